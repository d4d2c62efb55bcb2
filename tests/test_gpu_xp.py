"""One process per rank without NCCL: host all-gather (gloo) for the control plane, CUDA-IPC peer memory and epoch
flags for the data plane -- the cross-process exchange of a grid operator (src/DistributedJets.jl:600-604, 647-651,
694-698 in the reference).  Kernels of different ranks wait on one another's flags, so every rank needs a GPU of its
own: the test needs >= 2 GPUs (the library itself refuses the peer-memory path when ranks share a GPU)."""
import os
import signal
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_multiprocess_exchange_over_peer_memory():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs one GPU per rank (>= 2 GPUs): kernels of different ranks wait on each other's flags")
    world, port = (4 if ngpu >= 4 else 2), free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_xp_worker.py")], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True, start_new_session=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=300)[0])
    except subprocess.TimeoutExpired:
        for p in procs:
            try:
                os.killpg(p.pid, signal.SIGKILL)
            except ProcessLookupError:
                pass
        raise AssertionError("xp worker hung")
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank}:\n{out[-4000:]}"
    assert f"xp ok {world}" in outs[0]
