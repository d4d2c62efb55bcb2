"""One process per GPU over NCCL (needs >= 2 GPUs; skipped on a 1-GPU box)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_spmd_operator_over_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 4 if n >= 4 else 2
    # own session: on a timeout the whole process group (torchrun and its workers) is killed, so a hung
    # exchange can never be left holding the GPUs
    p = subprocess.Popen([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(HERE, "_spmd_worker.py")],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = p.communicate(timeout=240)
    except subprocess.TimeoutExpired:
        import signal
        os.killpg(p.pid, signal.SIGKILL)
        out, err = p.communicate()
        raise AssertionError("spmd worker hung:\n" + out[-2000:] + err[-4000:])
    assert p.returncode == 0 and "spmd ok" in out, out[-3000:] + err[-3000:]
