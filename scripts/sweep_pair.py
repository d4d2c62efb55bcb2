"""Step time (forward + adjoint back to back, as bench.py times them) of the benchmark operator for a
list of schedule settings, inside one process, alternating over several trials.
    python scripts/sweep_pair.py <blocks> [steps] [setting ...]
A setting is comma-separated name=value pairs of context parameters, e.g. "auto_tile=0,fwd_tc=448,prefetch=0"."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
nblk = int(sys.argv[1])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
settings = sys.argv[3:] or ["prefetch=0", "prefetch=65536"]
DEFAULTS = {"persistent": 0, "l2_keep": 96, "wave_fit": 1, "auto_tile": 1, "fwd_tc": 512, "adj_tc": 64, "fwd_tx": 64, "adj_ru": 4, "prefetch": 32768, "pdl": 1,
            "ld_policy": 1}
ctx = dj.init(world=1, devices=[0])
rng = np.random.default_rng(0)
pool = [np.asfortranarray(rng.random((4096, 4096), dtype=np.float32)) for _ in range(4)]
A = dj.JopDBlock(lambda i, j: dj.JopDense(pool[i % 4]) if i == j else dj.JopZeroBlock(np.float32, 4096, 4096),
                 dims=(nblk, nblk), pids=[0], dist=[1, 1], isdiag=True, shapes=([(4096,)] * nblk, [(4096,)] * nblk),
                 dtype=np.float32)
m, d, m2 = A.domain().rand(rng), A.range().zeros(), A.domain().zeros()
nbytes = 2 * A.algorithmic_bytes()


def run():
    for _ in range(5):
        A.mul(d, m); A.mul_adjoint(m2, d)
    ctx.sync()
    ctx.timer_start()
    for _ in range(steps):
        A.mul(d, m); A.mul_adjoint(m2, d)
    return ctx.timer_stop() / steps


best, tiles = {}, {}
for trial in range(3):
    for s in settings:
        for k, v in DEFAULTS.items():
            ctx.set_param(k, v)
        for kv in s.split(","):
            k, v = kv.split("=")
            ctx.set_param(k, int(v))
        A.replan()
        tiles[s] = (A.schedule_info(False)[0], A.schedule_info(True)[0])
        ms = run()
        best[s] = min(best.get(s, 1e9), ms)
for s in settings:
    ms = best[s]
    print(f"{s:52s} tiles={tiles[s]} {ms * 1e3:8.1f} us/step {nbytes / ms / 1e6:8.1f} GB/s", flush=True)
