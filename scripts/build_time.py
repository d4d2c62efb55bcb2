"""How long it takes to build operators (upload of dense block payloads from pageable host memory)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
ctx = dj.init(world=1, devices=[0])
rng = np.random.default_rng(0)
pool = [np.asfortranarray(rng.random((1024, 1024), dtype=np.float32)) for _ in range(4)]
for nb in (32, 64):
    t0 = time.perf_counter()
    A = dj.JopDBlock(lambda i, j: dj.JopDense(pool[(i + j) % 4]), dims=(nb, nb), pids=[0], dist=[1, 1],
                     shapes=([(1024,)] * nb, [(1024,)] * nb), dtype=np.float32)
    dt = time.perf_counter() - t0
    gb = nb * nb * 4 / 1024
    print(f"{nb}x{nb} blocks of 4 MiB: {dt:.2f} s, {gb / dt:.2f} GiB/s, {dt / (nb * nb) * 1e6:.0f} us per block")
    A.close()
big = np.asfortranarray(rng.random((8192, 8192), dtype=np.float32))
t0 = time.perf_counter()
A = dj.JopDBlock(lambda i, j: dj.JopDense(big), dims=(4, 4), pids=[0], dist=[1, 1], shapes=([(8192,)] * 4, [(8192,)] * 4), dtype=np.float32)
dt = time.perf_counter() - t0
print(f"4x4 blocks of 256 MiB: {dt:.2f} s, {4.0 / dt:.2f} GiB/s")
