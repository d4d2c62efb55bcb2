"""One launch of each DBArray kernel on a 2^28-element Float64 vector (2 GiB per operand), for `ncu --set full`:
eval (unary, three inputs), specialised lincomb, one-launch reductions (dot, norm)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402

dj = djets_loader.load()
ctx = dj.init(world=1, devices=[0])
nb, bl = 256, 1 << 20
R = dj.JetDSpace([[(bl,)] * nb], [0], np.float64, ctx)
x, y, z, out = R.zeros().fill(1.5), R.zeros().fill(-2.0), R.zeros().fill(0.25), R.zeros()
L = dj.lazy
for _ in range(2):                                                   # second round is the one to look at (warm)
    out.eval_(abs(L(x)), wide=False)
    out.eval_(L(x) * L(y) / (abs(L(z)) + 1.0), wide=False)
    out.lincomb_([0.1, 0.2, 0.3], [x, y, z])
    x.dot(y)
    x.norm()
ctx.sync()
