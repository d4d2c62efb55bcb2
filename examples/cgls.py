"""CGLS (conjugate gradients on the normal equations) written against the public interface only:
``mul!`` / adjoint ``mul!``, ``dot`` / ``norm`` and axpy-style broadcasts on DBArrays -- the solver
loop BASELINE config 5 names.  Everything stays on the GPUs; only scalars come back to the host.

    python examples/cgls.py            (needs a B200; uses 4 ranks on the visible GPUs)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import djets_loader  # noqa: E402


def cgls(A, b, iters=50, tol=1e-10):
    """min ||A x - b||_2.  Per iteration: one forward, one adjoint, two dots, three fused updates."""
    x = A.domain().zeros()
    r = b.copy()                                   # r = b - A x
    s = A.T @ r
    p = s.copy()
    q = A.range().zeros()
    gamma = float(s.dot(s))
    gamma0 = gamma
    history = [float(r.norm())]
    for _ in range(iters):
        A.mul(q, p)                                # q = A p
        alpha = gamma / float(q.dot(q))
        x.lincomb_([1.0, alpha], [x, p])           # x .+= alpha .* p
        r.lincomb_([1.0, -alpha], [r, q])          # r .-= alpha .* q
        A.mul_adjoint(s, r)                        # s = A' r
        gamma_new = float(s.dot(s))
        history.append(float(r.norm()))
        if gamma_new <= tol * tol * gamma0:
            break
        p.lincomb_([1.0, gamma_new / gamma], [s, p])   # p .= s .+ beta .* p
        gamma = gamma_new
    return x, history


def cgls_graph(A, b, iters=50, check_every=10, tol=1e-10):
    """The same iteration with NO host round trip: dots stay on the devices (``dot_s``), alpha and beta are
    formed there (``Scalar`` arithmetic), the updates read them as coefficients, and the whole iteration is
    captured once and replayed with one launch.  The host looks at the residual every ``check_every`` steps."""
    dj = sys.modules["djets_b200"]
    ctx, T = A.ctx, A.dtype
    x = A.domain().zeros()
    r = b.copy()
    s = A.T @ r
    p = s.copy()
    q = A.range().zeros()
    gamma, gamma_new, delta, alpha, beta, rnorm = (dj.Scalar(0.0, T, ctx) for _ in range(6))
    s.dot_s(s, gamma)
    gamma0 = float(gamma)
    history = [float(r.norm())]
    SC = sys.modules["djets_b200._lib"]
    with ctx.capture() as step:
        A.mul(q, p)                                            # q = A p
        q.dot_s(q, delta)
        alpha.assign(SC.SC_DIV, gamma, delta)                  # alpha = gamma / (q . q)
        x.assign(dj.lazy(x) + alpha * dj.lazy(p))              # x .+= alpha .* p
        r.assign(dj.lazy(r) - alpha * dj.lazy(q))              # r .-= alpha .* q
        A.mul_adjoint(s, r)                                    # s = A' r
        s.dot_s(s, gamma_new)
        beta.assign(SC.SC_DIV, gamma_new, gamma)
        p.assign(dj.lazy(s) + beta * dj.lazy(p))               # p .= s .+ beta .* p
        gamma.assign(SC.SC_COPY, gamma_new)
        r.norm_s(2, rnorm)
    for it in range(1, iters + 1):
        step.launch()
        if it % check_every == 0 or it == iters:
            history.append(float(rnorm))
            if float(gamma) <= tol * tol * gamma0:
                break
    return x, history


if __name__ == "__main__":
    dj = djets_loader.load()
    dj.init(world=4)
    rng = np.random.default_rng(0)
    nb, bs = 8, 256
    kind = lambda i, j: (i % 2 == 0) and (j % 2 == 1)      # noqa: E731  diagonal blocks, benchmark/benchmarks.jl:83-89
    A = dj.blockop(lambda i, j: dj.JopDiagonal(1.0 + rng.random(bs)) if kind(i, j)
                   else dj.JopDense(np.asfortranarray(rng.random((bs, bs)) / bs + (np.eye(bs) if i == j else 0))),
                   dims=(nb, nb), dist=[2, 2])
    x_true = dj.rand(A.domain(), rng)
    b = A @ x_true
    x, hist = cgls(A, b, iters=200)
    err = float((x - x_true).norm()) / float(x_true.norm())
    print(f"{len(hist) - 1} iterations, residual {hist[0]:.3e} -> {hist[-1]:.3e}, solution error {err:.3e}")
    x, hist = cgls_graph(A, b, iters=200)
    err = float((x - x_true).norm()) / float(x_true.norm())
    print(f"captured iteration: residual {hist[0]:.3e} -> {hist[-1]:.3e}, solution error {err:.3e}")
