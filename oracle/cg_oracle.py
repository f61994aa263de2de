"""CPU oracle for the predict-and-recompute CG hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``new_cg_variants_b200/`` imports it.

It is a numpy/scipy restatement of the reference solvers in
``predict_and_recompute/numerical_experiments/cg_variants/`` (reference paths below are
relative to ``/root/reference/predict_and_recompute/``):

  * every floating-point expression is evaluated with the same operands, in the same
    association order, with the same primitives (``scipy CSR @``, ``numpy @``, numpy
    elementwise ops) as the reference, so on one machine the histories are
    BIT-IDENTICAL to the reference's (pinned by ``tests/golden/make_golden.py``, which
    runs the real reference from ``/root/reference`` and stores its outputs as
    fixtures; ``tests/test_oracle.py`` re-checks the oracle against them);
  * the structure is NOT the reference's: one table-driven stepper instead of ten
    hand-unrolled functions, no per-iteration ``np.copy`` "update indexing" blocks, no
    ``callback(**locals())`` protocol.

Parity status: pinned end-to-end (per-iteration histories of all nine variants on the
fixtures in ``tests/golden/``); at the level of a single primitive (one ``A@v`` or one
``u@v``) the reference itself is unpinned.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps

# variant tag -> reference function name (numerical_experiments/cg_variants/__init__.py:64-74)
VARIANTS = {
    "hs": "hs_pcg",            # hs_cg.py:70-131
    "cg": "cg_pcg",            # cg_cg.py:77-146
    "gv": "gv_pcg",            # gv_cg.py:89-176
    "pr": "pr_pcg",            # pr_cg.py:93-171
    "m": "m_pcg",              # pr_cg.py:93-164,172-176
    "pipe_pr": "pipe_pr_pcg",  # pipe_pr_cg.py:109-193,201-205
    "pipe_p": "pipe_p_pcg",    # pipe_pr_cg.py:195-199
    "pipe_pr_m": "pipe_pr_m_pcg",  # pipe_pr_cg.py:213-217
    "pipe_p_m": "pipe_p_m_pcg",    # pipe_pr_cg.py:207-211
}
HISTORIES = ("error_A_norm", "residual_2_norm", "error_2_norm", "updated_residual_2_norm")


# --------------------------------------------------------------------------------------
# problem generators (figure_gen.py:31-44 for the set-up)
# --------------------------------------------------------------------------------------
def poisson2d(nx, ny=None):
    """5-point Dirichlet Laplacian, natural (x-fastest) ordering: kron(I,T)+kron(T,I).

    ``matrices/poisson_ca.mtx`` is this matrix for nx=ny=16."""
    ny = nx if ny is None else ny
    tx = sps.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(nx, nx))
    ty = sps.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(ny, ny))
    a = sps.kron(sps.identity(ny), tx) + sps.kron(ty, sps.identity(nx))
    a = sps.csr_matrix(a)
    a.sort_indices()
    return a


def poisson3d(nx, ny=None, nz=None):
    """7-point Dirichlet Laplacian (diag 6, off -1), natural ordering i = x + nx*(y + ny*z)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    t = lambda m: sps.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(m, m))
    eye = sps.identity
    a = (sps.kron(eye(nz), sps.kron(eye(ny), t(nx)))
         + sps.kron(eye(nz), sps.kron(t(ny), eye(nx)))
         + sps.kron(t(nz), sps.kron(eye(ny), eye(nx))))
    a = sps.csr_matrix(a)
    a.sort_indices()
    return a


def model_problem_spectrum(n, kappa=1e6, rho=0.9):
    """Diagonal model problem of scaling_experiments_mpi4py/scaling_tests.py:31-36."""
    lam1, lamn = 1.0 / kappa, 1.0
    return lam1 + (lamn - lam1) * np.arange(n) / (n - 1) * rho ** np.arange(n - 1, -1, -1, dtype="float")


def setup_problem(A):
    """x_true = 1/sqrt(N), b = A x_true, x0 = 0  (figure_gen.py:31-34)."""
    n = A.shape[0]
    x_true = np.ones(n) / np.sqrt(n)
    b = A @ x_true
    x0 = np.zeros(n)
    return x_true, b, x0


def jacobi_dinv(A):
    """The reference's Jacobi lambda is ``(1/A.diagonal())*x`` (figure_gen.py:43):
    reciprocal first, then an elementwise product."""
    return 1 / A.diagonal()


# --------------------------------------------------------------------------------------
# the instrumentation the four standard callbacks compute
# --------------------------------------------------------------------------------------
def _record(hist, k, A, b, x, r, x_true):
    if hist is None:
        return
    if x_true is not None:
        e = x - x_true                                    # error_A_norm.py:47, error_2_norm.py:47
        hist["error_A_norm"][k] = np.sqrt(e.T @ (A @ e))  # error_A_norm.py:48  (A@e, not A@x-b)
        hist["error_2_norm"][k] = np.linalg.norm(e)       # error_2_norm.py:48
    hist["residual_2_norm"][k] = np.linalg.norm(b - A @ x)  # residual_2_norm.py:41
    hist["updated_residual_2_norm"][k] = np.linalg.norm(r)  # updated_residual_2_norm.py:40


# --------------------------------------------------------------------------------------
# the solver: one stepper for all nine preconditioned variants
# --------------------------------------------------------------------------------------
def _blas_dot(u, v):
    return u @ v


def _pairwise_dot(u, v):
    return np.sum(u * v)                     # numpy pairwise summation, products rounded


def _reversed_dot(u, v):
    return u[::-1] @ v[::-1]                 # strided BLAS path, opposite order


def _lanes8_dot(u, v):
    t = u * v                                # 8 interleaved accumulators, like a SIMD ddot
    m = (t.shape[0] // 8) * 8
    return (t[:m].reshape(-1, 8).sum(axis=0).sum() + t[m:].sum()) if m else t.sum()


def _strided_dot(u, v):
    t = u * v                                # 1024 strided sequential partials, then a tree: the shape of a
    m = (t.shape[0] // 1024) * 1024          # grid-stride GPU reduction
    return (t[:m].reshape(-1, 1024).sum(axis=0).sum() + t[m:].sum()) if m else t.sum()


def _extended_dot(u, v):
    return float(u.astype(np.longdouble) @ v.astype(np.longdouble))   # nearly exact


# alternative summation orders for the rounding-sensitivity ensemble
DOT_ORDERS = {"blas": _blas_dot, "pairwise": _pairwise_dot, "reversed": _reversed_dot,
              "lanes8": _lanes8_dot, "strided1024": _strided_dot, "extended": _extended_dot}


def solve(variant, A, b, x0, max_iter, dinv=None, x_true=None, history=True, return_state=False,
          dot=None, w_replace=None):
    """Run ``max_iter-1`` iterations of ``variant`` exactly as the reference does.

    dinv: None for the identity preconditioner, else the vector the Jacobi lambda
    multiplies by.  Returns the reference's ``output`` dict (name, max_iter and the four
    history arrays, index 0 = initial state); with ``return_state`` also the final
    vectors/scalars (used by the single-iteration kernel tests).

    w_replace: GV only -- the reference's residual-replacement predicate (gv_cg.py:89,156-158),
    called with the same keywords; None = never (its default).

    dot: inner-product implementation; the default is numpy's ``u @ v`` (what the reference
    calls).  Other summation orders (``DOT_ORDERS``) are used ONLY to measure how sensitive
    the reference's own curves are to rounding order (tests/golden/make_golden.py).
    """
    if variant not in VARIANTS:
        raise ValueError(f"unknown variant {variant!r}")
    if dot is None:
        dot = _blas_dot          # the reference's `u @ v`
    M = (lambda v: v) if dinv is None else (lambda v: dinv * v)
    pipe = variant.startswith("pipe")
    meurant = variant == "m" or variant.endswith("_m")
    recompute_w = variant in ("pipe_pr", "pipe_pr_m")       # pipe_pr_cg.py:181-182
    out = {"name": VARIANTS[variant], "max_iter": max_iter}
    hist = None
    if history:
        hist = {h: np.zeros(max_iter) for h in HISTORIES if x_true is not None or "error" not in h}
        out.update(hist)

    # ---- initialisation (hs_cg.py:83-94, cg_cg.py:90-104, gv_cg.py:105-121,
    #      pr_cg.py:106-120, pipe_pr_cg.py:122-140)
    x = np.copy(x0)
    r = np.copy(b - A @ x)
    rt = M(r)
    p = np.copy(rt)
    w = wt = s = st = u = ut = None
    eta = dl = gam = None
    if variant == "hs":
        nu = dot(r, rt)
        s = A @ p
        mu = dot(p, s)
    elif variant == "cg":
        w = A @ rt
        nu = dot(r, rt)
        eta = dot(w, rt)
        s = A @ p
        mu = dot(p, s)
    elif variant == "gv":
        w = A @ rt
        wt = M(w)
        s = np.copy(w)
        st = np.copy(wt)
        u = A @ wt
        nu = dot(r, rt)
        eta = dot(w, r)            # gv_cg.py:115 (never consumed)
        mu = dot(p, s)
    elif not pipe:             # pr / m
        nu = dot(rt, r)
        s = A @ p
        st = M(s)
        mu = dot(p, s)
        dl = dot(r, st)
        gam = dot(st, s)
    else:                      # pipe_* family
        nu = dot(rt, r)
        s = A @ p
        st = M(s)
        w = np.copy(s)
        wt = np.copy(st)
        u = A @ st
        ut = M(u)
        mu = dot(p, s)
        dl = dot(r, st)
        gam = dot(st, s)
    a = nu / mu
    beta = 0
    _record(hist, 0, A, b, x, r, x_true)

    wk_flags = {}                  # gv_cg.py:123: scratch dict handed to the w_replace predicate
    for k in range(1, max_iter):
        a1, nu1 = a, nu
        if variant == "hs":                                   # hs_cg.py:117-125
            x = x + a1 * p
            r = r - a1 * s
            rt = M(r)
            nu = dot(r, rt)
            beta = nu / nu1
            p = rt + beta * p
            s = A @ p
            mu = dot(p, s)
        elif variant == "cg":                                 # cg_cg.py:130-140
            x = x + a1 * p
            r = r - a1 * s
            rt = M(r)
            w = A @ rt
            nu = dot(r, rt)
            eta = dot(w, rt)
            beta = nu / nu1
            p = rt + beta * p
            s = w + beta * s
            mu = eta - (beta / a1) * nu
        elif variant == "gv":                                 # gv_cg.py:151-170
            r_prev = r
            x = x + a1 * p
            r = r - a1 * s
            rt = rt - a1 * st
            w = w - a1 * u
            if w_replace is not None and w_replace(k=k, A=A, b=b, x=x, w=w, r=r, r_=r_prev, u=u, s=s, p=p,
                                                   wk_replace_flags=wk_flags):
                w = A @ r                                     # gv_cg.py:156-158
            wt = M(w)
            t = A @ wt
            nu = dot(r, rt)
            eta = dot(w, rt)
            beta = nu / nu1
            p = rt + beta * p
            s = w + beta * s
            st = wt + beta * st
            u = t + beta * u
            mu = eta - (beta / a1) * nu
        elif not pipe:                                        # pr_cg.py:146-158
            dl1, gam1 = dl, gam
            x = x + a1 * p
            r = r - a1 * s
            rt = rt - a1 * st
            nu = -nu1 + a1 ** 2 * gam1 if meurant else nu1 - 2 * a1 * dl1 + a1 ** 2 * gam1
            beta = nu / nu1
            p = rt + beta * p
            s = A @ p
            st = M(s)
            mu = dot(p, s)
            dl = dot(r, st)
            gam = dot(st, s)
            nu = dot(rt, r)
        else:                                                 # pipe_pr_cg.py:169-187
            dl1, gam1 = dl, gam
            x = x + a1 * p
            r = r - a1 * s
            rt = rt - a1 * st
            w = w - a1 * u
            wt = wt - a1 * ut
            nu = -nu1 + a1 ** 2 * gam1 if meurant else nu1 - 2 * a1 * dl1 + a1 ** 2 * gam1
            beta = nu / nu1
            p = rt + beta * p
            s = w + beta * s
            st = wt + beta * st
            u = A @ st
            ut = M(u)
            if recompute_w:
                w = A @ rt
                wt = M(w)
            mu = dot(p, s)
            dl = dot(r, st)
            gam = dot(st, s)
            nu = dot(rt, r)
        a = nu / mu
        _record(hist, k, A, b, x, r, x_true)

    if return_state:
        out["_state"] = dict(x=x, r=r, rt=rt, p=p, s=s, st=st, w=w, wt=wt, u=u, ut=ut,
                             nu=nu, mu=mu, a=a, beta=beta, eta=eta, delta=dl, gamma=gam)
    return out


# --------------------------------------------------------------------------------------
# summary metrics of figure_gen.py:80-89
# --------------------------------------------------------------------------------------
def convergence_metrics(error_A_norm, tol=1e-5):
    """(iterations to rel. A-norm error <= tol [0 = never], log10 of the minimum)."""
    rel = np.asarray(error_A_norm) / error_A_norm[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        return int(np.argmin(rel > tol)), float(np.log10(np.nanmin(rel)))


def departure_index(ref_hist, exact_hist, tol=1e-10):
    """k* of SURVEY.md section 8c: first k at which the finite-precision curve leaves
    the extended-precision (``exact_pcg``) curve by more than ``tol`` relative; searched
    over exact_pcg's non-zero prefix (it breaks early, exact_cg.py:149-150)."""
    ex = np.asarray(exact_hist, dtype=np.float64)
    nz = np.nonzero(ex)[0]
    m = min(len(ref_hist), (nz[-1] + 1) if len(nz) else 0)
    if m == 0:
        return 0
    rel = np.abs(np.asarray(ref_hist[:m], dtype=np.float64) - ex[:m]) / ex[:m]
    bad = np.nonzero(rel > tol)[0]
    return int(bad[0]) if len(bad) else int(m)


# --------------------------------------------------------------------------------------
# mpi4py-style fixed-iteration variants (scaling_experiments_mpi4py/cg_variants/*.py)
# restated for ONE rank (Allreduce == copy) on a diagonal operator; used as KATs.
# --------------------------------------------------------------------------------------
def solve_mpi_style(variant, lam, b, max_iter):
    """Un-preconditioned, x0 = 0, fixed ``max_iter`` iterations, no history.
    ``lam`` is the diagonal of the (diagonal) model matrix.  Returns x.

    hs: hs_cg.py:36-60   cg: cg_cg.py:46-68   gv: gv_cg.py:51-77
    pr: pr_cg.py:49-73   pipe_pr: pipe_pr_cg.py:58-83."""
    A = lambda v: lam * v
    x = np.zeros_like(b)
    if variant == "hs":
        r = np.copy(b); p = np.zeros_like(b); nu = 1.0
        for _ in range(max_iter):
            nu_ = nu
            nu = r @ r
            beta = nu / nu_
            p = p * beta; p = p + r
            s = A(p)
            mu = p @ s
            alpha = nu / mu
            x = x + alpha * p
            r = r - alpha * s
        return x
    if variant in ("cg", "gv"):
        r = np.copy(b); p = np.zeros_like(b); s = np.zeros_like(b)
        nu = 1.0; alpha = 0.0                      # np.ones / np.zeros initial scalars
        if variant == "gv":
            w = A(r); u = np.zeros_like(b)         # gv_cg.py:42-44
        for k in range(max_iter):
            if variant == "cg":
                w = A(r)
            nu_ = nu
            nu = r @ r; eta = r @ w
            if variant == "gv":
                t = A(w)
            beta = nu / nu_
            p = p * beta; p = p + r
            s = s * beta; s = s + w
            if variant == "gv":
                u = u * beta; u = u + t
            mu = eta - (beta / alpha) * nu if k > 0 else eta
            alpha = nu / mu
            x = x + alpha * p
            r = r - alpha * s
            if variant == "gv":
                w = w - alpha * u
        return x
    if variant in ("pr", "pipe_pr"):
        r = np.copy(b); p = np.copy(b); s = A(r)
        w = None
        for _ in range(max_iter):
            mu = p @ s; delta = r @ s; gamma = s @ s; nu_ = r @ r
            if variant == "pipe_pr":
                wp = A(r); u = A(s)
            alpha = nu_ / mu
            x = x + alpha * p
            r = r - alpha * s
            if variant == "pipe_pr":
                w = wp - alpha * u
            nu = nu_ - 2 * alpha * delta + alpha ** 2 * gamma
            beta = nu / nu_
            p = p * beta; p = p + r
            if variant == "pipe_pr":
                s = s * beta; s = s + w
            else:
                s = A(p)
        return x
    raise ValueError(variant)
