"""Drop-in solver functions with the reference's names and signatures
(numerical_experiments/cg_variants/__init__.py:19-44,64-74):

    f(A, b, x0, max_iter, preconditioner=lambda x: x, callbacks=[], **kwargs) -> output dict

for ``hs_pcg, cg_pcg, gv_pcg, pr_pcg, m_pcg, pipe_pr_pcg, pipe_p_pcg, pipe_pr_m_pcg,
pipe_p_m_pcg`` and the un-preconditioned twins ``*_cg`` (no ``preconditioner`` argument).
``gv_*`` also accept ``w_replace`` (gv_cg.py:89); only the default "never" runs on the GPU.

Every iteration runs in libcgx_b200 on the GPU (there is no CPU solve path).  ``A`` may be
a scipy sparse matrix, a dense ndarray, or a ``PoissonStencil``.  ``preconditioner`` stays an
opaque callable as in the reference; it is probed once and must act as a fixed diagonal
scaling (identity or Jacobi, figure_gen.py:40-44), otherwise ``NotImplementedError``.

Extra keyword arguments understood here (all optional, ignored by the reference):
``device=0``, ``path="auto"|"stream"|"persistent"``, ``session=Session`` (reuse an operator
already resident on the GPU), ``return_info=True`` (adds ``output['_info']``).
"""
from __future__ import annotations

import weakref

import numpy as np
import scipy.sparse as sps

from .. import _lib
from .. import callbacks as _cbk
from ..callbacks import DEVICE_HISTORIES
from ..operators import PoissonStencil, canonical_csr
from ..session import Session

_OWN_KEYS = ("device", "path", "session", "return_info")

# ---------------------------------------------------------------------------------------
# preconditioner probe
# ---------------------------------------------------------------------------------------
def probe_preconditioner(preconditioner, n):
    """Return None if ``preconditioner`` is the identity, the vector ``d`` if it is the
    diagonal scaling ``v -> d*v`` (bit for bit on a random probe); raise otherwise."""
    if preconditioner is None:
        return None
    ones = np.ones(n)
    d = np.asarray(preconditioner(ones), dtype=np.float64)
    if d.shape != (n,):
        raise NotImplementedError("preconditioner must map (n,) -> (n,)")
    v = np.random.default_rng(0).standard_normal(n)
    pv = np.asarray(preconditioner(v), dtype=np.float64)
    if not np.array_equal(pv, d * v):
        raise NotImplementedError(
            "only diagonal (identity / Jacobi) preconditioners run on the GPU path; "
            "preconditioner(v) != preconditioner(ones) * v and there is no CPU fallback")
    if np.array_equal(d, ones):
        return None
    return d


# ---------------------------------------------------------------------------------------
# operator cache: keep the last few matrices resident in HBM between calls, as a
# figure_gen-style driver runs nine variants on the same A (figure_gen.py:50-60)
# ---------------------------------------------------------------------------------------
_CACHE: list = []          # entries: [weakref(A) | None, fingerprint, device, Session, dinv]
_CACHE_SIZE = 2


def _fingerprint(A):
    if isinstance(A, PoissonStencil):
        return ("stencil", A.dim, A.nx, A.ny, A.nz, A.diag, A.off)
    if sps.issparse(A):
        if A.format not in ("csr", "csc", "coo", "bsr", "dia"):
            A = A.tocsr()                      # e.g. LIL keeps object-dtype rows
        d = A.data
        step = max(1, d.shape[0] // 4096)
        return ("sparse", A.shape, A.nnz, A.format, float(d.sum()), float(d[::step].sum()))
    a = np.asarray(A)
    return ("dense", a.shape, float(a.sum()), float(a.reshape(-1)[:: max(1, a.size // 4096)].sum()))


def _session_for(A, dinv, device):
    fp = _fingerprint(A)
    for ent in _CACHE:
        ref, efp, edev, sess, edinv = ent
        same = ref is not None and ref() is A
        if isinstance(A, PoissonStencil):
            same = efp == fp
        if same and efp == fp and edev == device and sess._ctx:
            if (dinv is None) != (edinv is None) or (dinv is not None and not np.array_equal(dinv, edinv)):
                sess.set_jacobi(dinv)
                ent[4] = None if dinv is None else dinv.copy()
            _CACHE.remove(ent)
            _CACHE.insert(0, ent)
            return sess
    sess = Session(A, dinv=dinv, device=device)
    try:
        ref = weakref.ref(A)
    except TypeError:
        ref = None
    _CACHE.insert(0, [ref, fp, device, sess, None if dinv is None else dinv.copy()])
    while len(_CACHE) > _CACHE_SIZE:
        _CACHE.pop().__getitem__(3).close()
    return sess


def clear_cache():
    while _CACHE:
        _CACHE.pop()[3].close()


# ---------------------------------------------------------------------------------------
# the driver behind every exported function
# ---------------------------------------------------------------------------------------
def _split_callbacks(callbacks):
    device, ticks, generic = [], [], []
    for cb in callbacks:
        name = getattr(cb, "__name__", "")
        # the four standard callbacks: this package's objects, or the reference's own (module
        # `callbacks.<name>` of numerical_experiments) -- a user callable that merely shares the name
        # is called like any other
        mod = getattr(cb, "__module__", "") or ""
        std = name in DEVICE_HISTORIES and (cb is getattr(_cbk, name, None) or mod.split(".")[0] == "callbacks"
                                            or mod.endswith("callbacks." + name))
        if std:
            device.append(name)
        elif hasattr(cb, "_cgx_print_every") or name == "pk":
            ticks.append(cb)          # print_k(K): progress only
        else:
            generic.append(cb)
    return device, ticks, generic


def _ensure_x_true(A, b, extra, needed):
    """error_* histories need x_true; if the caller gave none, solve for it once on the
    host exactly as the reference callbacks do (error_A_norm.py:36-39)."""
    if not needed or "x_true" in extra:
        return
    import scipy.sparse.linalg as spla
    if isinstance(A, PoissonStencil):
        A = A.tocsr()
    solver = spla.spsolve if sps.issparse(A) else np.linalg.solve
    extra["x_true"] = solver(A.astype(np.double), np.asarray(b, dtype=np.double))


def _solve(name, tag, A, b, x0, max_iter, preconditioner, callbacks, kwargs):
    own = {k: kwargs.pop(k) for k in _OWN_KEYS if k in kwargs}
    device = own.get("device", 0)
    path = own.get("path", "auto")
    n = len(b)
    output = {"name": name, "max_iter": max_iter}
    dev_hist, ticks, generic = _split_callbacks(list(callbacks))
    _ensure_x_true(A, b, kwargs, any("error" in h for h in dev_hist))
    x_true = kwargs.get("x_true")

    sess = own.get("session")
    dinv = probe_preconditioner(preconditioner, n)
    if sess is None:
        sess = _session_for(A, dinv, device)
    else:
        sess.set_jacobi(dinv)                  # an explicit session still honours `preconditioner`

    if not generic:
        _, hist, info = sess.solve(tag, b, x0, max_iter, x_true=x_true, histories=tuple(dev_hist),
                                   path=path, return_x=False)
        for h in dev_hist:
            output[h] = hist[h]
    else:
        info = _solve_stepwise(sess, tag, A, b, x0, max_iter, x_true, dev_hist, generic, output,
                               kwargs, path)
    for cb in ticks:   # leave the terminal as print_k would after the last iteration
        try:
            cb(output=output, k=max_iter - 1, max_iter=max_iter)
        except Exception:
            pass
    if own.get("return_info"):
        output["_info"] = info
    return output


def _solve_stepwise(sess, tag, A, b, x0, max_iter, x_true, dev_hist, generic, output, extra, path):
    """Arbitrary callbacks: step the GPU one iteration at a time and hand each callback the
    reference's keyword set (hs_cg.py:97-98,128-129).  Slow (a device round trip per
    iteration) but still no CPU arithmetic on the solve itself."""
    sess.load_problem(b, x0, x_true)
    sess.begin(tag, max_iter, histories=tuple(dev_hist), path=path)
    predicted = tag in ("pr", "m") or tag.startswith("pipe")
    b_arr = np.asarray(b, dtype=np.float64)
    a_k1 = a_k2 = 0.0
    b_k = b_k1 = 0.0
    x_k1 = r_k1 = None
    sc = sess.scalars()
    next_b = sc["b"]
    for k in range(max_iter):
        if k > 0:
            sess.advance(1)
            sc_new = sess.scalars()
            a_k2, a_k1 = a_k1, sc["a"]
            b_k1 = b_k
            b_k = next_b if predicted else sc_new["b"]
            sc = sc_new
            next_b = sc["b"]
        x_k = sess.vector("x")
        r_k = sess.vector("r")
        loc = dict(output=output, A=A, b=b_arr, x0=x0, x_k=x_k, r_k=r_k, k=k, max_iter=max_iter,
                   n=len(b_arr), kwargs=extra, a_k=sc["a"], a_k1=a_k1, a_k2=a_k2, b_k=b_k, b_k1=b_k1,
                   nu_k=sc["nu"], mu_k=sc["mu"], x_k1=x_k1, r_k1=r_k1)
        for cb in generic:
            cb(**loc)
        x_k1, r_k1 = x_k, r_k
    _, hist = sess.fetch(want_x=False, want_hist=True)
    for h in dev_hist:
        output[h] = hist[_lib.HIST_NAMES.index(h)].copy()
    return sess.get_info()


def _never(**kwargs):
    return False


def _make(name, tag, preconditioned, gv=False):
    if preconditioned and gv:
        def f(A, b, x0, max_iter, w_replace=_never, preconditioner=lambda x: x, callbacks=[], **kwargs):
            _check_w_replace(w_replace)
            return _solve(name, tag, A, b, x0, max_iter, preconditioner, callbacks, kwargs)
    elif preconditioned:
        def f(A, b, x0, max_iter, preconditioner=lambda x: x, callbacks=[], **kwargs):
            return _solve(name, tag, A, b, x0, max_iter, preconditioner, callbacks, kwargs)
    elif gv:
        def f(A, b, x0, max_iter, w_replace=_never, callbacks=[], **kwargs):
            _check_w_replace(w_replace)
            return _solve(name, tag, A, b, x0, max_iter, None, callbacks, kwargs)
    else:
        def f(A, b, x0, max_iter, callbacks=[], **kwargs):
            return _solve(name, tag, A, b, x0, max_iter, None, callbacks, kwargs)
    f.__name__ = f.__qualname__ = name
    f.__doc__ = f"{name}: GPU implementation of the reference's `{name}` (variant tag {tag!r})."
    return f


def _check_w_replace(w_replace):
    if w_replace is not _never:
        # the reference's default is `lambda **kwargs: False` (gv_cg.py:89); accept any
        # callable that declines at k=1 without looking at vectors, reject the rest
        try:
            if not w_replace(k=1, wk_replace_flags={}):
                return
        except Exception:
            pass
        raise NotImplementedError("gv residual replacement (w_replace) is not supported on the GPU path")


_TAGS = [("hs", "hs"), ("cg", "cg"), ("gv", "gv"), ("pr", "pr"), ("m", "m"), ("pipe_pr", "pipe_pr"),
         ("pipe_p", "pipe_p"), ("pipe_pr_m", "pipe_pr_m"), ("pipe_p_m", "pipe_p_m")]
__all__ = ["probe_preconditioner", "clear_cache"]
for _stem, _tag in _TAGS:
    for _suffix, _pre in (("_pcg", True), ("_cg", False)):
        _fname = _stem + _suffix
        globals()[_fname] = _make(_fname, _tag, _pre, gv=(_tag == "gv"))
        __all__.append(_fname)
del _stem, _tag, _suffix, _pre, _fname
