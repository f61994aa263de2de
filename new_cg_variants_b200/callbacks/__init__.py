"""Callback objects with the reference's names (numerical_experiments/callbacks/*.py).

Passed in ``callbacks=[...]`` to the solvers of ``new_cg_variants_b200.cg_variants``:

* the four standard ones -- ``error_A_norm``, ``residual_2_norm``, ``error_2_norm``,
  ``updated_residual_2_norm`` (the list of figure_gen.py:37) -- are *recognised by name*
  and computed ON THE GPU inside the solve, one fused matrix pass per iteration; the
  functions below are never called for them on that path;
* ``print_k(K)`` is honoured as a progress tick without forcing a device round trip;
* anything else (``save_x``, ``save_r``, or a user callable) makes the solver step the GPU
  one iteration at a time and call it with the reference's keyword protocol
  (``output, A, b, x_k, r_k, k, max_iter, kwargs, a_k1, a_k2, b_k1, r_k1, ...``).

The bodies here are plain host-side instrumentation with the reference's semantics, so
the same objects also work with any solver that follows the ``callback(**locals())``
protocol.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla

DEVICE_HISTORIES = ("error_A_norm", "residual_2_norm", "error_2_norm", "updated_residual_2_norm")


def _true_solution(kw):
    """x_true from the caller's kwargs, else a direct solve (error_A_norm.py:36-39)."""
    extra = kw["kwargs"]
    if "x_true" not in extra:
        A, b = kw["A"], kw["b"]
        solve = spla.spsolve if sps.issparse(A) else np.linalg.solve
        extra["x_true"] = solve(A.astype(np.double), b.astype(np.double))
    return extra["x_true"]


def _slot(kw, name, dtype=np.float64):
    out = kw["output"]
    if kw["k"] == 0:
        out[name] = np.zeros(kw["max_iter"], dtype=dtype)
    return out[name]


def error_A_norm(**kw):
    """sqrt(e.(A e)), e = x_k - x_true   (error_A_norm.py:47-48)."""
    A = kw["A"]
    e = kw["x_k"] - _true_solution(kw).astype(A.dtype)
    _slot(kw, "error_A_norm", A.dtype)[kw["k"]] = np.sqrt(e.T @ (A @ e))


def error_2_norm(**kw):
    """||x_k - x_true||_2   (error_2_norm.py:47-48)."""
    A = kw["A"]
    e = kw["x_k"] - _true_solution(kw).astype(A.dtype)
    _slot(kw, "error_2_norm", A.dtype)[kw["k"]] = np.linalg.norm(e)


def residual_2_norm(**kw):
    """||b - A x_k||_2   (residual_2_norm.py:41)."""
    _slot(kw, "residual_2_norm")[kw["k"]] = np.linalg.norm(kw["b"] - kw["A"] @ kw["x_k"])


def updated_residual_2_norm(**kw):
    """||r_k||_2 of the recursively updated residual (updated_residual_2_norm.py:40)."""
    _slot(kw, "updated_residual_2_norm")[kw["k"]] = np.linalg.norm(kw["r_k"])


def save_x(**kw):
    """output['x'][k] = x_k   (save_x.py)."""
    out, x, k = kw["output"], kw["x_k"], kw["k"]
    if k == 0:
        out["x"] = np.zeros((kw["max_iter"], len(x)), dtype=kw["A"].dtype)
    out["x"][k] = x


def save_r(**kw):
    """output['r'][k] = r_k   (save_r.py)."""
    out, r, k = kw["output"], kw["r_k"], kw["k"]
    if k == 0:
        out["r"] = np.zeros((kw["max_iter"], len(r)), dtype=kw["A"].dtype)
    out["r"][k] = r


def print_k(K):
    """Progress line every K iterations (print_k.py)."""
    def pk(**kw):
        if kw["k"] % K == 0:
            print(f"{kw['output']['name']}: iteration {kw['k']} of {kw['max_iter']}", end="\r")
    pk._cgx_print_every = K
    return pk


__all__ = ["error_A_norm", "error_2_norm", "residual_2_norm", "updated_residual_2_norm",
           "save_x", "save_r", "print_k", "DEVICE_HISTORIES"]
