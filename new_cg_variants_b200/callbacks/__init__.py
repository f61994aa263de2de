"""Callback objects with the reference's names (numerical_experiments/callbacks/*.py).

Passed in ``callbacks=[...]`` to the solvers of ``new_cg_variants_b200.cg_variants``:

* the four standard ones -- ``error_A_norm``, ``residual_2_norm``, ``error_2_norm``,
  ``updated_residual_2_norm`` (the list of figure_gen.py:37) -- are *recognised by name*
  and computed ON THE GPU inside the solve, one fused matrix pass per iteration; the
  functions below are never called for them on that path;
* ``print_k(K)`` is honoured as a progress tick without forcing a device round trip;
* ``save_x``, ``save_r``, ``lanczos_recurrence`` and ``updated_error_A_norm`` are served from
  device capture buffers: the GPU records x_k / r_k / (a, b) after every iteration and the
  bodies below run once on the host afterwards, fed from one device-to-host copy;
* any other callable makes the solver step the GPU one iteration at a time and call it with
  the reference's keyword protocol (``output, A, b, x_k, r_k, k, max_iter, kwargs, a_k1, a_k2,
  b_k1, r_k1, ...``); further state vectors (``p_k, s_k, rt_k, st_k, w_k, wt_k, u_k, t_k``) are
  copied off the device for callables that name them as parameters.

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


def updated_error_A_norm(**kw):
    """sqrt(r_k . A^{-1} r_k): the A-norm of the error the algorithm "sees" through its updated
    residual (updated_error_A_norm.py:43-45; a direct solve per iteration, as in the reference)."""
    A, r = kw["A"], kw["r_k"]
    solve = spla.spsolve if sps.issparse(A) else np.linalg.solve
    e = solve(A.astype(np.double), r.astype(np.double))
    _slot(kw, "updated_error_A_norm")[kw["k"]] = np.sqrt(e.T @ r)


def lanczos_recurrence(**kw):
    """Lanczos vectors / coefficients implied by the CG iterates and the loss of the three-term
    recurrence and of orthogonality (lanczos_recurrence.py:43-77)."""
    out, max_iter, k, r = kw["output"], kw["max_iter"], kw["k"], kw["r_k"]
    a_k1, b_k1 = kw["a_k1"], kw["b_k1"]
    A = kw["A"]
    if k == 0:
        out["lanczos_alpha"] = np.zeros(max_iter, dtype=A.dtype)
        out["lanczos_beta"] = np.zeros(max_iter, dtype=A.dtype)
        out["lanczos_z"] = np.zeros((len(r), max_iter), dtype=A.dtype)
        out["lanczos_z"][:, 0] = r / np.linalg.norm(r)
    elif k < max_iter:
        r_k1, a_k2 = kw["r_k1"], kw["a_k2"]
        out["lanczos_alpha"][k - 1] = 1 / a_k1 + b_k1 / a_k2 if k > 1 else 1 / a_k1
        out["lanczos_beta"][k - 1] = np.linalg.norm(r) / (a_k1 * np.linalg.norm(r_k1))
        out["lanczos_z"][:, k] = (-1) ** k * r / np.linalg.norm(r)
    if k == max_iter - 1:
        _lanczos_finish(out, A, max_iter)


def _lanczos_finish(out, A, max_iter):
    """lanczos_recurrence.py:63-77: T (max_iter x max_iter-1), E = A Z[:, :-1] - Z T."""
    T = sps.diags([out["lanczos_alpha"], out["lanczos_beta"][:max_iter - 2], out["lanczos_beta"][:max_iter - 1]],
                  [0, 1, -1], shape=(max_iter, max_iter - 1))
    Z = out["lanczos_z"]
    E = A @ Z[:, :-1] - Z @ T
    out["lanczos_3_term_error"] = np.linalg.norm(E, axis=0)
    out["lanczos_orthogonality"] = np.abs(np.einsum("ji,ji->i", out["lanczos_beta"][:max_iter - 1] * Z[:, :-1], Z[:, 1:]))


# Callbacks served from the DEVICE capture buffers (x_k / r_k / scalars recorded on the GPU during
# the solve, one copy at the end) instead of a host round trip per iteration.
CAPTURE_CALLBACKS = ("save_x", "save_r", "lanczos_recurrence", "updated_error_A_norm")


def print_k(K):
    """Progress line every K iterations (print_k.py)."""
    def pk(**kw):
        if kw["k"] % K == 0:
            print(f"{kw['output']['name']}: iteration {kw['k']} of {kw['max_iter']}", end="\r")
    pk._cgx_print_every = K
    return pk


__all__ = ["error_A_norm", "error_2_norm", "residual_2_norm", "updated_residual_2_norm",
           "save_x", "save_r", "lanczos_recurrence", "updated_error_A_norm", "print_k", "DEVICE_HISTORIES",
           "CAPTURE_CALLBACKS"]
