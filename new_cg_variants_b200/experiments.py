"""Matplotlib-free restatement of the reference's numerical-experiment driver
(numerical_experiments/figure_gen.py:21-124), running every solve on the GPU path:

    test_matrix(A, max_iter, title, preconditioner=None|'jacobi', variants=[...], data_dir=...)
    parse_convergence_data(matrix_name, preconditioner, variants, A=..., data_dir=...)
    gen_convergence_table(data_dir, out_path)

The `.npy` dictionaries and the `convergence.txt` rows have the reference's format, so they
can be diffed against `numerical_experiments/data/*` and `figures/convergence_table_data.tex`.
`exact_pcg` (extended precision, figure_gen.py:53-56) is not part of the GPU path: pass the
reference's function in `variants` if it is wanted; it is then called exactly as there.
"""
from __future__ import annotations

import glob
import os

import numpy as np

from . import callbacks as _cb
from . import cg_variants as _cg

# figure_gen.py:347-348 (method list) and :360 (table columns)
ALL_METHODS = [_cg.hs_pcg, _cg.cg_pcg, _cg.m_pcg, _cg.pr_pcg, _cg.gv_pcg, _cg.pipe_pr_m_pcg, _cg.pipe_pr_pcg,
               _cg.pipe_p_pcg, _cg.pipe_p_m_pcg]
TABLE_METHODS = ["hs_pcg", "cg_pcg", "m_pcg", "pr_pcg", "gv_pcg", "pipe_pr_m_pcg", "pipe_pr_pcg"]


def setup_problem(A):
    """figure_gen.py:31-34: x_true = ones/sqrt(N), b = A x_true, x0 = 0."""
    N = A.get_shape()[0] if hasattr(A, "get_shape") else A.shape[0]
    x_true = np.ones(N) / np.sqrt(N)
    b = A @ x_true
    x0 = np.zeros(N)
    return x_true, b, x0


def test_matrix(A, max_iter, title, preconditioner=None, variants=(), data_dir="./data", save=True, **solver_kwargs):
    """figure_gen.py:21-60.  Returns {method name: output dict}."""
    N = A.get_shape()[0] if hasattr(A, "get_shape") else A.shape[0]
    x_true, b, x0 = setup_problem(A)
    callbacks = [_cb.error_A_norm, _cb.residual_2_norm, _cb.error_2_norm, _cb.updated_residual_2_norm]
    prec = lambda x: x                                            # noqa: E731
    prec_long = lambda x: x                                       # noqa: E731
    if preconditioner == "jacobi":
        dinv = 1 / A.diagonal()                                   # figure_gen.py:43 (hoisted: same products)
        prec = lambda x: dinv * x                                 # noqa: E731
        prec_long = lambda x: (1 / A.diagonal().astype(np.longdouble)) * x    # noqa: E731
    out_dir = os.path.join(data_dir, f"{title}_{preconditioner}")
    if save:
        os.makedirs(out_dir, exist_ok=True)
    trials = {}
    for method in variants:
        if method.__name__ == "exact_pcg":                        # the reference's own function, its own call
            trial = method(A.astype(np.longdouble), b.astype(np.longdouble), x0.astype(np.longdouble),
                           min(max_iter, N), callbacks=callbacks, x_true=x_true.astype(np.longdouble),
                           preconditioner=prec_long)
        else:
            trial = method(A, b, x0, max_iter, callbacks=callbacks, x_true=x_true, preconditioner=prec,
                           **solver_kwargs)
        trials[method.__name__] = trial
        if save:
            np.save(os.path.join(out_dir, method.__name__), trial, allow_pickle=True)
    return trials


def convergence_metrics(error_A_norm, error_tol=1e-5):
    """figure_gen.py:80-89: (first k with relative A-norm error <= tol, 0 = never;
    log10 of the smallest relative A-norm error)."""
    rel = np.asarray(error_A_norm) / error_A_norm[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        return int(np.argmin(rel > error_tol)), float(np.log10(np.nanmin(rel)))


def parse_convergence_data(matrix_name, preconditioner=None, variants=TABLE_METHODS, A=None, n=None, nnz=None,
                           data_dir="./data", trials=None, write=True):
    """figure_gen.py:62-124: one LaTeX table row (also written to <dir>/convergence.txt)."""
    if A is not None:
        n, nnz = A.shape[0], A.nnz
    out_dir = os.path.join(data_dir, f"{matrix_name}_{preconditioner}")
    min_iters, min_errors = [], []
    for method in variants:
        name = method if isinstance(method, str) else method.__name__
        trial = trials[name] if trials is not None else \
            np.load(os.path.join(out_dir, name + ".npy"), allow_pickle=True).item()
        it, acc = convergence_metrics(trial["error_A_norm"])
        min_iters.append(it)
        min_errors.append(acc)
    formatted_matrix_name = r"\texttt{" + matrix_name.replace("_", r"\_") + r"}"
    formatted_preconditioner = "Jac." if preconditioner == "jacobi" else "-"
    data = f"{formatted_matrix_name} & {formatted_preconditioner} & {n} & {nnz}"
    data_iter = data_err = ""
    for k in range(len(min_errors)):
        formatted_min_iter = min_iters[k] if min_iters[k] != 0 else "-"
        mi_bold = "\\tableemph" if ((min_iters[k] > 1.1 * min_iters[0]) or (min_iters[k] == 0)) else ""
        me_bold = "\\tableemph" if (min_errors[k] > .9 * min_errors[0]) else ""
        data_iter += f"& {mi_bold}{{{formatted_min_iter}}}"
        data_err += f"&{me_bold}{{{min_errors[k]:1.2f}}}"
    data += data_iter + data_err + "\\\\ \n"
    if write:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "convergence.txt"), "w") as fh:
            fh.write(data)
    return data, min_iters, min_errors


def gen_convergence_table(data_dir="./data", out_path="./figures/convergence_table_data.tex"):
    """figure_gen.py:117-124: concatenate the rows, un-preconditioned cases first."""
    rows = []
    for suffix in ("None", "jacobi"):
        for path in sorted(glob.glob(os.path.join(data_dir, f"*{suffix}", "convergence.txt"))):
            rows.append(open(path).read())
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    with open(out_path, "w") as fh:
        fh.write("".join(rows))
    return rows
