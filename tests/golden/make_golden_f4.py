#!/usr/bin/env python3
"""Golden outputs of the reference for SURVEY.md section 8 row f4 (build container only):
GV residual replacement (gv_cg.py:156-158) with a periodic and a vector-reading predicate, and the
callbacks save_x / save_r / lanczos_recurrence / updated_error_A_norm, on nos4 + Jacobi.
Asserts that oracle/cg_oracle.py reproduces the replacement runs bit for bit.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_f4.py   -> tests/golden/f4.npz
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/predict_and_recompute"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "numerical_experiments"))
warnings.filterwarnings("ignore")

import cg_variants as ref_solvers            # noqa: E402
import callbacks as ref_callbacks            # noqa: E402
from oracle import cg_oracle as orc          # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers                               # noqa: E402

PERIODIC = lambda **kw: kw["k"] % 7 == 0                                                  # noqa: E731
# replace when the updated and the implied residual of w drift apart (reads vectors)
DRIFT = lambda **kw: np.linalg.norm(kw["w"] - kw["A"] @ kw["r"]) > 1e-9 * np.linalg.norm(kw["w"])   # noqa: E731


def main():
    A = helpers.load_matrix("nos4")
    x_true, b, x0 = orc.setup_problem(A)
    d = 1 / A.diagonal()
    prec = lambda v: d * v                                                                # noqa: E731
    std = [ref_callbacks.error_A_norm, ref_callbacks.residual_2_norm, ref_callbacks.error_2_norm,
           ref_callbacks.updated_residual_2_norm]
    out = {}
    for name, pred in (("periodic7", PERIODIC), ("drift", DRIFT)):
        ref = ref_solvers.gv_pcg(A, b, x0, 80, w_replace=pred, preconditioner=prec, callbacks=std, x_true=x_true)
        o = orc.solve("gv", A, b, x0, 80, dinv=d, x_true=x_true, w_replace=pred)
        for h in orc.HISTORIES:
            assert np.array_equal(o[h], np.asarray(ref[h], dtype=np.float64)), (name, h)
            out[f"gv_{name}/{h}"] = np.asarray(ref[h], dtype=np.float64)
    extra = [ref_callbacks.save_x, ref_callbacks.save_r, ref_callbacks.lanczos_recurrence,
             ref_callbacks.updated_error_A_norm]
    for fn in ("hs_pcg", "pr_pcg", "pipe_pr_pcg"):
        ref = getattr(ref_solvers, fn)(A, b, x0, 40, preconditioner=prec, callbacks=std + extra, x_true=x_true)
        for key in ("x", "r", "lanczos_alpha", "lanczos_beta", "lanczos_z", "lanczos_3_term_error",
                    "lanczos_orthogonality", "updated_error_A_norm"):
            out[f"{fn}/{key}"] = np.asarray(ref[key], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "f4.npz"), **out)
    print("wrote f4.npz:", len(out), "arrays; oracle == reference on both replacement runs")


if __name__ == "__main__":
    main()
