#!/usr/bin/env python3
"""CSR roofline datapoint: the PETSc driver's banded model problem (scaling_experiments_petsc/
ex2b.c:86-97; strong_scaling_tests.py:49-56): n = 650 000, half-bandwidth k = 32 (65 nnz/row),
off-diagonals 1e-4, diagonal 1 + (i/(n-1)) (kappa-1) rho^(n-1-i), kappa = 1e6, rho = 0.95,
x* = ones, no preconditioner (-pc_type none, strong_scaling_tests.py:44).  Fixed iteration count, CSR-stream SpMV kernels (stream path).

    python tools/csr_bench.py [--n 650000] [--k 32] [--iters 500]

Prints one JSON object: us/iteration per variant, the fused CSR SpMV kernel's achieved GB/s on
its algorithmic bytes (12 nnz + 4 (n+1) + 8 n words), final error ||x - 1||_2 / sqrt(n).
"""
import argparse, json, os, sys, time
import numpy as np
import scipy.sparse as sps
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from new_cg_variants_b200 import Session          # noqa: E402

SP_WORDS = {"sp_hs": 2, "sp_cg": 3, "sp_gv": 2, "sp_pr": 3, "sp_pipe_r": 4}    # vector words per row


def model_matrix(n, k, kappa=1e6, rho=0.95, off=1e-4):
    i = np.arange(n, dtype=np.float64)
    diag = 1.0 + (i / (n - 1)) * (kappa - 1) * rho ** (n - 1 - i)
    offs = list(range(-k, 0)) + list(range(1, k + 1))
    A = sps.diags([np.full(n - abs(o), off) for o in offs], offs, shape=(n, n), format="csr") + sps.diags(diag)
    A = sps.csr_matrix(A)
    A.sort_indices()
    return A


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=650000)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--variants", default="hs,cg,pr,gv,pipe_pr")
    ap.add_argument("--sweep", default="", help="';'-separated option sets, each 'name=value,name=value' (cgx_set_option); "
                    "one JSON object per set, e.g. 'csr_bulk=0;csr_bulk=1,csr_bulk_ctas=3'")
    args = ap.parse_args()
    t0 = time.time()
    A = model_matrix(args.n, args.k)
    n, nnz = A.shape[0], A.nnz
    x_true = np.ones(n)
    b, x0 = A @ x_true, np.zeros(n)
    workload = f"banded model problem n={n} k={args.k} nnz={nnz} unpreconditioned, {args.iters} iterations"
    sets = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in st.split(",") if kv) for st in args.sweep.split(";")]
    with Session(A) as s:
        s.load_problem(b, x0, None)
        for opts in sets:
            for name, value in opts.items():
                s.set_option(name, value)
            out = {"workload": workload, "options": opts, "build_s": round(time.time() - t0, 1), "variants": {}}
            for v in args.variants.split(","):
                best = min(s.run(v, args.iters + 1, histories=(), path="stream")["loop_ms"] for _ in range(3))
                x, _ = s.fetch(want_hist=False)
                s.set_profile(True)
                s.run(v, args.iters + 1, histories=(), path="stream")
                prof = s.get_profile()
                s.set_profile(False)
                row = {"us_per_iteration": round(1e3 * best / args.iters, 2),
                       "rel_error": float(np.linalg.norm(x - x_true) / np.sqrt(n)), "kernels": {}}
                for kname, (ms, cnt) in prof.items():
                    us = 1e3 * ms / cnt
                    row["kernels"][kname] = {"us": round(us, 2)}
                    if kname in SP_WORDS:
                        bytes_ = 12.0 * nnz + 4.0 * (n + 1) + 8.0 * n * SP_WORDS[kname]
                        row["kernels"][kname]["GBps"] = round(bytes_ / (us * 1e-6) / 1e9)
                        row["kernels"][kname]["algorithmic_bytes"] = bytes_
                out["variants"][v] = row
                print(opts, v, row, file=sys.stderr, flush=True)
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
