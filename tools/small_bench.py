#!/usr/bin/env python3
"""Latency-bound cases (BASELINE.json configs[1] and the single-GPU end of configs[4]):
microseconds per iteration of every variant on SuiteSparse fixtures and Poisson 64^3,
stream path (2-3 launches per iteration) vs persistent path (one cooperative launch).

    python tools/small_bench.py [--iters 2000]

Prints one JSON object; instrumentation off (callbacks=[] protocol), CUDA-event loop time,
best of 3.  Also runs the longest case of figure_gen.py (bcsstk18, Jacobi: 2700 iterations;
un-preconditioned: 1 750 000 iterations) with the four histories on, as a wall-clock number.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import helpers
    from helpers import orc
    from new_cg_variants_b200 import PoissonStencil, Session

    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--long", action="store_true", help="also the 1.75 M-iteration bcsstk18 case")
    args = ap.parse_args()
    cases = []
    for name in ("bcsstk03", "nos4", "bcsstk16", "bcsstk18"):
        A = helpers.load_matrix(name)
        cases.append((name, A, A))
    for g in (32, 64):
        S = PoissonStencil(g, g, g, dim=3)
        cases.append((f"poisson3d_{g}", S, None))
    out = {"unit": "us/iteration", "iters": args.iters, "cases": {}}
    for name, op, A in cases:
        n = op.shape[0]
        x_true = np.ones(n) / np.sqrt(n)
        b, x0 = op @ x_true, np.zeros(n)
        dinv = 1 / op.diagonal()
        row = {"n": n, "nnz": int(op.nnz)}
        with Session(op, dinv=dinv) as s:
            s.load_problem(b, x0, None)
            for path in ("stream", "persistent"):
                for v in ("hs", "cg", "gv", "pr", "pipe_pr"):
                    best = None
                    for _ in range(4):
                        info = s.run(v, args.iters + 1, histories=(), path=path)
                        best = info["loop_ms"] if best is None else min(best, info["loop_ms"])
                    row[f"{path}/{v}"] = round(1e3 * best / args.iters, 3)
        out["cases"][name] = row
        print(name, row, file=sys.stderr, flush=True)
    # figure_gen.py:270-272 -- bcsstk18
    A = helpers.load_matrix("bcsstk18")
    x_true, b, x0 = orc.setup_problem(A)
    longs = [("bcsstk18_jacobi", 2700, orc.jacobi_dinv(A))]
    if args.long:
        longs.append(("bcsstk18_None", 1750000, None))
    for name, max_iter, dinv in longs:
        with Session(A, dinv=dinv) as s:
            t0 = time.perf_counter()
            x, hist, info = s.solve("pipe_pr", b, x0, max_iter, x_true=x_true, path="persistent")
            dt = time.perf_counter() - t0
            it, acc = orc.convergence_metrics(hist["error_A_norm"])
            out["cases"][name] = {"max_iter": max_iter, "wall_s": dt, "loop_ms": info["loop_ms"],
                                  "us_per_iteration_with_histories": 1e3 * info["loop_ms"] / (max_iter - 1),
                                  "iters_to_1e-5": it, "log10_accuracy": acc, "launches": info["kernel_launches"]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
