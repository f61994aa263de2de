#!/usr/bin/env python3
"""Persistent-kernel CTA shape sweep: us/iteration for (threads per CTA, CTA cap).
    python tools/pers_tune.py [--grid 64] [--iters 1000] [--one T,CTAS,variant]   (--one: a single run, for ncu)"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from new_cg_variants_b200 import PoissonStencil, Session   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=64)
ap.add_argument("--iters", type=int, default=1000)
ap.add_argument("--one", default="")
args = ap.parse_args()
S = PoissonStencil(args.grid, args.grid, args.grid, dim=3)
n = S.shape[0]
x_true = np.ones(n) / np.sqrt(n)
b, x0 = S @ x_true, np.zeros(n)
out = {}
with Session(S, dinv=1 / S.diagonal()) as s:
    s.load_problem(b, x0, None)
    if args.one:
        T, C, v = args.one.split(",")
        s.set_option("pers_threads", int(T)); s.set_option("pers_ctas", int(C))
        info = s.run(v, args.iters + 1, path="persistent")
        print(json.dumps({"us_per_iteration": 1e3 * info["loop_ms"] / args.iters}))
        sys.exit(0)
    for T in (128, 256, 512):
        for C in (64, 128, 148, 256, 296, 444, 592):
            s.set_option("pers_threads", T); s.set_option("pers_ctas", C)
            row = {}
            for v in ("hs", "cg", "pr", "gv", "pipe_pr"):
                try:
                    best = min(s.run(v, args.iters + 1, path="persistent")["loop_ms"] for _ in range(3))
                    row[v] = round(1e3 * best / args.iters, 2)
                except Exception as e:
                    row[v] = str(e)[-60:]
            out[f"T{T}/C{C}"] = row
            print(f"T{T}/C{C}", row, file=sys.stderr, flush=True)
print(json.dumps(out))
