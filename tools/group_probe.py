#!/usr/bin/env python3
"""Per-kernel device time of the partitioned kernels with NO inter-GPU effects: two ranks
emulated on one GPU (shared stream, producers always complete before consumers) vs the plain
single-GPU kernels on a problem of the per-rank size."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from new_cg_variants_b200 import PoissonStencil, Session          # noqa: E402
from new_cg_variants_b200.dist import GroupSession                 # noqa: E402

nx = ny = 256
ppr, world, iters = 32, 2, 100
out = {}
S = PoissonStencil(nx, ny, ppr * world, dim=3)
n = S.shape[0]
b, x0 = S @ (np.ones(n) / np.sqrt(n)), np.zeros(n)
grp = GroupSession(S, world, dinv=1 / S.diagonal())
grp.load_problem(b, x0, None)
for v in ("pr", "pipe_pr", "hs"):
    for _ in range(2):
        grp.begin(v, iters + 1, ()); grp.advance(iters)
    ms = grp.members[0].get_info()["loop_ms"]
    for m in grp.members:
        m.set_profile(True)
    grp.begin(v, iters + 1, ()); grp.advance(iters)
    prof = grp.members[0].get_profile()
    for m in grp.members:
        m.set_profile(False)
    out[f"group2/{v}"] = {"us_per_iteration_both_ranks": 1e3 * ms / iters,
                          "kernels_us_rank0": {k: round(1e3 * x[0] / x[1], 2) for k, x in prof.items()}}
grp.close()
S1 = PoissonStencil(nx, ny, ppr, dim=3)
n1 = S1.shape[0]
with Session(S1, dinv=1 / S1.diagonal()) as one:
    one.load_problem(S1 @ (np.ones(n1) / np.sqrt(n1)), np.zeros(n1), None)
    for v in ("pr", "pipe_pr", "hs"):
        best = min(one.run(v, iters + 1, histories=(), path="stream")["loop_ms"] for _ in range(3))
        one.set_profile(True)
        one.run(v, iters + 1, histories=(), path="stream")
        prof = one.get_profile()
        one.set_profile(False)
        out[f"single/{v}"] = {"us_per_iteration": 1e3 * best / iters,
                              "kernels_us": {k: round(1e3 * x[0] / x[1], 2) for k, x in prof.items()}}
for k, v in out.items():
    print(k, v)
