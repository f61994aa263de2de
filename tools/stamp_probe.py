#!/usr/bin/env python3
"""Time stamps inside the vector pass (CTA 0): start -> scalars folded -> CTA done; two ranks
emulated on one GPU vs a single context at the per-rank size."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from new_cg_variants_b200 import PoissonStencil, Session, _lib          # noqa: E402
from new_cg_variants_b200.dist import GroupSession                        # noqa: E402

lib = _lib.load()
def stamps(ctx):
    out = (C.c_uint64 * 16)()
    _lib.check(lib.cgx_debug_times(ctx, out))
    return [int(v) for v in out[:10]]

S = PoissonStencil(256, 256, 64, dim=3); n = S.shape[0]
b, x0 = S @ (np.ones(n) / np.sqrt(n)), np.zeros(n)
grp = GroupSession(S, 2, dinv=1 / S.diagonal())
grp.load_problem(b, x0, None)
for m in grp.members: m.set_option("debug_skip", 2)
for v in ("pr", "pipe_pr", "hs"):
    for it in (20, 21, 22):
        grp.begin(v, it + 1, ()); grp.advance(it)
        t = stamps(grp.members[0]._ctx)
        print("group2", v, "fold_us", (t[1] - t[0]) / 1965., "cta0_total_us", (t[2] - t[0]) / 1965.,
              "detail(SM cycles since start): scal_loaded %d rec0_fetched %d rec0_applied %d rec1_fetched %d rec1_applied %d stored %d"
              % tuple(x - t[0] for x in (t[4], t[5], t[6], t[7], t[8], t[9])))
grp.close()
S1 = PoissonStencil(256, 256, 32, dim=3); n1 = S1.shape[0]
with Session(S1, dinv=1 / S1.diagonal()) as one:
    one.set_option("debug_skip", 2)
    one.load_problem(S1 @ (np.ones(n1) / np.sqrt(n1)), np.zeros(n1), None)
    for v in ("pr", "pipe_pr", "hs"):
        one.run(v, 21, histories=(), path="stream")
        t = stamps(one._ctx)
        print("single", v, "fold_us", (t[1] - t[0]) / 1965., "cta0_total_us", (t[2] - t[0]) / 1965.)
