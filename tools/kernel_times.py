#!/usr/bin/env python3
"""Per-kernel-class device time (event pair per launch) of the stream path on a Poisson grid.
    python tools/kernel_times.py [grid] [dim] [variants] [option=value ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from new_cg_variants_b200 import PoissonStencil, Session  # noqa: E402
gs = sys.argv[1] if len(sys.argv) > 1 else "256"
dims = [int(t) for t in gs.split("x")]
grid = dims[0]
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
variants = (sys.argv[3] if len(sys.argv) > 3 else "hs,cg,gv,pr,pipe_pr").split(",")
S = PoissonStencil(*dims, dim=3) if len(dims) == 3 else (PoissonStencil(grid, grid, grid, dim=3) if dim == 3 else PoissonStencil(grid, grid, 1, dim=2))
n = S.shape[0]
b = S @ (np.ones(n) / np.sqrt(n))
with Session(S, dinv=1 / S.diagonal()) as s:
    for opt in sys.argv[4:]:
        k, v = opt.split("=")
        s.set_option(k, int(v))
    s.load_problem(b, np.zeros(n), None)
    for v in variants:
        s.run(v, 61, histories=(), path="stream")
        s.set_profile(True)
        s.run(v, 61, histories=(), path="stream")
        prof = s.get_profile()
        s.set_profile(False)
        t = min(s.run(v, 201, histories=(), path="stream")["loop_ms"] for _ in range(3)) / 200
        print(v, "us/iter", round(1e3 * t, 1), {k: round(1e3 * ms / cnt, 1) for k, (ms, cnt) in prof.items()}, flush=True)
