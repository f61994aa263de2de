#!/usr/bin/env python3
"""Small single-GPU workload for ncu: `iters` iterations of one variant on a Poisson grid.
    python tools/prof_target.py [variant] [grid] [dim] [iters] [option=value ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from new_cg_variants_b200 import PoissonStencil, Session  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "pr"
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 3
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 12
S = PoissonStencil(grid, grid, grid, dim=3) if dim == 3 else PoissonStencil(grid, grid, 1, dim=2)
n = S.shape[0]
x_true = np.ones(n) / np.sqrt(n)
b = S @ x_true
with Session(S, dinv=1 / S.diagonal()) as s:
    for opt in sys.argv[5:]:
        k, v = opt.split("=")
        s.set_option(k, int(v))
    s.load_problem(b, np.zeros(n), None)
    info = s.run(variant, iters + 1, histories=(), path="stream")
    info = s.run(variant, iters + 1, histories=(), path="stream")
    print(variant, grid, dim, "loop ms/iter", info["loop_ms"] / iters, "launches", info["kernel_launches"])
