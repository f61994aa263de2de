#!/usr/bin/env python3
"""Small single-GPU CSR workload for ncu: the banded model problem, `iters` iterations of one variant.
    python tools/prof_csr_target.py [variant] [iters] [n]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from new_cg_variants_b200 import Session  # noqa: E402
from new_cg_variants_b200.experiments import banded_model_problem  # noqa: E402
variant = sys.argv[1] if len(sys.argv) > 1 else "pr"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
n = int(sys.argv[3]) if len(sys.argv) > 3 else 650000
A, b, x_true = banded_model_problem(n)
with Session(A) as s:
    s.load_problem(b, np.zeros(n), None)
    info = s.run(variant, iters + 1, histories=(), path="stream")
    info = s.run(variant, iters + 1, histories=(), path="stream")
    print(variant, "us/iter", 1e3 * info["loop_ms"] / iters)
