#!/usr/bin/env python3
"""Where csr_bulk_kernel's CTA 0 spends its cycles (option debug_skip & 2): per role, total cycles
of the role's loop and the cycles it spent blocked on an mbarrier.  Banded model problem."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from new_cg_variants_b200 import Session, _lib          # noqa: E402
from tools.csr_bench import model_matrix                 # noqa: E402

lib = _lib.load()
A = model_matrix(650000, 32)
n = A.shape[0]
b, x0 = A @ np.ones(n), np.zeros(n)
with Session(A) as s:
    s.load_problem(b, x0, None)
    for opts in sys.argv[1:] or ["debug_skip=2"]:
        for kv in opts.split(","):
            s.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        for v in ("pr", "pipe_pr"):
            s.run(v, 21, histories=(), path="stream")
            out = (C.c_uint64 * 16)()
            _lib.check(lib.cgx_debug_times(s._ctx, out))
            t = [int(x) for x in out]
            print(opts, v, "producer total/blocked %d/%d  gather %d/%d  summing %d/%d (SM cycles, CTA 0, last SpMV pass)"
                  % (t[10], t[11], t[12], t[13], t[14], t[15]), flush=True)
