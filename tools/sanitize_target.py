#!/usr/bin/env python3
"""Small workload touching every kernel family, for `compute-sanitizer --tool memcheck|racecheck`:
all nine variants on a SuiteSparse fixture (CSR, stream + persistent incl. the shared-memory slab),
on a small Poisson grid (TMA stencil + generic stencil + elided CG/GV), and on a 3-slab partition
emulated on one GPU (stream protocol and the persistent kernel's in-kernel exchange)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers                                                       # noqa: E402
from new_cg_variants_b200 import PoissonStencil, Session            # noqa: E402
from new_cg_variants_b200.dist import GroupSession                  # noqa: E402
from new_cg_variants_b200 import _lib                               # noqa: E402

TAGS = list(_lib.VARIANT_IDS)
A = helpers.load_matrix("bcsstk03")
n = A.shape[0]
xt = np.ones(n) / np.sqrt(n)
b, x0 = A @ xt, np.zeros(n)
with Session(A, dinv=1 / A.diagonal()) as s:
    for path in ("stream", "persistent"):
        for t in TAGS:
            s.solve(t, b, x0, 8, x_true=xt, path=path)
for shape in ((34, 10, 6), (7, 5, 4)):
    S = PoissonStencil(*shape, dim=3)
    n = S.shape[0]
    xt = np.ones(n) / np.sqrt(n)
    b, x0 = S @ xt, np.zeros(n)
    for dinv in (1 / S.diagonal(), 1 / (S.diagonal() + np.arange(n) % 3)):
        with Session(S, dinv=dinv) as s:
            for path in ("stream", "persistent"):
                for t in TAGS:
                    s.solve(t, b, x0, 6, x_true=xt, path=path)
S = PoissonStencil(34, 10, 9, dim=3)
n = S.shape[0]
xt = np.ones(n) / np.sqrt(n)
b, x0 = S @ xt, np.zeros(n)
grp = GroupSession(S, 3, dinv=1 / S.diagonal())
for path in ("stream", "persistent"):
    for t in TAGS:
        grp.solve(t, b, x0, 6, x_true=xt, path=path)
grp.close()
print("sanitize_target ok")
