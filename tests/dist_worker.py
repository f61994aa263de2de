"""torchrun worker: one rank per GPU runs the partitioned solve (DistSession, CUDA IPC
windows) and rank 0 compares with the single-GPU emulation of the same partition
(GroupSession) -- they must agree bit for bit -- for both scalar-exchange modes."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from new_cg_variants_b200 import PoissonStencil
    from new_cg_variants_b200.dist import DistSession, GroupSession
    from new_cg_variants_b200 import _lib

    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--shape", default="64,20,24")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")          # plumbing only: window handles and barriers
    nx, ny, nz = map(int, args.shape.split(","))
    S = PoissonStencil(nx, ny, nz, dim=3)
    n = S.shape[0]
    x_true = np.ones(n) / np.sqrt(n)
    b, x0 = S @ x_true, np.zeros(n)
    dinv = 1 / S.diagonal()
    hist = _lib.HIST_NAMES
    ok = True
    for mode in ("p2p", "nccl"):
        sess = DistSession(S, dinv=dinv, device=local, mode=mode)
        results = {}
        paths = ("stream", "persistent") if mode == "p2p" else ("stream",)
        for path in paths:
            for tag in ("hs", "cg", "gv", "pr", "m", "pipe_pr", "pipe_p"):
                x_loc, h, info = sess.solve(tag, b, x0, 25, x_true=x_true, histories=hist, path=path)
                x = sess.gather_x(x_loc)
                results[(tag, path)] = (x, h)
        sess.close()
        if rank == 0:
            grp = GroupSession(S, world, dinv=dinv, devices=[local] * world)
            for (tag, path), (x, h) in results.items():
                xg, hg, _ = grp.solve(tag, b, x0, 25, x_true=x_true, path=path)
                if mode == "p2p":         # same rank-ordered sums: same bits
                    same = np.array_equal(x, xg) and all(np.array_equal(h[k], hg[k], equal_nan=True) for k in hist)
                else:                     # NCCL picks its own summation tree: rounding-level agreement
                    same = np.allclose(x, xg, rtol=1e-9, atol=1e-13) and \
                        all(np.allclose(h[k][:10], hg[k][:10], rtol=1e-10) for k in hist)
                print(f"[{mode}/{path}] {tag}: {'match' if same else 'MISMATCH'}", flush=True)
                ok = ok and same
            grp.close()
        dist.barrier()
    # general CSR row partition: multi-process run == single-GPU emulation, bit for bit
    import json
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import helpers
    A = helpers.load_matrix("bcsstk16")
    xt, bA, x0A = helpers.orc.setup_problem(A)
    dA = helpers.orc.jacobi_dinv(A)
    sess = DistSession(A, dinv=dA, device=local)
    res = {}
    for tag in ("hs", "cg", "gv", "pr", "pipe_pr"):
        x_loc, h, info = sess.solve(tag, bA, x0A, 20, x_true=xt, histories=hist, path="stream")
        res[tag] = (sess.gather_x(x_loc), h)
    sess.close()
    if rank == 0:
        grp = GroupSession(A, world, dinv=dA, devices=[local] * world)
        for tag, (x, h) in res.items():
            xg, hg, _ = grp.solve(tag, bA, x0A, 20, x_true=xt)
            same = np.array_equal(x, xg) and all(np.array_equal(h[k], hg[k], equal_nan=True) for k in hist)
            print(f"[csr partition] {tag}: {'match' if same else 'MISMATCH'}", flush=True)
            ok = ok and same
        grp.close()
    dist.barrier()
    # the reference's scaling_tests.py protocol on its own model problem (dense column blocks), P ranks:
    # final errors within x2 of the reference's 1-rank run (tests/golden/mpi_kat.json)
    from new_cg_variants_b200 import scaling_tests as st
    kat = json.load(open(os.path.join(helpers.GOLDEN, "mpi_kat.json")))
    out = st.run(kat["pr"]["n"], kat["pr"]["max_iter"], "worker", save=False, verbose=False)
    if rank == 0:
        for tag, fn in (("hs", "hs_cg"), ("cg", "cg_cg"), ("gv", "gv_cg"), ("pr", "pr_cg"), ("pipe_pr", "pipe_pr_cg")):
            ratio = out[fn]["error"] / kat[tag]["error"]
            good = 0.5 <= ratio <= 2.0 and out[fn]["timings"]["tot"] > 0
            print(f"[scaling_tests] {fn}: error {out[fn]['error']:.3e} (reference {kat[tag]['error']:.3e}) {'ok' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
    dist.barrier()
    # the reference's distributed signature f(comm, A, b, max_iter) -> (x_local, times)
    from new_cg_variants_b200 import Session, cg_variants_mpi4py as m
    comm = m.GpuComm()
    r0, r1 = rank * (n // world), (rank + 1) * (n // world)
    xs = {}
    for fn in (m.hs_cg, m.pr_cg, m.pipe_pr_cg):
        x_loc, times = fn(comm, S, b[r0:r1], 20)
        parts = [None] * world
        dist.all_gather_object(parts, x_loc)
        xs[fn.__name__] = np.concatenate(parts)
        assert (times is not None) == (rank == 0)
    m.clear_sessions()
    if rank == 0:
        with Session(S) as one:
            for name, tag in (("hs_cg", "hs"), ("pr_cg", "pr"), ("pipe_pr_cg", "pipe_pr")):
                x1, _, _ = one.solve(tag, b, x0, 21, histories=())
                same = np.allclose(xs[name], x1, rtol=1e-9, atol=1e-13)
                print(f"[mpi4py-shaped] {name}: {'match' if same else 'MISMATCH'}", flush=True)
                ok = ok and same
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0 and flag.item():
        print("dist_worker ok", flush=True)
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
