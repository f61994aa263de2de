#!/usr/bin/env python3
"""Allreduce-hiding experiment (BASELINE.json configs[4]): strong-scaling
tail of the 3-D Poisson problem, fixed iteration count, one rank per GPU under torchrun.

    torchrun --nproc-per-node N tools/overlap_bench.py [--grid 64] [--iters 2000]

For every variant and scalar-exchange mode it times the iteration loop (CUDA events, max
over ranks, best of `--reps`) with the exchange live and with it stubbed to a local stand-in
(cgx_set_option "stub_allreduce"):

    exposed   = t_iter(live) - t_iter(stub)
    hidden(v) = 1 - exposed(v) / exposed(hs)        (also vs pr)

Rank 0 prints one JSON object.  The numbers of a stubbed run are numerically meaningless.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from new_cg_variants_b200 import PoissonStencil
    from new_cg_variants_b200.dist import DistSession

    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=64)
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--modes", default="p2p,nccl")
    ap.add_argument("--variants", default="hs,cg,pr,gv,pipe_pr")
    ap.add_argument("--path", default="stream", choices=["stream", "persistent"])
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S = PoissonStencil(args.grid, args.grid, args.grid, dim=3)
    n = S.shape[0]
    x_true = np.ones(n) / np.sqrt(n)
    b, x0 = S @ x_true, np.zeros(n)
    dinv = 1 / S.diagonal()
    res = {}
    if args.path == "persistent":
        args.modes = "p2p"          # the persistent kernel exchanges scalars in-kernel only
    for mode in args.modes.split(","):
        sess = DistSession(S, dinv=dinv, device=local, mode=mode)
        sess.load_problem(b, x0, None)
        for v in args.variants.split(","):
            for stub in (0, 1):
                sess.set_option("stub_allreduce", stub)
                best = None
                for rep in range(args.reps + 1):
                    dist.barrier()
                    info = sess.run(v, args.iters + 1, path=args.path)
                    t = torch.tensor([info["loop_ms"]], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    if rep > 0:
                        best = t.item() if best is None else min(best, t.item())
                res[(mode, v, stub)] = 1e3 * best / args.iters          # us / iteration
            sess.set_option("stub_allreduce", 0)
        sess.close()
        dist.barrier()
    if rank == 0:
        out = {"workload": f"poisson3d_{args.grid} jacobi, {args.iters} iterations, {world} GPUs, {args.path} path",
               "unit": "us/iteration", "modes": {}}
        for mode in args.modes.split(","):
            m = {}
            for v in args.variants.split(","):
                live, stub = res[(mode, v, 0)], res[(mode, v, 1)]
                m[v] = {"live": live, "stub": stub, "exposed": live - stub}
            for v in m:
                for base in ("hs", "pr"):
                    if base in m and m[base]["exposed"] > 0:
                        m[v][f"hidden_vs_{base}"] = 1 - m[v]["exposed"] / m[base]["exposed"]
            out["modes"][mode] = m
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
