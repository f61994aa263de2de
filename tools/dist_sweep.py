#!/usr/bin/env python3
"""Partitioned PR-CG: us/iteration for a few library options (chunk counts of the fused kernel,
two-kernel path, stubbed exchange).   torchrun --nproc-per-node N tools/dist_sweep.py [--grid 256]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from new_cg_variants_b200 import PoissonStencil
    from new_cg_variants_b200.dist import DistSession
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--variant", default="pr")
    ap.add_argument("--sets", default="default;fused_chunks=1;fused_chunks=2;fused_chunks=3;fused_chunks=4;fused_chunks=8;pr_fused=0")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    S = PoissonStencil(args.grid, args.grid, args.grid, dim=3)
    n = S.shape[0]
    b, x0 = S @ (np.ones(n) / np.sqrt(n)), np.zeros(n)
    out = {}
    for spec in args.sets.split(";"):
        sess = DistSession(S, dinv=1 / S.diagonal(), device=local)
        sess.load_problem(b, x0, None)
        if spec != "default":
            for kv in spec.split(","):
                k, v = kv.split("=")
                sess.set_option(k, int(v))
        best = None
        for _ in range(4):
            dist.barrier()
            info = sess.run(args.variant, args.iters + 1, histories=(), path="stream")
            t = torch.tensor([info["loop_ms"]], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = t.item() if best is None else min(best, t.item())
        out[spec] = round(1e3 * best / args.iters, 2)
        sess.close()
    if rank == 0:
        print(json.dumps({"world": world, "grid": args.grid, "variant": args.variant, "us_per_iteration": out}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
