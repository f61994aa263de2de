#!/usr/bin/env python3
"""Where does the multi-GPU stream path lose time?  Per-kernel device time (event pair per
launch) of a slab-partitioned run vs ONE GPU running a problem of the per-rank size.

    torchrun --nproc-per-node 2 tools/dist_probe.py --nx 256 --ny 256 --planes-per-rank 32
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from new_cg_variants_b200 import PoissonStencil, Session
    from new_cg_variants_b200.dist import DistSession
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--ny", type=int, default=256)
    ap.add_argument("--planes-per-rank", type=int, default=32)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--variants", default="pr,pipe_pr,hs")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    out = {}

    def run(sess, tag, label):
        best = None
        for _ in range(3):
            dist.barrier()
            info = sess.run(tag, args.iters + 1, histories=(), path="stream")
            best = info["loop_ms"] if best is None else min(best, info["loop_ms"])
        sess.set_profile(True)
        sess.run(tag, args.iters + 1, histories=(), path="stream")
        prof = sess.get_profile()
        sess.set_profile(False)
        out[label] = {"us_per_iteration": 1e3 * best / args.iters,
                      "kernels_us": {k: round(1e3 * v[0] / v[1], 2) for k, v in prof.items()}}

    S = PoissonStencil(args.nx, args.ny, args.planes_per_rank * world, dim=3)
    n = S.shape[0]
    b, x0 = S @ (np.ones(n) / np.sqrt(n)), np.zeros(n)
    sess = DistSession(S, dinv=1 / S.diagonal(), device=local)
    sess.load_problem(b, x0, None)
    for v in args.variants.split(","):
        run(sess, v, f"dist{world}/{v}")
    # timing experiments (numerically meaningless): scalar exchange local only / no halo traffic
    sess.set_option("stub_allreduce", 1)
    for v in args.variants.split(","):
        run(sess, v, f"dist{world}+stub_scalars/{v}")
    sess.set_option("debug_skip", 1)
    for v in args.variants.split(","):
        run(sess, v, f"dist{world}+stub_scalars+no_halo/{v}")
    sess.set_option("stub_allreduce", 0)
    for v in args.variants.split(","):
        run(sess, v, f"dist{world}+no_halo/{v}")
    sess.set_option("debug_skip", 0)
    sess.close()
    if rank == 0:
        S1 = PoissonStencil(args.nx, args.ny, args.planes_per_rank, dim=3)
        n1 = S1.shape[0]
        with Session(S1, dinv=1 / S1.diagonal(), device=local) as one:
            one.load_problem(S1 @ (np.ones(n1) / np.sqrt(n1)), np.zeros(n1), None)
            for v in args.variants.split(","):
                # no other rank runs now: plain barrier-free timing
                best = min(one.run(v, args.iters + 1, histories=(), path="stream")["loop_ms"] for _ in range(3))
                one.set_profile(True)
                one.run(v, args.iters + 1, histories=(), path="stream")
                prof = one.get_profile()
                one.set_profile(False)
                out[f"single/{v}"] = {"us_per_iteration": 1e3 * best / args.iters,
                                      "kernels_us": {k: round(1e3 * x[0] / x[1], 2) for k, x in prof.items()}}
        print(json.dumps(out, indent=1))
    dist.barrier() if False else None
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
