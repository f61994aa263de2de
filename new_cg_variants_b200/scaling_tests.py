#!/usr/bin/env python3
"""The reference's strong-scaling driver protocol (scaling_experiments_mpi4py/scaling_tests.py:29-86)
on the GPU path:

    python -m new_cg_variants_b200.scaling_tests <n> <max_iter> <trial_name> [--data-dir ./data]
    python -m torch.distributed.run --nproc-per-node P -m new_cg_variants_b200.scaling_tests <n> <max_iter> <trial>

Same steps as the reference: rank 0 builds the model spectrum (kappa = 1e6, rho = 0.9), the
eigenvalues are scattered, every rank fills its dense (n, n/P) column block of A with its diagonal
block, b is normalised so that the solution is ones/sqrt(n); the five variants run `max_iter`
iterations each after a barrier; the solution is gathered on rank 0, which prints the error and
saves  <data-dir>/<n>/<variant>_<trial_name>.npy = {"error": ..., "timings": {...}}  -- the
format of the reference's data/12288/*.npy, so the two sets can be diffed.
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from . import cg_variants_mpi4py as m


def run(n, max_iter, trial_name, data_dir="./data", comm=None, save=True, verbose=True):
    comm = comm or m.GpuComm()
    size, rank = comm.Get_size(), comm.Get_rank()
    assert n % size == 0, "n must be a multiple of the number of processes"
    if rank == 0:
        kappa, rho = 1e6, 0.9
        lambda1, lambdan = 1 / kappa, 1
        Lambda = lambda1 + (lambdan - lambda1) * np.arange(n) / (n - 1) * rho ** np.arange(n - 1, -1, -1, dtype="float")
        sendbuf = Lambda.reshape(size, -1)
    else:
        sendbuf = None
    comm.Barrier()
    if rank == 0 and verbose:
        print(f"trial name: {trial_name}\nstart distributing to {size} ranks")
    b = np.empty(n // size, dtype="float")
    comm.Scatter(sendbuf, b, root=0)
    A = np.zeros((n, n // size), dtype="float")
    A[rank * (n // size):(rank + 1) * (n // size)] += np.diag(b)
    b /= np.sqrt(n)
    comm.Barrier()
    if rank == 0 and verbose:
        print("done distributing")
    results = {}
    for variant in (m.hs_cg, m.cg_cg, m.gv_cg, m.pr_cg, m.pipe_pr_cg):
        comm.Barrier()
        sol, t = variant(comm, A, b, max_iter)
        sol_raw = np.empty([size, n // size], dtype="float") if rank == 0 else None
        comm.Gather(sol, sol_raw, root=0)
        if rank == 0:
            sol_raw = np.reshape(sol_raw, (n))
            error = np.linalg.norm(np.ones(n) / np.sqrt(n) - sol_raw)
            if verbose:
                print(f"{variant.__name__} error: {error}")
            res = {"error": error, "timings": t}
            results[variant.__name__] = res
            if save:
                os.makedirs(os.path.join(data_dir, str(n)), exist_ok=True)
                np.save(os.path.join(data_dir, str(n), f"{variant.__name__}_{trial_name}"), res, allow_pickle=True)
    m.clear_sessions()
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("n", type=int)
    ap.add_argument("max_iter", type=int)
    ap.add_argument("trial_name")
    ap.add_argument("--data-dir", default="./data")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("gloo")
    run(args.n, args.max_iter, args.trial_name, args.data_dir)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
