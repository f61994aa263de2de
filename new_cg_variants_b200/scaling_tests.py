#!/usr/bin/env python3
"""Strong-scaling driver in the protocol of the reference's
scaling_experiments_mpi4py/scaling_tests.py:29-86, running on the GPU path:

    python -m new_cg_variants_b200.scaling_tests <n> <max_iter> <trial_name> [--data-dir ./data]
    python -m torch.distributed.run --nproc-per-node P -m new_cg_variants_b200.scaling_tests <n> <max_iter> <trial>

What the protocol fixes (and this module reproduces):
  * the model problem: eigenvalues  lambda_i = 1/kappa + (1 - 1/kappa) (i/(n-1)) rho^(n-1-i),
    kappa = 1e6, rho = 0.9, built on rank 0 and scattered in blocks of n/P (scaling_tests.py:31-45);
  * the operand every solver receives: this rank's dense (n, n/P) COLUMN block of diag(lambda)
    and b = lambda_block / sqrt(n), so that the solution is ones/sqrt(n) (:47-53);
  * five variants (hs, cg, gv, pr, pipe_pr), each started after a barrier with `max_iter`
    iterations; the local solutions are gathered on rank 0 (:60-72);
  * the record: error = ||ones/sqrt(n) - x||_2, printed as "<variant> error: <value>", and saved as
    <data-dir>/<n>/<variant>_<trial_name>.npy = {"error": ..., "timings": {...}} (:74-86) -- the
    layout of the reference's data/12288/*.npy, so the two sets can be diffed.
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from . import cg_variants_mpi4py as solvers

VARIANTS = ("hs_cg", "cg_cg", "gv_cg", "pr_cg", "pipe_pr_cg")
KAPPA, RHO = 1e6, 0.9


def model_spectrum(n, kappa=KAPPA, rho=RHO):
    i = np.arange(n)
    return 1 / kappa + (1 - 1 / kappa) * i / (n - 1) * rho ** np.arange(n - 1, -1, -1, dtype="float")


def column_block(eigs_local, n, rank):
    """(n, n/P) column block of diag(lambda): zeros except this rank's diagonal block."""
    m = len(eigs_local)
    block = np.zeros((n, m))
    block[rank * m:(rank + 1) * m] += np.diag(eigs_local)
    return block


def run(n, max_iter, trial_name, data_dir="./data", comm=None, save=True, verbose=True):
    comm = comm or solvers.GpuComm()
    ranks, me = comm.Get_size(), comm.Get_rank()
    if n % ranks:
        raise ValueError("n must be a multiple of the number of processes")
    say = print if (verbose and me == 0) else (lambda *a, **k: None)
    comm.Barrier()
    say(f"trial name: {trial_name}")
    say(f"start distributing to {ranks} ranks")
    eigs = np.empty(n // ranks)
    comm.Scatter(model_spectrum(n).reshape(ranks, -1) if me == 0 else None, eigs, root=0)
    A = column_block(eigs, n, me)
    b = eigs / np.sqrt(n)
    comm.Barrier()
    say("done distributing")
    exact = np.ones(n) / np.sqrt(n)
    records = {}
    for name in VARIANTS:
        comm.Barrier()
        x_local, timings = getattr(solvers, name)(comm, A, b, max_iter)
        gathered = np.empty((ranks, n // ranks)) if me == 0 else None
        comm.Gather(x_local, gathered, root=0)
        if me != 0:
            continue
        err = np.linalg.norm(exact - gathered.reshape(n))
        say(f"{name} error: {err}")
        records[name] = {"error": err, "timings": timings}
        if save:
            out_dir = os.path.join(data_dir, str(n))
            os.makedirs(out_dir, exist_ok=True)
            np.save(os.path.join(out_dir, f"{name}_{trial_name}"), records[name], allow_pickle=True)
    solvers.clear_sessions()
    return records


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("n", type=int)
    ap.add_argument("max_iter", type=int)
    ap.add_argument("trial_name")
    ap.add_argument("--data-dir", default="./data")
    args = ap.parse_args()
    multi = int(os.environ.get("WORLD_SIZE", "1")) > 1
    if multi:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("gloo")
    run(args.n, args.max_iter, args.trial_name, args.data_dir)
    if multi:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
