#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 CG path (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--variant pr] [--grid 256] [--dim 3] [--iters 200]

Workload (BASELINE.json configs[3], the one the metric is quoted on): 3-D Poisson 7-point,
256^3 grid (16.8 M unknowns), Jacobi-preconditioned, x_true = 1/sqrt(n), b = A x_true,
x0 = 0, fp64.  A "step" is one solve of `iters` CG iterations from x0 (initialisation
included).  N GPUs shard the grid into z-slabs (strong scaling, one problem).

  value  : iterations/s with b, x0 resident in HBM, timed with CUDA events on the library's
           launch stream (max over ranks), instrumentation off (the reference's
           callbacks=[] timing protocol)
  e2e    : the same through the C-ABI call with HOST buffers (cgx_solve_host): pinned
           b/x0 copied in, x copied out (into a pinned buffer), inside the timed region
           (wall clock + sync)
  roofline / cpu_baseline : see DESIGN.md
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_V = {"hs": 13, "cg": 14, "pr": 14, "m": 14, "gv": 21, "pipe_pr": 23, "pipe_p": 23,
       "pipe_pr_m": 23, "pipe_p_m": 23}          # fp64 words/row/iteration
# algorithmic fp64 words per row per launch of each kernel class (reads + writes; "+d" = one
# more when a Jacobi vector is read) -- DESIGN.md "Kernels"
CLASS_WORDS = {"ew_hs1": (3, 1), "ew_hs2": (5, 1), "ew_cg": (11, 1), "ew_gv": (19, 1), "ew_pr": (9, 1),
               "ew_pipe_r": (14, 1), "ew_pipe_n": (18, 1), "sp_hs": (2, 0), "sp_cg": (3, 0), "sp_gv": (2, 0),
               "sp_pr": (3, 1), "sp_pipe_r": (4, 0), "sp_pipe_n": (2, 0),
               "pr_fused": (10, 0)}        # one launch per PR-CG iteration: R x,r,p,s,rt  W x,r,p,s,rt
# with a constant Jacobi diagonal (or none) on the TMA stencil path CG-CG does not stream r~ and
# GV does not stream w~ (DESIGN.md "Kernels"): their kernels then move fewer words
CLASS_WORDS_ELIDED = {"ew_cg": (9, 0), "sp_cg": (2, 0), "ew_gv": (17, 0)}
REF_FUN = {"hs": "hs_pcg", "cg": "cg_pcg", "gv": "gv_pcg", "pr": "pr_pcg", "m": "m_pcg",
           "pipe_pr": "pipe_pr_pcg"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); smmax.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smmax) if smmax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(grid, dim=3):
    from new_cg_variants_b200 import PoissonStencil
    S = PoissonStencil(grid, grid, grid, dim=3) if dim == 3 else PoissonStencil(grid, grid, 1, dim=2)
    n = S.shape[0]
    x_true = np.ones(n) / np.sqrt(n)           # figure_gen.py:31-34
    b = S @ x_true
    x0 = np.zeros(n)
    dinv = 1 / S.diagonal()                    # figure_gen.py:43
    return S, b, x0, x_true, dinv


def pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


# ------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's CPU path: its algorithm restated in numpy/scipy (oracle/cg_oracle.py,
    bit-identical to the reference's *_pcg on the same machine) on the same workload.  The
    reference is pure Python and /root/reference does not travel to the GPU box, so this is
    the "port" kind.  Each step is a bounded sample: `ref_iters` iterations from x0."""
    if rank != 0:
        return
    from oracle import cg_oracle as orc
    from threadpoolctl import threadpool_info
    t0 = time.time()
    A = orc.poisson3d(args.grid) if args.dim == 3 else orc.poisson2d(args.grid)
    x_true, b, x0 = orc.setup_problem(A)
    dinv = orc.jacobi_dinv(A)
    build_s = time.time() - t0
    its = args.ref_iters
    for _ in range(min(args.warmup, 1)):       # one warm-up solve
        orc.solve(args.variant, A, b, x0, its + 1, dinv=dinv, history=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.solve(args.variant, A, b, x0, its + 1, dinv=dinv, history=False)
    dt = time.perf_counter() - t0
    value = its * args.steps / dt
    blas = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    n = A.shape[0]
    line = {
        "impl": "reference", "metric": f"CG iterations/s ({REF_FUN.get(args.variant, args.variant)}, Jacobi, {args.dim}-D Poisson {args.grid}^{args.dim})",
        "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"poisson{args.dim}d_{args.grid} {REF_FUN.get(args.variant, args.variant)} jacobi (scipy CSR, nnz={A.nnz})",
                   "iters_per_step": its, "n": n},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": blas, "kind": "port",
                         "sample": f"{args.steps} x {its} iterations of the numpy/scipy restatement on the full {args.grid}^{args.dim} CSR matrix "
                                   f"(scipy SpMV single-threaded, OpenBLAS dots {blas} threads, host has {os.cpu_count()} cpus; "
                                   f"matrix build {build_s:.1f}s untimed)"},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "hbm_gbs_model": value * (8 * n * W_V[args.variant] + 12 * A.nnz + 4 * (n + 1)) / 1e9,
    }
    emit(line)


# ------------------------------------------------------------------------------------ our arm
def cpu_baseline_sample(args):
    from oracle import cg_oracle as orc
    from threadpoolctl import threadpool_info
    A = orc.poisson3d(args.grid) if args.dim == 3 else orc.poisson2d(args.grid)
    x_true, b, x0 = orc.setup_problem(A)
    dinv = orc.jacobi_dinv(A)
    its = args.cpu_iters
    orc.solve(args.variant, A, b, x0, 2, dinv=dinv, history=False)      # warm-up
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        orc.solve(args.variant, A, b, x0, its + 1, dinv=dinv, history=False)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    blas = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    return {"value": its / best, "unit": "iterations/s", "cores": blas, "kind": "port",
            "sample": f"min of 2 runs of {its} iterations of oracle/cg_oracle.py ({REF_FUN.get(args.variant, args.variant)}, callbacks=[], "
                      f"precomputed dinv) on the full {args.grid}^{args.dim} scipy CSR matrix; scipy SpMV single-threaded, "
                      f"OpenBLAS {blas} threads, host {os.cpu_count()} cpus"}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from new_cg_variants_b200 import Session

    torch.cuda.set_device(local_rank)
    if world > 1:
        from new_cg_variants_b200.dist import DistSession      # row-partitioned multi-GPU path
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    S, b, x0, x_true, dinv = build_problem(args.grid, args.dim)
    n = S.shape[0]
    its = args.iters
    variant = args.variant
    if world > 1:
        sess = DistSession(S, dinv=dinv, device=local_rank, rank=rank, world=world)
    else:
        sess = Session(S, dinv=dinv, device=local_rank)
    sess.load_problem(b, x0, None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        info = sess.run(variant, its + 1, histories=(), path=args.path)
        return info["setup_ms"] + info["loop_ms"], info["loop_ms"], info["kernel_launches"]

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms = loop_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        a, l, k = step_resident()
        dev_ms += a; loop_ms += l; launches += k
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([dev_ms, loop_ms, wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, loop_ms, wall_ms = t.tolist()
    value = its * args.steps / (dev_ms / 1e3)

    # ---- e2e through the C-ABI call with host buffers (pinned), copies inside the timing
    e2e = None
    if world == 1:
        tb, hb = pinned(b)
        tx0, hx0 = pinned(x0)
        txo, hxo = pinned(np.empty(n))          # the result lands in pinned host memory too
        for _ in range(2):
            sess.solve(variant, hb, hx0, its + 1, histories=(), path=args.path, x_out=hxo)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            x, _, info = sess.solve(variant, hb, hx0, its + 1, histories=(), path=args.path, x_out=hxo)
        barrier()
        e_dt = time.perf_counter() - t0
        e2e = {"value": its * args.steps / e_dt, "unit": "iterations/s",
               "h2d_bytes_per_step": info["h2d_bytes"], "d2h_bytes_per_step": info["d2h_bytes"],
               "ms_per_step": 1e3 * e_dt / args.steps,
               "api": "cgx_solve_host (Session.solve): pinned host b,x0 in, x out (pinned)"}
    else:
        tb, hb = pinned(sess.local(b))
        tx0, hx0 = pinned(sess.local(x0))
        txo, hxo = pinned(np.empty(sess.n))
        e2e = sess.e2e_bench(variant, hb, hx0, its + 1, args.steps, barrier, x_out=hxo)
        t = torch.tensor([e2e["ms_per_step"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e["ms_per_step"] = t.item()
        e2e["value"] = its / (t.item() / 1e3)

    # ---- per-kernel timing for the roofline (separate, untimed pass with event pairs; on
    #      N > 1 the rows are this rank's slab and the times include waiting for the peers)
    n_loc = n if world == 1 else sess.n
    peak, peak_src = measured_peak()
    sess.set_profile(True)
    sess.run(variant, its + 1, histories=(), path="stream")
    prof = sess.get_profile()
    sess.set_profile(False)
    top = max(prof, key=lambda c: prof[c][0])
    ms, cnt = prof[top]
    # a constant Jacobi diagonal (every Poisson stencil) travels as a scalar: no dinv stream
    dinv_stream = 0 if np.all(dinv == dinv[0]) else 1
    words_tab = dict(CLASS_WORDS)
    if not dinv_stream:
        words_tab.update(CLASS_WORDS_ELIDED)
    cw = lambda c: words_tab[c][0] + words_tab[c][1] * dinv_stream
    words = cw(top)
    bytes_per_launch = 8.0 * n_loc * words
    achieved = bytes_per_launch / (ms / cnt * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(top)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic if world == 1 else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": ms / cnt,
                "rows_per_launch": n_loc,
                "share_of_loop": ms / sum(v[0] for v in prof.values()),
                "kernels": {c: {"avg_ms": v[0] / v[1], "launches": v[1],
                                "words_per_row": cw(c),
                                "GBps": 8.0 * n_loc * cw(c) / (v[0] / v[1] * 1e-3) / 1e9}
                            for c, v in prof.items() if c in CLASS_WORDS}}
    # ---- the other variants on the same problem (2 solves each, resident inputs)
    variants = {}
    for v in ("hs", "cg", "m", "gv", "pr", "pipe_pr"):
        sess.run(v, its + 1, histories=(), path=args.path)
        info = sess.run(v, its + 1, histories=(), path=args.path)
        lm = info["loop_ms"]
        if world > 1:
            t = torch.tensor([lm], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            lm = t.item()
        ips = its / (lm / 1e3)
        variants[v] = {"iterations_per_s_loop": ips, "ms_per_iteration": lm / its,
                       "hbm_gbs_model": ips * 8 * n * W_V[v] / 1e9,
                       "pct_of_8TBs_per_gpu": 100 * ips * 8 * n * W_V[v] / 8e12 / world}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(args)

    if rank == 0:
        b_iter = 8.0 * n * W_V[variant]
        line = {
            "metric": f"CG iterations/s ({REF_FUN.get(variant, variant)}, Jacobi, {args.dim}-D Poisson {args.grid}^{args.dim})",
            "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"poisson{args.dim}d_{args.grid} {REF_FUN.get(variant, variant)} jacobi (matrix-free {2 * args.dim + 1}-point stencil)",
                       "n": n, "iters_per_step": its, "path": args.path, "partition": f"z-slabs x{world}",
                       "l2": "no flush needed: one iteration streams %.2f GB >> 126 MB L2" % (b_iter / 1e9)},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "wall_ms_per_step": wall_ms / args.steps, "loop_ms_per_iteration": loop_ms / (its * args.steps),
            "hbm_gbs_model": value * b_iter / 1e9, "pct_of_8TBs": 100 * value * b_iter / 8e12 / world,
            "pct_of_measured_peak": 100 * value * b_iter / 1e9 / measured_peak()[0] / world,
            "roofline": roofline, "cpu_baseline": cpu, "variants": variants,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else any library
    prints (e.g. NCCL's version banner) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for the duration of the run
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="pr", choices=sorted(W_V))
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--dim", type=int, default=3, choices=[2, 3],
                    help="3: BASELINE configs[3] (default, the headline); 2 with --grid 4096: configs[2]")
    ap.add_argument("--iters", type=int, default=200, help="CG iterations per step (our arm)")
    ap.add_argument("--ref-iters", type=int, default=3, help="CG iterations per step (reference arm)")
    ap.add_argument("--cpu-iters", type=int, default=10, help="iterations of the cpu_baseline sample")
    ap.add_argument("--path", default="auto", choices=["auto", "stream", "persistent"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.steps < 1 or args.warmup < 0:
        ap.error("steps >= 1, warmup >= 0")
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        ap.error("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
