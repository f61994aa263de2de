"""CPU (-m "not gpu"): host-side logic of the row-partitioned multi-GPU path -- slab
partition, slicing of global vectors, the handle exchange through torch.distributed
(gloo, world_size 2) -- and that a rank fails loudly without a GPU."""
import os
import socket

import numpy as np
import pytest

from new_cg_variants_b200 import PoissonStencil
from new_cg_variants_b200 import dist as cdist


def test_partition_planes_tiles_the_grid():
    for nz in (1, 2, 7, 8, 64, 256):
        for world in (1, 2, 3, 4, 8):
            if nz < world:
                with pytest.raises(ValueError):
                    cdist.partition_planes(nz, world)
                continue
            parts = cdist.partition_planes(nz, world)
            assert parts[0][0] == 0 and parts[-1][1] == nz
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1


def test_row_ranges_and_2d_slab_view():
    S3 = PoissonStencil(6, 5, 9, dim=3)
    assert cdist.slab_grid(S3) == (6, 5, 9)
    ranges = [cdist.row_range(S3, 4, r) for r in range(4)]
    assert ranges[0][0] == 0 and ranges[-1][1] == S3.shape[0]
    assert all(r[0] % 30 == 0 and r[1] % 30 == 0 for r in ranges)
    S2 = PoissonStencil(8, 10, 1, dim=2)            # 2-D: planes are grid rows
    assert cdist.slab_grid(S2) == (8, 1, 10)
    assert cdist.row_range(S2, 2, 1) == (40, 80)
    # the slab view is the same operator: nx x 1 x ny 3-D stencil with the 2-D diagonal
    A2 = S2.tocsr()
    A3 = PoissonStencil(8, 1, 10, dim=3, diag=S2.diag).tocsr()
    assert (A2 != A3).nnz == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        S = PoissonStencil(4, 3, 5, dim=3)
        n = S.shape[0]
        v = np.arange(n, dtype=np.float64)
        r0, r1 = cdist.row_range(S, world, rank)
        # what DistSession does with the window handles: one fixed-size payload per rank, rank order
        payload = bytes([rank]) * 64
        got = cdist.exchange_bytes(payload)
        ok_handles = got == [bytes([r]) * 64 for r in range(world)]
        # gather_x: local slices concatenated in rank order reproduce the global vector
        parts = cdist.exchange_bytes(v[r0:r1].tobytes())
        glob = np.concatenate([np.frombuffer(p, dtype=np.float64) for p in parts])
        ok_gather = np.array_equal(glob, v)
        # a rank without a GPU must fail loudly (no CPU fallback)
        try:
            cdist.DistSession(S, dinv=None, device=0)
            loud = False
        except Exception as e:          # CgxError (no CUDA device) -- never a silent CPU path
            loud = "CUDA" in str(e) or "cuda" in str(e) or "device" in str(e)
        # the two collectives of the reference's scaling_tests.py around the solvers (mpi4py semantics)
        from new_cg_variants_b200 import cg_variants_mpi4py as m
        comm = m.GpuComm()
        part = np.empty(3)
        comm.Scatter(np.arange(6.0).reshape(world, -1) if rank == 0 else None, part, root=0)
        ok_scatter = np.array_equal(part, np.arange(6.0)[3 * rank:3 * rank + 3])
        full = np.empty((world, 3)) if rank == 0 else None
        comm.Gather(part * 2, full, root=0)
        ok_scatter = ok_scatter and (rank != 0 or np.array_equal(full.reshape(-1), 2 * np.arange(6.0)))
        ok_gather = ok_gather and ok_scatter
        q.put((rank, ok_handles, ok_gather, loud))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchange_and_loud_failure():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.is_available():
        pytest.skip("CPU-only test (on a GPU box the real path is exercised by test_gpu_dist.py)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True, True), (1, True, True, True)]


def test_nccl_library_path_resolves():
    p = cdist.nccl_library_path()
    assert p.endswith("libnccl.so.2")


def test_csr_row_partition_lists_reassemble_the_matrix():
    """General CSR row partition: local blocks with remapped columns + the
    per-neighbour send/receive lists reproduce y = A v exactly when the ghost entries are gathered
    the way the device does it (numpy emulation of csr_halo_push + the extended-vector SpMV)."""
    import scipy.sparse as sps
    import helpers
    rng = np.random.default_rng(2)
    mats = [helpers.load_matrix("bcsstk16"), helpers.load_matrix("1138_bus"), sps.diags(rng.uniform(1, 2, 37)).tocsr(),
            helpers.orc.poisson2d(12)]
    for A in mats:
        n = A.shape[0]
        v = rng.standard_normal(n)
        ref = A @ v
        for world in (1, 2, 3, 5, 8):
            ranges = cdist.block_rows(n, world)
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            blks = [cdist.csr_local_block(A[a:b], a, b) for a, b in ranges]
            ghosts = [blk[3] for blk in blks]
            lists = [cdist.csr_exchange_lists(ghosts, ranges, r) for r in range(world)]
            stage = [np.full(len(g), np.nan) for g in ghosts]
            for r in range(world):                                 # every rank pushes what the others need
                recv, send_count, send_idx, send_off, nghost_of = lists[r]
                assert list(nghost_of) == [len(g) for g in ghosts] and recv.sum() == len(ghosts[r])
                pos = 0
                for q in range(world):
                    cnt = int(send_count[q])
                    stage[q][send_off[q]:send_off[q] + cnt] = v[ranges[r][0]:ranges[r][1]][send_idx[pos:pos + cnt]]
                    pos += cnt
            for r, (a, b) in enumerate(ranges):
                indptr, indices, data, ghost = blks[r]
                assert not np.isnan(stage[r]).any()
                ext = np.concatenate([v[a:b], stage[r]])
                Aloc = sps.csr_matrix((data, indices, indptr), shape=(b - a, b - a + len(ghost)))
                assert np.array_equal(Aloc @ ext, ref[a:b])            # same stored order -> same bits
