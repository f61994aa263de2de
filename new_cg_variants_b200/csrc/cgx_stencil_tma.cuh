// cgx_stencil_tma.cuh -- TMA-staged matrix-free stencil SpMV with fused epilogues (sm_100a).
//
// One CTA owns a BX x BY column of the grid and marches through a chunk of z-planes.  Each
// plane (with its one-point x/y halo) is brought into shared memory by ONE bulk-tensor copy
// (cp.async.bulk.tensor.3d, SASS UTMALDG) that completes on an mbarrier; out-of-range
// coordinates are zero-filled by the TMA unit, which is exactly the Dirichlet boundary, so
// the load path has no boundary branches.  A ring of kRing planes keeps the copy of plane
// z+2 in flight while plane z is computed from planes z-1, z, z+1, so every input value
// is read from L2/HBM once per CTA column (+ halo) instead of seven times.
//
// The row sum is still evaluated in canonical-CSR term order with separately rounded
// multiply and add (absent neighbours are skipped by a select, not by adding a zero), so
// results stay bit-identical to scipy's `A @ v` on the equivalent CSR matrix.
#pragma once
#include <cuda.h>

#include "cgx_kernels.cuh"

namespace cgx {

constexpr int kTX = 128;              // tile extent in x (points)
constexpr int kTY = 8;                // tile extent in y
constexpr int kPX = kTX + 4;          // box extent incl. halo: starts at x0-2 (TMA needs 16-B aligned starts)
constexpr int kPY = kTY + 2;
constexpr int kPlane = kPX * kPY;     // doubles per staged plane
constexpr int kPlaneStride = ((kPlane * 8 + 127) / 128) * 128 / 8;   // 128-B aligned slots
constexpr int kRing = 4;
constexpr int kTmaThreads = 256;
constexpr int kPtsPerThread = kTX * kTY / kTmaThreads;               // 4

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a copy that never lands (bad descriptor) must fail loudly, not hang the GPU.
// The flag is checked by the host after every synchronisation (CGX_ERR_CUDA).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (global_ns() - t0 > 1000000000ull) { atomicExch(err, 1); return; }
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2),
      "r"(smem_u32(bar))
      : "memory");
}

// Geometry of one launch.  zlo/zhi: does a plane exist below z=0 / above z=nz-1 of THIS
// slab (multi-GPU partition).  Those ghost planes live in the rank's window as LL words,
// stored by the neighbours' vector passes, and are copied into the ring slot by fill_ghost.
struct TmaGeom {
  int nx, ny, nz;
  int ntx, nty;                  // tiles in x, y
  int has_zlo, has_zhi;
  int march_y;                   // 2-D grid: the "planes" of the march are the y-tiles (8 rows each) of a
                                 // column of x-tiles, so consecutive tiles pipeline through the ring like
                                 // z-planes do (nz = number of y-tiles, nty = 1; no z neighbours)
  double diag, off;
  int nchunk;                    // z-chunks per column (kernels that take (column, chunk) work units)
  int* err;                      // device word set when a plane copy did not land within 1 s
};

#define NV_OF(MODE) ((MODE) == SP_PIPE_R ? 2 : 1)
// MODE: SP_* of cgx_kernels.cuh.  PM: 0 identity, 1 Jacobi vector, 2 Jacobi scalar.
// Resident CTAs per SM: 5 (one RHS; shared-memory bound, registers capped to match) or 2.
template <int MODE, int PM, bool MEUR>
__global__ void __launch_bounds__(kTmaThreads, (NV_OF(MODE) == 2) ? 2 : 4)
stencil_tma_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                   const TmaGeom G, const Args g) {
  constexpr int NV = (MODE == SP_PIPE_R) ? 2 : 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // [kRing][NV][kPlaneStride], 128-byte aligned whatever static shared memory precedes it
  double* smem = reinterpret_cast<double*>(
      smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
  __shared__ __align__(8) uint64_t bar[kRing];

  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kRing; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int ly = tid >> 5;              // 0..7   row of the tile
  const int lx = tid & 31;              // lane: points lx, lx+32, lx+64, lx+96
  const i64 plane_pts = (i64)G.nx * G.ny;
  constexpr uint32_t kBytes = (uint32_t)(kPlane * 8 * NV);

  double red[kNRed] = {0.0, 0.0, 0.0, 0.0};
  uint32_t L = 0;                        // loads issued so far by this CTA (ring position)

  // Static even split: the (column, plane) pairs, column-major, are cut into gridDim.x
  // contiguous ranges, so every CTA streams the same number of planes (+-1) whatever the
  // grid shape, and the assignment (hence the summation order of the fused dots) is a
  // function of the problem size only.  A range that crosses a column end is processed as
  // two z-segments.
  const i64 total = (i64)G.ntx * G.nty * G.nz;
  const i64 range_end = ((i64)blockIdx.x + 1) * total / gridDim.x;
  for (i64 pos = (i64)blockIdx.x * total / gridDim.x; pos < range_end;) {
    const int col = (int)(pos / G.nz);
    const int z0 = (int)(pos - (i64)col * G.nz);
    const int z1 = (int)min((i64)G.nz, (i64)z0 + (range_end - pos));
    pos += z1 - z0;
    const int tx = col % G.ntx;
    const int ty = col / G.ntx;
    const int x0 = tx * kTX, y0 = ty * kTY;
    const uint32_t Lbase = L;            // load index of plane z0-1

    auto issue = [&](int z, uint32_t li) {          // one thread: plane z -> slot li % kRing
      const int slot = li % kRing;
      const bool exists = (z >= 0 || G.has_zlo) && (z < G.nz || G.has_zhi);
      if (!exists) {                                // beyond the domain: never read (selects)
        mbar_arrive(&bar[slot]);
        return;
      }
      double* dst = smem + (size_t)slot * NV * kPlaneStride;
      if (z < 0 || z >= G.nz) return;               // ghost plane: filled by fill_ghost (all threads)
      mbar_arrive_expect_tx(&bar[slot], kBytes);
      const int ty0 = G.march_y ? z * kTY : y0, tz = G.march_y ? 0 : z;
      tma_load_3d(dst, &tm0, x0 - 2, ty0 - 1, tz, &bar[slot]);
      if constexpr (NV == 2) tma_load_3d(dst + kPlaneStride, &tm1, x0 - 2, ty0 - 1, tz, &bar[slot]);
    };
    auto wait_load = [&](uint32_t li) { mbar_wait(&bar[li % kRing], (li / kRing) & 1u, G.err); };
    // Ghost plane of a slab (multi-GPU): the neighbour rank's vector pass stored it into this
    // rank's window as LL words (each 8-byte word = half a double + the halo epoch).  All
    // threads poll their elements and write the tile -- zero outside the domain, like the TMA
    // fill -- into the ring slot; no flag, no fence.  Uniform call (every thread of the CTA).
    auto is_ghost = [&](int z) { return (z < 0 && G.has_zlo) || (z >= G.nz && G.has_zhi); };
    auto fill_ghost = [&](int z, uint32_t li) {
      const int slot = li % kRing;
      const int side = z < 0 ? 0 : 1;
      WinHdr* w = g.d.win[g.d.rank];
      double* dst = smem + (size_t)slot * NV * kPlaneStride;
      constexpr int kPer = (kPlane + kTmaThreads - 1) / kTmaThreads;      // elements per thread
      const u64 tag = g.hin_epoch & 0xffffffffull;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const u64* q = g.d.ghl + ghl_off(g.d, g.hin_ch + v, g.hin_par, side);
        u64 lo[kPer], hi[kPer];
        const u64* src[kPer];
        // all loads first (16 bytes = both words of a value), then validate / re-poll
#pragma unroll
        for (int m = 0; m < kPer; ++m) {
          const int idx = tid + m * kTmaThreads;
          const int py = idx / kPX, px = idx - py * kPX;
          const int gx = x0 - 2 + px, gyy = y0 - 1 + py;
          src[m] = (idx < kPlane && gx >= 0 && gx < G.nx && gyy >= 0 && gyy < G.ny) ? q + 2 * (gyy * G.nx + gx) : nullptr;
          lo[m] = hi[m] = tag << 32;                                       // outside the domain: +0.0, "valid"
          if (g.dbg & 1) src[m] = nullptr;                                 // timing experiment: no halo traffic
          if (src[m])
            asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo[m]), "=l"(hi[m]) : "l"(src[m]) : "memory");
        }
#pragma unroll
        for (int m = 0; m < kPer; ++m) {
          const int idx = tid + m * kTmaThreads;
          if (src[m] && ((lo[m] >> 32) != tag || (hi[m] >> 32) != tag)) {
            lo[m] = ll_poll(src[m], tag, &w->error);
            hi[m] = ll_poll(src[m] + 1, tag, &w->error);
          }
          if (idx < kPlane)
            dst[v * kPlaneStride + idx] = __longlong_as_double((long long)((lo[m] & 0xffffffffull) | (hi[m] << 32)));
        }
      }
      __syncthreads();
      if (tid == 0) mbar_arrive(&bar[slot]);
    };

    if (tid == 0) {
      issue(z0 - 1, Lbase);
      issue(z0, Lbase + 1);
      issue(z0 + 1, Lbase + 2);
    }
    if (is_ghost(z0 - 1)) fill_ghost(z0 - 1, Lbase);
    if (is_ghost(z0 + 1)) fill_ghost(z0 + 1, Lbase + 2);
    wait_load(Lbase);
    wait_load(Lbase + 1);

    int gy = y0 + ly;                                  // (march_y: set per step)
    bool row_ok = gy < G.ny;

    // Operands that do not go through shared memory (r for the fused dots, a Jacobi vector)
    // are fetched ONE PLANE AHEAD into registers: their global-load latency is covered by a
    // whole plane of work instead of stalling the first fused dot (ncu: 29 % of the stall
    // samples sat on that DFMA when the load was issued in the same iteration).
    constexpr bool kNeedR = (MODE == SP_CG || MODE == SP_PR);
    constexpr bool kNeedD = (MODE == SP_PR && PM == 1);
    double rv_n[kPtsPerThread], dvv_n[kPtsPerThread];
    auto fetch_direct = [&](int z) {
      const int gyz = G.march_y ? z * kTY + ly : gy;
      const bool rok = gyz < G.ny;
      const i64 ib = (G.march_y ? 0 : (i64)z * plane_pts) + (i64)gyz * G.nx + x0 + lx;
#pragma unroll
      for (int m = 0; m < kPtsPerThread; ++m) {
        const bool ok = rok && (x0 + lx + 32 * m) < G.nx;
        rv_n[m] = (kNeedR && ok) ? g.r[ib + 32 * m] : 0.0;
        dvv_n[m] = (kNeedD && ok) ? g.dinv[ib + 32 * m] : 0.0;
      }
    };
    if constexpr (kNeedR || kNeedD) fetch_direct(z0);

    for (int z = z0; z < z1; ++z) {
      const uint32_t j = (uint32_t)(z - z0);
      if (tid == 0 && z + 2 <= z1) issue(z + 2, Lbase + j + 3);    // slot of plane z-2: free
      if (z + 2 <= z1 && is_ghost(z + 2)) fill_ghost(z + 2, Lbase + j + 3);
      double rv[kPtsPerThread], dvv[kPtsPerThread];
      if (G.march_y) { gy = z * kTY + ly; row_ok = gy < G.ny; }
      const i64 ibase = (G.march_y ? 0 : (i64)z * plane_pts) + (i64)gy * G.nx + x0 + lx;
#pragma unroll
      for (int m = 0; m < kPtsPerThread; ++m) { rv[m] = rv_n[m]; dvv[m] = dvv_n[m]; }
      if constexpr (kNeedR || kNeedD) { if (z + 1 < z1) fetch_direct(z + 1); }
      wait_load(Lbase + j + 2);                                     // plane z+1 has landed
      const double* pm = smem + (size_t)((Lbase + j) % kRing) * NV * kPlaneStride;
      const double* pc = smem + (size_t)((Lbase + j + 1) % kRing) * NV * kPlaneStride;
      const double* pp = smem + (size_t)((Lbase + j + 2) % kRing) * NV * kPlaneStride;
      const bool has_zm = !G.march_y && ((z > 0) || G.has_zlo);
      const bool has_zp = !G.march_y && ((z < G.nz - 1) || G.has_zhi);

      if (row_ok) {
#pragma unroll
        for (int m = 0; m < kPtsPerThread; ++m) {
          const int px = lx + 32 * m;              // 0..127 within the tile
          const int gx = x0 + px;
          if (gx < G.nx) {
            const int c = (ly + 1) * kPX + (px + 2);
            double y[NV], ctr[NV], raw[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              const double* qm = pm + v * kPlaneStride;
              const double* qc = pc + v * kPlaneStride;
              const double* qp = pp + v * kPlaneStride;
              // x/y neighbours outside the domain were zero-filled by the TMA unit: their term is
              // off * (+0.0) = +-0.0, and adding a signed zero never changes a running sum that
              // started at +0.0 (such a sum can never be -0.0), so no select is needed and the
              // bits equal scipy's sum over the stored entries only.  Absent z-planes are not
              // loaded at all (stale shared memory): those two terms keep their select.
              // SP_CG_E: the tile holds r; the operand is r~ = M r, formed here (the product
              // EW_CG would have stored), so r~ is neither written nor read from HBM
              auto in = [&](double q) {
                if constexpr ((MODE == SP_CG_E || MODE == SP_GV_E) && PM == 2) return mul_(g.dinv_s, q);
                else return q;
              };
              double acc = 0.0, t;
              t = add_(acc, mul_(G.off, in(qm[c])));        acc = has_zm ? t : acc;
              acc = add_(acc, mul_(G.off, in(qc[c - kPX])));
              acc = add_(acc, mul_(G.off, in(qc[c - 1])));
              raw[v] = qc[c];
              ctr[v] = in(raw[v]);
              acc = add_(acc, mul_(G.diag, ctr[v]));
              acc = add_(acc, mul_(G.off, in(qc[c + 1])));
              acc = add_(acc, mul_(G.off, in(qc[c + kPX])));
              t = add_(acc, mul_(G.off, in(qp[c])));        acc = has_zp ? t : acc;
              y[v] = acc;
            }
            const i64 i = ibase + 32 * m;
            auto M = [&](double v) {
              if constexpr (PM == 1) return mul_(dvv[m], v);
              else if constexpr (PM == 2) return mul_(g.dinv_s, v);
              else return v;
            };
            if constexpr (MODE == SP_HS) {                 // hs_cg.py:123-124
              g.s[i] = y[0];
              red[0] = fma(ctr[0], y[0], red[0]);
            } else if constexpr (MODE == SP_CG_E) {        // cg_cg.py:133-135 with r from the tile
              g.w[i] = y[0];
              red[0] = fma(raw[0], ctr[0], red[0]);
              red[1] = fma(y[0], ctr[0], red[1]);
            } else if constexpr (MODE == SP_CG) {          // cg_cg.py:133-135
              g.w[i] = y[0];
              red[0] = fma(rv[m], ctr[0], red[0]);
              red[1] = fma(y[0], ctr[0], red[1]);
            } else if constexpr (MODE == SP_GV || MODE == SP_GV_E) {          // gv_cg.py:161
              g.t[i] = y[0];
            } else if constexpr (MODE == SP_PR) {          // pr_cg.py:152-156
              g.s[i] = y[0];
              const double sti = M(y[0]);
              red[0] = fma(ctr[0], y[0], red[0]);
              red[1] = fma(rv[m], sti, red[1]);
              red[2] = fma(sti, y[0], red[2]);
            } else if constexpr (MODE == SP_PIPE_R) {      // pipe_pr_cg.py:179-182
              g.u[i] = y[0];
              g.w[i] = y[NV - 1];
            } else {                                       // SP_PIPE_N
              g.u[i] = y[0];
            }
          }
        }
      }
      __syncthreads();       // everyone is done with plane z-1 before its slot is refilled
    }
    L = Lbase + (uint32_t)(z1 - z0) + 2;   // planes z0-1 .. z1 were issued
  }

  spmv_close<MODE, MEUR>(g, red);
}

}  // namespace cgx
