// cgx_common.cuh -- shared device-side definitions for libcgx_b200 (sm_100a only).
//
// Numerical contract (DESIGN.md section "Arithmetic"): every elementwise update and every
// matrix row sum is evaluated with separately rounded IEEE multiply/add in the operand
// order of the reference's numpy/scipy expressions (no FMA contraction: the __d*_rn
// intrinsics are never fused by nvcc), so those steps are bit-identical to the reference.
// Only the inner products differ from OpenBLAS: they are accumulated with FMA in a fixed,
// run-to-run deterministic order (per-thread strided partial -> warp butterfly -> block ->
// fixed-order cross-block sum by the last-arriving block).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cgx {

typedef long long i64;

constexpr int kBlock = 256;          // threads per CTA for streaming kernels
constexpr int kMaxGrid = 148 * 16;   // upper bound on CTAs of any reducing kernel
constexpr int kNRed = 4;             // at most four fused inner products per pass

// Device-resident scalar recurrences.  One instance per context; written only by the
// finalising thread of a reducing kernel (or by init_scalars), read by every CTA of the
// next kernel.  Kernel boundaries order the accesses.
struct Scal {
  double a;    // alpha_{k-1} when an iteration starts, alpha_k when it ends
  double a1;   // previous alpha
  double b;    // beta the next vector pass applies
  double nu, nu1, mu, eta, del, gam;
  double tmp[8];      // initialisation dot products
  int breakdown;      // -1, or first k with a non-finite alpha/beta
  int pad;
};

// ---- arithmetic that must mirror numpy's two-rounding elementwise expressions ---------
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div_(double a, double b) { return __ddiv_rn(a, b); }
// x + a*p  and  r - a*s  exactly as numpy evaluates `x + a * p`, `r - a * s`
__device__ __forceinline__ double axpy_(double x, double a, double p) { return add_(x, mul_(a, p)); }
__device__ __forceinline__ double axmy_(double r, double a, double s) { return sub_(r, mul_(a, s)); }

// ---- 1- and 2-wide packs for 128-bit global accesses ----------------------------------
template <int W> struct Pk { double v[W]; };

template <int W> __device__ __forceinline__ Pk<W> ldp(const double* __restrict__ p, i64 i) {
  Pk<W> o;
  if constexpr (W == 2) {
    double2 t = *reinterpret_cast<const double2*>(p + i);
    o.v[0] = t.x; o.v[1] = t.y;
  } else {
    o.v[0] = p[i];
  }
  return o;
}
template <int W> __device__ __forceinline__ void stp(double* __restrict__ p, i64 i, const Pk<W>& o) {
  if constexpr (W == 2) {
    *reinterpret_cast<double2*>(p + i) = make_double2(o.v[0], o.v[1]);
  } else {
    p[i] = o.v[0];
  }
}

// ---- deterministic reductions ---------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;   // butterfly: every lane holds the same bits
}

// Sum NR per-thread values over the CTA; result valid in thread 0.  `sh` holds
// NR * (blockDim.x / 32) doubles.
template <int NR>
__device__ __forceinline__ void block_sum(double (&v)[NR], double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int j = 0; j < NR; ++j) v[j] = warp_sum(v[j]);
  __syncthreads();   // protect `sh` against the previous use
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < NR; ++j) sh[j * nw + wid] = v[j];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int j = 0; j < NR; ++j) {
      double t = (lane < nw) ? sh[j * nw + lane] : 0.0;
      v[j] = warp_sum(t);
    }
  }
}

// Grid-wide sum of NR values with a fixed summation order, finished by whichever CTA
// arrives last (ticket counter); `fin(acc)` runs in ONE thread with the totals.
// partials: [gridDim.x][NR].  The order in which CTAs arrive does not influence the bits.
template <int NR, class Fin>
__device__ __forceinline__ void grid_sum_finalize(double (&v)[NR], double* __restrict__ partials,
                                                  unsigned* __restrict__ ticket, Fin fin) {
  __shared__ double sh[NR * (kBlock / 32)];
  __shared__ bool is_last;
  block_sum<NR>(v, sh);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int j = 0; j < NR; ++j) __stcg(&partials[(i64)blockIdx.x * NR + j], v[j]);
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc[NR];
#pragma unroll
  for (int j = 0; j < NR; ++j) acc[j] = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
#pragma unroll
    for (int j = 0; j < NR; ++j) acc[j] += __ldcg(&partials[(i64)i * NR + j]);
  }
  block_sum<NR>(acc, sh);
  if (threadIdx.x == 0) {
    *ticket = 0u;
    fin(acc);
  }
}

__device__ __forceinline__ void note_breakdown(Scal* sc, int k, double a, double b) {
  if (sc->breakdown < 0 && !(isfinite(a) && isfinite(b))) sc->breakdown = k;
}

}  // namespace cgx
