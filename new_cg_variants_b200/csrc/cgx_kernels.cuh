// cgx_kernels.cuh -- streaming-path kernels: operators, fused vector passes, fused
// SpMV passes, instrumentation.  One template per dependency stage; the variant is a
// compile-time tag.  Reference statements: SURVEY.md section 8 / the file:line cited at each
// body (paths relative to predict_and_recompute/numerical_experiments/cg_variants/).
#pragma once
#include "cgx_common.cuh"

namespace cgx {

// =====================================================================================
// Operators.  row<NV>(i, ld, y): y[c] = sum_j A_ij * v_c[j] for NV right-hand sides whose
// entries are produced by ld(j, vals).  Accumulation mirrors scipy's csr_matvec
// (sum = 0; sum += a*v, stored order, multiply and add rounded separately).
// =====================================================================================
struct CsrOp {
  const int* __restrict__ ptr;
  const int* __restrict__ idx;
  const double* __restrict__ val;
  i64 n;
  template <int NV, class Ld>
  __device__ __forceinline__ void row(i64 i, Ld ld, double (&y)[NV]) const {
#pragma unroll
    for (int c = 0; c < NV; ++c) y[c] = 0.0;
    const int e = __ldg(ptr + i + 1);
    for (int jj = __ldg(ptr + i); jj < e; ++jj) {
      const double a = __ldg(val + jj);
      double v[NV];
      ld((i64)__ldg(idx + jj), v);
#pragma unroll
      for (int c = 0; c < NV; ++c) y[c] = add_(y[c], mul_(a, v[c]));
    }
  }
};

// Matrix-free Dirichlet Poisson stencil, natural ordering.  Visits the neighbours in
// ascending column order (z-1, y-1, x-1, centre, x+1, y+1, z+1), i.e. exactly the order
// of the canonical CSR matrix, so the result is bit-identical to scipy on that matrix.
struct StencilOp {
  int nx, ny, nz;
  double diag, off;
  i64 n;
  template <int NV, class Ld>
  __device__ __forceinline__ void row(i64 i, Ld ld, double (&y)[NV]) const {
    const int plane = nx * ny;
    const int ii = (int)i;
    const int z = ii / plane;
    const int rem = ii - z * plane;
    const int yy = rem / nx;
    const int xx = rem - yy * nx;
    double v[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) y[c] = 0.0;
#define CGX_ST_TERM(cond, j, coef)                                         \
    if (cond) {                                                            \
      ld((i64)(j), v);                                                     \
      _Pragma("unroll") for (int c = 0; c < NV; ++c) y[c] = add_(y[c], mul_((coef), v[c])); \
    }
    CGX_ST_TERM(z > 0, ii - plane, off)
    CGX_ST_TERM(yy > 0, ii - nx, off)
    CGX_ST_TERM(xx > 0, ii - 1, off)
    CGX_ST_TERM(true, ii, diag)
    CGX_ST_TERM(xx < nx - 1, ii + 1, off)
    CGX_ST_TERM(yy < ny - 1, ii + nx, off)
    CGX_ST_TERM(z < nz - 1, ii + plane, off)
#undef CGX_ST_TERM
  }
};

// =====================================================================================
// Kernel argument block (plain pointers; unused ones are null).
// =====================================================================================
struct Args {
  double* x; double* r; double* rt; double* p; double* s; double* st;
  double* w; double* wt; double* u; double* t;
  const double* dinv; double dinv_s;
  const double* b; const double* xtrue;
  Scal* sc; double* partials; unsigned* ticket;
  double* hist; int hist_len; unsigned hist_mask;
  i64 n; int k;
};

// Stage tags
enum {
  EW_HS1 = 0,   // r -= a s ; nu = r.(M r)
  EW_HS2,       // x += a p ; p = M r + b p
  EW_CG,        // [deferred p,s] ; x,r ; rt = M r
  EW_GV,        // [deferred p,s,st,u] ; x,r,rt,w ; wt = M w ; nu,eta
  EW_PR,        // x,r,rt ; p ; nu = rt.r
  EW_PIPE_R,    // pipe family with recompute of w (pipe_pr, pipe_pr_m)
  EW_PIPE_N     // pipe family without recompute (pipe_p, pipe_p_m)
};
enum { SP_PLAIN = 0, SP_HS, SP_CG, SP_GV, SP_PR, SP_PIPE_R, SP_PIPE_N, SP_RESID };

template <int KID> struct EwTraits { static constexpr int NR = 0; };
template <> struct EwTraits<EW_HS1> { static constexpr int NR = 1; };
template <> struct EwTraits<EW_GV> { static constexpr int NR = 2; };
template <> struct EwTraits<EW_PR> { static constexpr int NR = 1; };
template <> struct EwTraits<EW_PIPE_R> { static constexpr int NR = 4; };
template <> struct EwTraits<EW_PIPE_N> { static constexpr int NR = 4; };

// Predicted nu and the beta it gives (pr_cg.py:149-150, pipe_pr_cg.py:174-175):
//   PR: nu' = nu - 2 a del + a^2 gam      M: nu' = -nu + a^2 gam      b = nu'/nu
__device__ __forceinline__ double predict_beta(bool meurant, double nu, double a, double del,
                                               double gam) {
  const double a2g = mul_(mul_(a, a), gam);
  const double nup = meurant ? add_(-nu, a2g) : add_(sub_(nu, mul_(mul_(2.0, a), del)), a2g);
  return div_(nup, nu);
}

// -------------------------------------------------------------------------------------
// Fused vector pass: one HBM sweep over the state vectors of stage KID.
// -------------------------------------------------------------------------------------
// PM: preconditioner mode -- 0 identity, 1 Jacobi vector g.dinv, 2 Jacobi scalar g.dinv_s
// (a constant diagonal, e.g. every Poisson stencil: same products, one HBM stream less).
template <int KID, int PM, int W>
__device__ __forceinline__ void ew_body(const Args& g, i64 i, double a, double b,
                                        double (&red)[kNRed]) {
  constexpr bool PREC = PM != 0;
  Pk<W> dv{};
  if constexpr (PM == 1) dv = ldp<W>(g.dinv, i);
  const double ds = g.dinv_s;
  auto M = [&](double v, int l) { return PM == 1 ? mul_(dv.v[l], v) : (PM == 2 ? mul_(ds, v) : v); };

  if constexpr (KID == EW_HS1) {             // hs_cg.py:118-120
    Pk<W> r = ldp<W>(g.r, i), s = ldp<W>(g.s, i);
#pragma unroll
    for (int l = 0; l < W; ++l) {
      r.v[l] = axmy_(r.v[l], a, s.v[l]);
      red[0] = fma(r.v[l], M(r.v[l], l), red[0]);
    }
    stp<W>(g.r, i, r);
  } else if constexpr (KID == EW_HS2) {      // hs_cg.py:117,119,122
    Pk<W> x = ldp<W>(g.x, i), p = ldp<W>(g.p, i), r = ldp<W>(g.r, i);
#pragma unroll
    for (int l = 0; l < W; ++l) {
      x.v[l] = axpy_(x.v[l], a, p.v[l]);
      p.v[l] = axpy_(M(r.v[l], l), b, p.v[l]);
    }
    stp<W>(g.x, i, x); stp<W>(g.p, i, p);
  } else if constexpr (KID == EW_CG) {       // cg_cg.py:137-138 (deferred), :130-132
    Pk<W> x = ldp<W>(g.x, i), r = ldp<W>(g.r, i), rt = ldp<W>(g.rt, i), p = ldp<W>(g.p, i),
          s = ldp<W>(g.s, i), w = ldp<W>(g.w, i);
#pragma unroll
    for (int l = 0; l < W; ++l) {
      p.v[l] = axpy_(rt.v[l], b, p.v[l]);
      s.v[l] = axpy_(w.v[l], b, s.v[l]);
      x.v[l] = axpy_(x.v[l], a, p.v[l]);
      r.v[l] = axmy_(r.v[l], a, s.v[l]);
      rt.v[l] = M(r.v[l], l);
    }
    stp<W>(g.p, i, p); stp<W>(g.s, i, s); stp<W>(g.x, i, x); stp<W>(g.r, i, r);
    stp<W>(g.rt, i, rt);
  } else if constexpr (KID == EW_GV) {       // gv_cg.py:165-168 (deferred), :151-154,160,162-163
    Pk<W> x = ldp<W>(g.x, i), r = ldp<W>(g.r, i), rt = ldp<W>(g.rt, i), p = ldp<W>(g.p, i),
          s = ldp<W>(g.s, i), st = ldp<W>(g.st, i), w = ldp<W>(g.w, i), wt = ldp<W>(g.wt, i),
          u = ldp<W>(g.u, i), t = ldp<W>(g.t, i);
#pragma unroll
    for (int l = 0; l < W; ++l) {
      p.v[l] = axpy_(rt.v[l], b, p.v[l]);
      s.v[l] = axpy_(w.v[l], b, s.v[l]);
      st.v[l] = axpy_(wt.v[l], b, st.v[l]);
      u.v[l] = axpy_(t.v[l], b, u.v[l]);
      x.v[l] = axpy_(x.v[l], a, p.v[l]);
      r.v[l] = axmy_(r.v[l], a, s.v[l]);
      rt.v[l] = axmy_(rt.v[l], a, st.v[l]);
      w.v[l] = axmy_(w.v[l], a, u.v[l]);
      wt.v[l] = M(w.v[l], l);
      red[0] = fma(r.v[l], rt.v[l], red[0]);
      red[1] = fma(w.v[l], rt.v[l], red[1]);
    }
    stp<W>(g.p, i, p); stp<W>(g.s, i, s); stp<W>(g.st, i, st); stp<W>(g.u, i, u);
    stp<W>(g.x, i, x); stp<W>(g.r, i, r); stp<W>(g.rt, i, rt); stp<W>(g.w, i, w);
    stp<W>(g.wt, i, wt);
  } else if constexpr (KID == EW_PR) {       // pr_cg.py:146-148,151,157
    Pk<W> x = ldp<W>(g.x, i), r = ldp<W>(g.r, i), rt = ldp<W>(g.rt, i), p = ldp<W>(g.p, i),
          s = ldp<W>(g.s, i);
#pragma unroll
    for (int l = 0; l < W; ++l) {
      x.v[l] = axpy_(x.v[l], a, p.v[l]);
      r.v[l] = axmy_(r.v[l], a, s.v[l]);
      rt.v[l] = axmy_(rt.v[l], a, M(s.v[l], l));       // st_{k-1} = M s_{k-1} exactly
      p.v[l] = axpy_(rt.v[l], b, p.v[l]);
      red[0] = fma(rt.v[l], r.v[l], red[0]);
    }
    stp<W>(g.x, i, x); stp<W>(g.r, i, r); stp<W>(g.rt, i, rt); stp<W>(g.p, i, p);
  } else {                                   // pipe_pr_cg.py:169-178,183-186
    constexpr bool RECOMP = (KID == EW_PIPE_R);
    Pk<W> x = ldp<W>(g.x, i), r = ldp<W>(g.r, i), rt = ldp<W>(g.rt, i), p = ldp<W>(g.p, i),
          s = ldp<W>(g.s, i), st = ldp<W>(g.st, i), w = ldp<W>(g.w, i), u = ldp<W>(g.u, i);
    Pk<W> wt;
    if constexpr (!RECOMP && PREC) wt = ldp<W>(g.wt, i);
#pragma unroll
    for (int l = 0; l < W; ++l) {
      x.v[l] = axpy_(x.v[l], a, p.v[l]);
      r.v[l] = axmy_(r.v[l], a, s.v[l]);
      rt.v[l] = axmy_(rt.v[l], a, st.v[l]);
      // wt_{k-1}: = M w_{k-1} exactly when w is recomputed every iteration (or M = I),
      // otherwise its own recurrence.  ut_{k-1} = M u_{k-1} always.
      double wt_old;
      if constexpr (!RECOMP && PREC) wt_old = wt.v[l]; else wt_old = M(w.v[l], l);
      const double wn = axmy_(w.v[l], a, u.v[l]);
      const double wtn = axmy_(wt_old, a, M(u.v[l], l));
      p.v[l] = axpy_(rt.v[l], b, p.v[l]);
      s.v[l] = axpy_(wn, b, s.v[l]);
      st.v[l] = axpy_(wtn, b, st.v[l]);
      w.v[l] = wn;
      if constexpr (!RECOMP && PREC) wt.v[l] = wtn;
      red[0] = fma(p.v[l], s.v[l], red[0]);      // mu
      red[1] = fma(r.v[l], st.v[l], red[1]);     // delta
      red[2] = fma(st.v[l], s.v[l], red[2]);     // gamma
      red[3] = fma(rt.v[l], r.v[l], red[3]);     // nu (recomputed)
    }
    stp<W>(g.x, i, x); stp<W>(g.r, i, r); stp<W>(g.rt, i, rt); stp<W>(g.p, i, p);
    stp<W>(g.s, i, s); stp<W>(g.st, i, st);
    if constexpr (!RECOMP) {
      stp<W>(g.w, i, w);
      if constexpr (PREC) stp<W>(g.wt, i, wt);
    }
  }
}

template <int KID, int PM, bool MEURANT>
__global__ void __launch_bounds__(kBlock) ew_kernel(const Args g) {
  const double a = g.sc->a, b = g.sc->b;
  double red[kNRed] = {0.0, 0.0, 0.0, 0.0};
  const i64 nv = g.n >> 1;
  const i64 stride = (i64)gridDim.x * kBlock;
  for (i64 i = (i64)blockIdx.x * kBlock + threadIdx.x; i < nv; i += stride)
    ew_body<KID, PM, 2>(g, 2 * i, a, b, red);
  if ((g.n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
    ew_body<KID, PM, 1>(g, g.n - 1, a, b, red);

  constexpr int NR = EwTraits<KID>::NR;
  if constexpr (NR > 0) {
    double v[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) v[j] = red[j];
    Scal* sc = g.sc;
    const int k = g.k;
    grid_sum_finalize<NR>(v, g.partials, g.ticket, [=](const double* acc) {
      if constexpr (KID == EW_HS1) {            // hs_cg.py:120-121
        const double nu1 = sc->nu;
        sc->nu1 = nu1; sc->nu = acc[0];
        sc->b = div_(acc[0], nu1);
        note_breakdown(sc, k, sc->a, sc->b);
      } else if constexpr (KID == EW_GV) {      // gv_cg.py:162-164,169-170
        const double nu1 = sc->nu, a1 = sc->a, nu = acc[0], eta = acc[1];
        const double bb = div_(nu, nu1);
        const double mu = sub_(eta, mul_(div_(bb, a1), nu));
        sc->nu1 = nu1; sc->nu = nu; sc->eta = eta; sc->b = bb; sc->mu = mu;
        sc->a1 = a1; sc->a = div_(nu, mu);
        note_breakdown(sc, k, sc->a, bb);
      } else if constexpr (KID == EW_PR) {      // pr_cg.py:157 (consumed by the SpMV pass)
        sc->nu1 = sc->nu; sc->nu = acc[0];
      } else {                                  // pipe_pr_cg.py:183-187 then :174-175
        const double mu = acc[0], del = acc[1], gam = acc[2], nu = acc[3];
        sc->nu1 = sc->nu; sc->nu = nu; sc->mu = mu; sc->del = del; sc->gam = gam;
        const double an = div_(nu, mu);
        sc->a1 = sc->a; sc->a = an;
        sc->b = predict_beta(MEURANT, nu, an, del, gam);
        note_breakdown(sc, k, an, sc->b);
      }
    });
  }
}

// -------------------------------------------------------------------------------------
// Fused SpMV pass: one sweep over the matrix (or stencil) with the stage's epilogue.
// -------------------------------------------------------------------------------------
template <int MODE> struct SpTraits { static constexpr int NR = 0; };
template <> struct SpTraits<SP_HS> { static constexpr int NR = 1; };
template <> struct SpTraits<SP_CG> { static constexpr int NR = 2; };
template <> struct SpTraits<SP_PR> { static constexpr int NR = 3; };

// Scalar recurrences that close a fused SpMV pass (run by one thread with the grid totals).
template <int MODE, bool MEURANT>
__device__ __forceinline__ void spmv_finalize(Scal* sc, int k, const double* acc) {
  if constexpr (MODE == SP_HS) {             // hs_cg.py:124-125
    sc->mu = acc[0];
    sc->a1 = sc->a; sc->a = div_(sc->nu, acc[0]);
    note_breakdown(sc, k, sc->a, sc->b);
  } else if constexpr (MODE == SP_CG) {      // cg_cg.py:134-136,139-140
    const double nu1 = sc->nu, a1 = sc->a, nu = acc[0], eta = acc[1];
    const double bb = div_(nu, nu1);
    const double mu = sub_(eta, mul_(div_(bb, a1), nu));
    sc->nu1 = nu1; sc->nu = nu; sc->eta = eta; sc->b = bb; sc->mu = mu;
    sc->a1 = a1; sc->a = div_(nu, mu);
    note_breakdown(sc, k, sc->a, bb);
  } else if constexpr (MODE == SP_PR) {      // pr_cg.py:154-158 then :149-150
    const double mu = acc[0], del = acc[1], gam = acc[2], nu = sc->nu;
    sc->mu = mu; sc->del = del; sc->gam = gam;
    const double an = div_(nu, mu);
    sc->a1 = sc->a; sc->a = an;
    sc->b = predict_beta(MEURANT, nu, an, del, gam);
    note_breakdown(sc, k, an, sc->b);
  }
}

template <class Op, int MODE, int PM, bool MEURANT>
__global__ void __launch_bounds__(kBlock) spmv_kernel(const Op A, const Args g, const double* vin,
                                                     double* vout) {
  double red[kNRed] = {0.0, 0.0, 0.0, 0.0};
  const i64 stride = (i64)gridDim.x * kBlock;
  for (i64 i = (i64)blockIdx.x * kBlock + threadIdx.x; i < g.n; i += stride) {
    if constexpr (MODE == SP_PLAIN) {            // y = A v
      double y[1];
      A.template row<1>(i, [&](i64 j, double (&v)[1]) { v[0] = vin[j]; }, y);
      vout[i] = y[0];
    } else if constexpr (MODE == SP_RESID) {     // r = b - A x0   (e.g. hs_cg.py:84)
      double y[1];
      A.template row<1>(i, [&](i64 j, double (&v)[1]) { v[0] = vin[j]; }, y);
      vout[i] = sub_(g.b[i], y[0]);
    } else if constexpr (MODE == SP_HS) {        // hs_cg.py:123-124
      double y[1];
      A.template row<1>(i, [&](i64 j, double (&v)[1]) { v[0] = g.p[j]; }, y);
      g.s[i] = y[0];
      red[0] = fma(g.p[i], y[0], red[0]);
    } else if constexpr (MODE == SP_CG) {        // cg_cg.py:133-135
      double y[1];
      A.template row<1>(i, [&](i64 j, double (&v)[1]) { v[0] = g.rt[j]; }, y);
      g.w[i] = y[0];
      const double rti = g.rt[i];
      red[0] = fma(g.r[i], rti, red[0]);
      red[1] = fma(y[0], rti, red[1]);
    } else if constexpr (MODE == SP_GV) {        // gv_cg.py:161
      double y[1];
      A.template row<1>(i, [&](i64 j, double (&v)[1]) { v[0] = g.wt[j]; }, y);
      g.t[i] = y[0];
    } else if constexpr (MODE == SP_PR) {        // pr_cg.py:152-156
      double y[1];
      A.template row<1>(i, [&](i64 j, double (&v)[1]) { v[0] = g.p[j]; }, y);
      g.s[i] = y[0];
      const double sti = PM == 1 ? mul_(g.dinv[i], y[0]) : (PM == 2 ? mul_(g.dinv_s, y[0]) : y[0]);
      red[0] = fma(g.p[i], y[0], red[0]);
      red[1] = fma(g.r[i], sti, red[1]);
      red[2] = fma(sti, y[0], red[2]);
    } else if constexpr (MODE == SP_PIPE_R) {    // pipe_pr_cg.py:179-182: one matrix pass, 2 RHS
      double y[2];
      A.template row<2>(i, [&](i64 j, double (&v)[2]) { v[0] = g.st[j]; v[1] = g.rt[j]; }, y);
      g.u[i] = y[0];
      g.w[i] = y[1];
    } else {                                     // SP_PIPE_N: pipe_pr_cg.py:179-180
      double y[1];
      A.template row<1>(i, [&](i64 j, double (&v)[1]) { v[0] = g.st[j]; }, y);
      g.u[i] = y[0];
    }
  }
  constexpr int NR = SpTraits<MODE>::NR;
  if constexpr (NR > 0) {
    double v[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) v[j] = red[j];
    Scal* sc = g.sc;
    const int k = g.k;
    grid_sum_finalize<NR>(v, g.partials, g.ticket, [=](const double* acc) {
      spmv_finalize<MODE, MEURANT>(sc, k, acc);
    });
  }
}

// -------------------------------------------------------------------------------------
// Instrumentation = the four standard callbacks in one matrix pass (callbacks/*.py):
//   e = x - x_true ; error_A_norm = sqrt(e.(A e)) ; residual_2_norm = ||b - A x|| ;
//   error_2_norm = ||e|| ; updated_residual_2_norm = ||r||.
// Excluded from the roofline traffic model.
// -------------------------------------------------------------------------------------
template <class Op, bool HAS_XTRUE>
__global__ void __launch_bounds__(kBlock) instrument_kernel(const Op A, const Args g) {
  double red[4] = {0.0, 0.0, 0.0, 0.0};
  const i64 stride = (i64)gridDim.x * kBlock;
  for (i64 i = (i64)blockIdx.x * kBlock + threadIdx.x; i < g.n; i += stride) {
    double y[2] = {0.0, 0.0};
    if constexpr (HAS_XTRUE) {
      A.template row<2>(i, [&](i64 j, double (&v)[2]) {
        v[0] = g.x[j]; v[1] = sub_(v[0], g.xtrue[j]); }, y);
      const double e = sub_(g.x[i], g.xtrue[i]);
      red[0] = fma(e, y[1], red[0]);
      red[2] = fma(e, e, red[2]);
    } else {
      double y1[1];
      A.template row<1>(i, [&](i64 j, double (&v)[1]) { v[0] = g.x[j]; }, y1);
      y[0] = y1[0];
    }
    const double res = sub_(g.b[i], y[0]);
    red[1] = fma(res, res, red[1]);
    const double ri = g.r[i];
    red[3] = fma(ri, ri, red[3]);
  }
  double* hist = g.hist;
  const int L = g.hist_len, k = g.k;
  const unsigned mask = g.hist_mask;
  grid_sum_finalize<4>(red, g.partials, g.ticket, [=](const double* acc) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (mask & (1u << j)) hist[(i64)j * L + k] = sqrt(acc[j]);
  });
}

// -------------------------------------------------------------------------------------
// Small helpers used only by the (non-timed) initialisation.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) scale_kernel(const double* __restrict__ dinv,
                                                      const double* __restrict__ v,
                                                      double* __restrict__ out, i64 n) {
  const i64 stride = (i64)gridDim.x * kBlock;
  for (i64 i = (i64)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride)
    out[i] = dinv ? mul_(dinv[i], v[i]) : v[i];
}

// sc->tmp[slot] = u . (dinv ? dinv*v : v)
__global__ void __launch_bounds__(kBlock) dot_kernel(const double* __restrict__ u,
                                                    const double* __restrict__ v,
                                                    const double* __restrict__ dinv, i64 n,
                                                    Scal* sc, int slot, double* partials,
                                                    unsigned* ticket) {
  double red[1] = {0.0};
  const i64 stride = (i64)gridDim.x * kBlock;
  for (i64 i = (i64)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
    const double vi = dinv ? mul_(dinv[i], v[i]) : v[i];
    red[0] = fma(u[i], vi, red[0]);
  }
  grid_sum_finalize<1>(red, partials, ticket, [=](const double* acc) { sc->tmp[slot] = acc[0]; });
}

// Initial scalars from the initialisation dots.  tmp: 0 nu, 1 mu, 2 eta, 3 delta, 4 gamma.
__global__ void init_scalars_kernel(Scal* sc, int variant_class, int meurant) {
  // variant_class: 0 HS, 1 CG/GV (mu := p.s), 2 PR/M/pipe (predict first beta)
  const double nu = sc->tmp[0], mu = sc->tmp[1];
  sc->nu = nu; sc->nu1 = nu; sc->mu = mu; sc->eta = sc->tmp[2];
  sc->del = sc->tmp[3]; sc->gam = sc->tmp[4];
  sc->a1 = 0.0;
  sc->a = div_(nu, mu);
  sc->b = 0.0;
  sc->breakdown = -1;
  if (variant_class == 2) sc->b = predict_beta(meurant != 0, nu, sc->a, sc->del, sc->gam);
  note_breakdown(sc, 0, sc->a, sc->b);
}

}  // namespace cgx
