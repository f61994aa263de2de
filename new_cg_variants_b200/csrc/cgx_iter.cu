// cgx_iter.cu -- stage s of one iteration for one preconditioner mode (compiled three times:
// -DCGX_PM=0 identity, 1 Jacobi vector, 2 Jacobi with a constant diagonal).
#include "cgx_launch.cuh"

#ifndef CGX_PM
#error "compile with -DCGX_PM=0|1|2"
#endif

// CSR row partition: the ghost entries of the vector(s) the SpMV pass is about to read
static void csr_push_stage(cgx_ctx* c, const Args& g) {
  switch (c->variant) {
    case CGX_HS: case CGX_PR: case CGX_M: launch_halo_push(c, g, c->vec[V_P], 0); break;
    case CGX_CG: launch_halo_push(c, g, c->vec[V_RT], 0); break;
    case CGX_GV: launch_halo_push(c, g, c->vec[V_WT], 0); break;
    case CGX_PIPE_PR: case CGX_PIPE_PR_M:
      launch_halo_push2(c, g, c->vec[V_ST], c->vec[V_RT], 0);
      break;
    default: launch_halo_push(c, g, c->vec[V_ST], 0); break;      // pipe_p, pipe_p_m
  }
}

template <int PM>
static void iter_stage_pm(cgx_ctx* c, int s, const Args& g) {
  const int core = core_stages(c);
  if (s >= core) {
    const int t = s - core;
    if (c->dist.world > 1) {
      if (t == 0) launch_halo_push(c, g, c->vec[V_X], 2);
      else if (t == 1) launch_instrument(c, g);
      else launch_hist_consume(c, g);
    } else {
      if (c->hist_mask) launch_instrument(c, g);
      if (c->capture) launch_capture(c, g);
    }
    return;
  }
  if (csr_dist(c)) {                 // ... ew stage(s), push, SpMV pass
    if (s == core - 2) { csr_push_stage(c, g); return; }
    if (s == core - 1) s -= 1;
  }
  switch (c->variant) {
    case CGX_HS:
      if (s == 0) launch_ew<EW_HS1, PM, false>(c, g);
      else if (s == 1) launch_ew<EW_HS2, PM, false>(c, g);
      else launch_spmv<SP_HS, PM, false>(c, g, nullptr, nullptr);
      break;
    case CGX_CG:
      if constexpr (PM != 1) {
        if (c->cg_elide) {
          if (s == 0) launch_ew<EW_CG_E, PM, false>(c, g); else launch_spmv<SP_CG_E, PM, false>(c, g, nullptr, nullptr);
          break;
        }
      }
      if (s == 0) launch_ew<EW_CG, PM, false>(c, g); else launch_spmv<SP_CG, PM, false>(c, g, nullptr, nullptr);
      break;
    case CGX_GV:
      if constexpr (PM != 1) {
        if (c->cg_elide) {
          if (s == 0) launch_ew<EW_GV_E, PM, false>(c, g); else launch_spmv<SP_GV_E, PM, false>(c, g, nullptr, nullptr);
          break;
        }
      }
      if (s == 0) launch_ew<EW_GV, PM, false>(c, g); else launch_spmv<SP_GV, PM, false>(c, g, nullptr, nullptr);
      break;
    case CGX_PR:
      if (s == 0) launch_ew<EW_PR, PM, false>(c, g); else launch_spmv<SP_PR, PM, false>(c, g, nullptr, nullptr);
      break;
    case CGX_M:
      if (s == 0) launch_ew<EW_PR, PM, true>(c, g); else launch_spmv<SP_PR, PM, true>(c, g, nullptr, nullptr);
      break;
    case CGX_PIPE_PR:
      if (s == 0) launch_ew<EW_PIPE_R, PM, false>(c, g); else launch_spmv<SP_PIPE_R, PM, false>(c, g, nullptr, nullptr);
      break;
    case CGX_PIPE_PR_M:
      if (s == 0) launch_ew<EW_PIPE_R, PM, true>(c, g); else launch_spmv<SP_PIPE_R, PM, true>(c, g, nullptr, nullptr);
      break;
    case CGX_PIPE_P:
      if (s == 0) launch_ew<EW_PIPE_N, PM, false>(c, g); else launch_spmv<SP_PIPE_N, PM, false>(c, g, nullptr, nullptr);
      break;
    case CGX_PIPE_P_M:
      if (s == 0) launch_ew<EW_PIPE_N, PM, true>(c, g); else launch_spmv<SP_PIPE_N, PM, true>(c, g, nullptr, nullptr);
      break;
  }
}

#define CGX_CAT_(a, b) a##b
#define CGX_CAT(a, b) CGX_CAT_(a, b)
void CGX_CAT(cgx_iter_stage_pm, CGX_PM)(cgx_ctx* c, int s, const Args& g) { iter_stage_pm<CGX_PM>(c, s, g); }
