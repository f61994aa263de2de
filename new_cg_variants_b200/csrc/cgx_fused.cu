// cgx_fused.cu -- host side of the single-launch PR-CG / M-CG iteration (cgx_stencil_fused.cuh).
#include "cgx_host.h"
#include "cgx_stencil_fused.cuh"

static const int kFusedVec[3] = {V_P, V_S, V_RT};

double* cgx_cur_vec(cgx_ctx* c, int v) {
  if (c->pr_fused && (c->fpar & 1))
    for (int j = 0; j < 3; ++j)
      if (kFusedVec[j] == v) return c->alt[j];
  return c->vec[v];
}

// Called from cgx_begin (after the state vectors exist and setup_tma has run).  Leaves
// c->pr_fused false -- the two-kernel path runs -- when the operator or the preconditioner is
// not one the fused kernel handles.
int cgx_fused_prepare(cgx_ctx* c) {
  c->pr_fused = false;
  c->fpar = 0;
  if (c->op_kind != 2 || !c->use_tma || c->pm == 1 || c->dist.world > 1) return CGX_OK;
  const StencilOp& S = c->sten;
  for (int j = 0; j < 3; ++j) {
    if (!c->alt[j]) CU(cudaMalloc(&c->alt[j], sizeof(double) * c->n));
    if (!tma_encode_dims(c->vec[kFusedVec[j]], S.nx, S.ny, S.nz, &c->ftmap[0][j])) return CGX_OK;
    if (!tma_encode_dims(c->alt[j], S.nx, S.ny, S.nz, &c->ftmap[1][j])) return CGX_OK;
  }
  c->pr_fused = true;
  return CGX_OK;
}

template <int PM, bool MEUR>
static void launch_fused_t(cgx_ctx* c, const Args& g, int cur) {
  const size_t smem = fused_smem_bytes();
  const int per_sm = ctx_occupancy(c, (const void*)pr_fused_kernel<PM, MEUR>, kFThreads, smem);
  // (column, z-chunk) work units, all co-resident when the columns fit: every chunk then marches
  // in lockstep over its planes (L2 serves the halos neighbouring columns share)
  TmaGeom G = c->geom;
  const int cap = per_sm * c->sm_count, ncols = G.ntx * G.nty;
  G.nchunk = std::max(1, std::min(cap / std::max(1, ncols), G.nz / std::max(1, c->fused_min_planes)));
  if (c->fused_chunks > 0) G.nchunk = std::min(c->fused_chunks, G.nz);
  const int grid = (int)std::min<i64>((i64)ncols * G.nchunk, cap);
  pr_fused_kernel<PM, MEUR><<<grid, kFThreads, smem, c->stream>>>(c->ftmap[cur][0], c->ftmap[cur][1], c->ftmap[cur][2],
                                                                  G, g);
}

void cgx_launch_pr_fused(cgx_ctx* c, Args g) {
  const int cur = c->fpar, nxt = cur ^ 1;
  g.p = nxt ? c->alt[0] : c->vec[V_P];
  g.s = nxt ? c->alt[1] : c->vec[V_S];
  g.rt = nxt ? c->alt[2] : c->vec[V_RT];
  {
    ProfScope ps(c, PC_FUSED);
    const bool meur = c->variant == CGX_M;
    if (c->pm == 2) { if (meur) launch_fused_t<2, true>(c, g, cur); else launch_fused_t<2, false>(c, g, cur); }
    else { if (meur) launch_fused_t<0, true>(c, g, cur); else launch_fused_t<0, false>(c, g, cur); }
    c->launches++;
  }
  c->fpar = nxt;
}
