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
  if (c->op_kind != 2 || !c->use_tma || c->pm == 1) return CGX_OK;
  // partitioned: the records travel peer to peer (the NCCL mode keeps the two-kernel path)
  if (c->dist.world > 1 && (c->dist.mode == 2 || !c->halo_ll)) return CGX_OK;
  // thin slabs: a march of < ~16 planes per CTA no longer amortises the pipeline fill of the fused
  // kernel (measured at 8 GPUs, 32 planes per rank: 53.4 us/iteration fused, 50.4 two kernels; at 2
  // GPUs, 128 planes: 137 vs 151) -- option "fused_min_slab" moves the switch
  if (c->dist.world > 1 && c->sten.nz < c->fused_min_slab) return CGX_OK;
  const StencilOp& S = c->sten;
  for (int j = 0; j < 3; ++j) {
    if (!c->alt[j]) CU(cudaMalloc(&c->alt[j], sizeof(double) * c->n));
    if (!tma_encode_dims(c->vec[kFusedVec[j]], S.nx, S.ny, S.nz, &c->ftmap[0][j])) return CGX_OK;
    if (!tma_encode_dims(c->alt[j], S.nx, S.ny, S.nz, &c->ftmap[1][j])) return CGX_OK;
  }
  // the LL ghost planes are shared with the other partitioned kernels (tags = epochs): keep the
  // tags unique by starting above every epoch any of them has used
  for (int ch = 0; ch < 3; ++ch) c->fepoch = std::max(c->fepoch, c->hepoch[ch]);
  if (c->dist.world > 1 && !c->d_gscr) CU(cudaMalloc(&c->d_gscr, sizeof(double) * 4 * (size_t)c->dist.plane));
  c->pr_fused = true;
  return CGX_OK;
}

template <int PM, bool MEUR, bool DIST>
static void launch_fused_t(cgx_ctx* c, const Args& g, int cur) {
  const size_t smem = fused_smem_bytes();
  const int per_sm = ctx_occupancy(c, (const void*)pr_fused_kernel<PM, MEUR, DIST>, kFThreads, smem);
  // (column, z-chunk) work units, all co-resident when the columns fit: every chunk then marches
  // in lockstep over its planes (L2 serves the halos neighbouring columns share)
  TmaGeom G = c->geom;
  const int cap = per_sm * c->sm_count, ncols = G.ntx * G.nty;
  G.nchunk = std::max(1, std::min(cap / std::max(1, ncols), G.nz / std::max(1, c->fused_min_planes)));
  if (c->fused_chunks > 0) G.nchunk = std::min(c->fused_chunks, G.nz);
  const int grid = (int)std::min<i64>((i64)ncols * G.nchunk, cap);
  launch_k(pr_fused_kernel<PM, MEUR, DIST>, grid, kFThreads, smem, c->stream, use_pdl(c), c->ftmap[cur][0], c->ftmap[cur][1],
           c->ftmap[cur][2], G, g);
}

void cgx_launch_pr_fused(cgx_ctx* c, Args g) {
  const int cur = c->fpar, nxt = cur ^ 1;
  g.p = nxt ? c->alt[0] : c->vec[V_P];
  g.s = nxt ? c->alt[1] : c->vec[V_S];
  g.rt = nxt ? c->alt[2] : c->vec[V_RT];
  const bool dist = c->dist.world > 1;
  g.gscr = c->d_gscr;
  Plan p;
  p.consume = true;
  p.produce = FK_PIPE;
  plan_apply(c, g, p);
  if (dist) {       // LL ghost planes of p, s, rt: one epoch counter for the three channels
    g.hin_epoch = c->fepoch; g.hin_par = (int)(g.hin_epoch & 1);
    g.hout_epoch = c->fepoch + 1; g.hout_par = (int)(g.hout_epoch & 1);
  }
  {
    ProfScope ps(c, PC_FUSED);
    const bool meur = c->variant == CGX_M;
#define CGX_FUSED_GO(PM, D) do { if (meur) launch_fused_t<PM, true, D>(c, g, cur); else launch_fused_t<PM, false, D>(c, g, cur); } while (0)
    if (c->pm == 2) { if (dist) CGX_FUSED_GO(2, true); else CGX_FUSED_GO(2, false); }
    else { if (dist) CGX_FUSED_GO(0, true); else CGX_FUSED_GO(0, false); }
#undef CGX_FUSED_GO
    c->launches++;
  }
  plan_commit(c, g, p);
  if (dist) c->fepoch++;
  c->fpar = nxt;
}

// partitioned runs: after the initialisation, hand the boundary planes of p, s, rt to the neighbours
void cgx_fused_push_initial_halo(cgx_ctx* c) {
  if (!c->pr_fused || c->dist.world <= 1) return;
  Args g = make_args(c);
  g.d = c->dist;
  g.hout_epoch = c->fepoch + 1; g.hout_par = (int)(g.hout_epoch & 1);
  fused_halo_init_kernel<<<grid_for(c, c->dist.plane), kBlock, 0, c->stream>>>(g);
  c->launches++;
  c->fepoch++;
}
