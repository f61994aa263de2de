// cgx_pers.cu -- launcher of the persistent cooperative kernel (cgx_persistent.cuh) for one
// operator kind and one preconditioner mode (compiled six times: -DCGX_PERS_OP=1 CSR | 2 matrix-free
// stencil, -DCGX_PERS_PM=0|1|2).
#include "cgx_host.h"

#ifndef CGX_PERS_OP
#error "compile with -DCGX_PERS_OP=1|2"
#endif

template <class Op, int PM>
static int pers_launch_pm(cgx_ctx** cs, int count, const PersGeom& G, int k0, int k1) {
  cgx_ctx* c0 = cs[0];
  std::vector<PersRank<Op>> h(count);
  for (int i = 0; i < count; ++i) {
    cgx_ctx* c = cs[i];
    PersRank<Op>& r = h[i];
    if constexpr (std::is_same<Op, CsrOp>::value) r.A = c->csr; else r.A = c->sten;
    Args g = make_args(c);
    Plan p;                                         // fills scpar / x_true epochs
    plan_apply(c, g, p);
    r.g = g;
    r.g.l2pol = 0;                                  // (the persistent kernel addresses shared memory)
    for (int v = 0; v < V_COUNT; ++v) r.vecs[v] = c->vec[v];
    r.vecs[10] = c->d_dinv; r.vecs[11] = nullptr;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) r.exp_[a][b] = c->d_exp[a][b];
    r.bar = c->d_pbar; r.part = c->d_ppart; r.out = c->d_pout;
    r.epoch0 = c->epoch;
    for (int ch = 0; ch < kChan; ++ch) r.hepoch0[ch] = c->hepoch[ch];
    CU(cudaMemsetAsync(c->d_pbar, 0, sizeof(u64) * 2, c0->stream));
    CU(cudaMemsetAsync(c->d_pout, 0, sizeof(PersOut), c0->stream));
  }
  CU(cudaMemcpyAsync(c0->d_prank, h.data(), sizeof(PersRank<Op>) * count, cudaMemcpyHostToDevice, c0->stream));
  CU(cudaStreamSynchronize(c0->stream));           // h is a host temporary
  PersLaunch L{};
  L.k0 = k0; L.k1 = k1; L.nb = G.nb; L.R = G.R; L.vmask = G.vmask; L.nslot = G.nslot; L.slab_cap = G.slab_cap;
  const PersRank<Op>* dr = static_cast<const PersRank<Op>*>(c0->d_prank);
  void* params[] = {(void*)&dr, (void*)&L};
  const void* fn = nullptr;
#define CGX_PV(V) case V: fn = (const void*)persistent_kernel<Op, V, PM>; break;
  switch (c0->variant) {
    CGX_PV(CGX_HS) CGX_PV(CGX_CG) CGX_PV(CGX_GV) CGX_PV(CGX_PR) CGX_PV(CGX_M) CGX_PV(CGX_PIPE_PR)
    CGX_PV(CGX_PIPE_P) CGX_PV(CGX_PIPE_PR_M) CGX_PV(CGX_PIPE_P_M)
  }
#undef CGX_PV
  if (!fn) return fail(CGX_ERR_ARG, "persistent path: unknown variant");
  CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G.smem));
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, G.T, G.smem));
  if ((i64)per_sm * c0->sm_count < (i64)G.nb * count)
    return fail(CGX_ERR_UNSUPPORTED, "persistent path: %d CTAs of %d threads / %zu B shared memory are not co-resident",
                G.nb * count, G.T, G.smem);
  CU(cudaLaunchCooperativeKernel(fn, dim3(G.nb * count), dim3(G.T), params, G.smem, c0->stream));
  c0->launches++;
  return CGX_OK;
}

#ifndef CGX_PERS_PM
#error "compile with -DCGX_PERS_PM=0|1|2"
#endif
#define CGX_CAT_(a, b, c) a##b##c
#define CGX_CAT(a, b, c) CGX_CAT_(a, b, c)
#if CGX_PERS_OP == 1
int CGX_CAT(cgx_pers_launch_csr, _pm, CGX_PERS_PM)(cgx_ctx** cs, int count, const PersGeom& G, int k0, int k1) {
  return pers_launch_pm<CsrOp, CGX_PERS_PM>(cs, count, G, k0, k1);
}
#else
int CGX_CAT(cgx_pers_launch_sten, _pm, CGX_PERS_PM)(cgx_ctx** cs, int count, const PersGeom& G, int k0, int k1) {
  return pers_launch_pm<StencilOp, CGX_PERS_PM>(cs, count, G, k0, k1);
}
#endif
