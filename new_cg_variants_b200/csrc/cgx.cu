// cgx.cu -- host side of libcgx_b200.so: context, device memory, the native iteration
// loop and the C ABI of include/cgx.h.  No torch, no cuBLAS/cuSPARSE, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cgx.h"
#include "cgx_kernels.cuh"
#include "cgx_stencil_tma.cuh"

using namespace cgx;

// cuTensorMapEncodeTiled is resolved through the runtime so the library has no link-time
// dependency on libcuda (it must load on the GPU-less build box for the ABI tests).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(CGX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                  \
  } while (0)

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
enum { V_X = 0, V_R, V_RT, V_P, V_S, V_ST, V_W, V_WT, V_U, V_T, V_COUNT };
static const char* kVecNames[V_COUNT] = {"x", "r", "rt", "p", "s", "st", "w", "wt", "u", "t"};

struct cgx_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  // operator
  int op_kind = 0;  // 0 none, 1 csr, 2 stencil
  CsrOp csr{};
  StencilOp sten{};
  int* d_ptr = nullptr;
  int* d_idx = nullptr;
  double* d_val = nullptr;
  i64 n = 0, nnz = 0;
  // preconditioner: pm = 0 identity, 1 Jacobi vector, 2 Jacobi with a constant diagonal
  double* d_dinv = nullptr;
  double dinv_s = 1.0;
  int pm = 0;
  // TMA-staged stencil path
  bool use_tma = false;
  bool no_tma = false;             // cgx_set_option("tma", 0): force the generic stencil kernel
  TmaGeom geom{};
  int tma_grid[2] = {0, 0};        // grid size for 1 / 2 right-hand sides
  CUtensorMap tmap[10];
  bool tmap_ok[10] = {};
  // problem
  double* d_b = nullptr;
  double* d_x0 = nullptr;
  double* d_xtrue = nullptr;
  bool own_problem = false, has_xtrue = false, problem_loaded = false;
  // state
  double* vec[V_COUNT] = {};
  Scal* d_sc = nullptr;
  double* d_partials = nullptr;
  unsigned* d_ticket = nullptr;
  double* d_hist = nullptr;
  int hist_len = 0;
  unsigned hist_mask = 0;
  bool ran = false;
  i64 launches = 0;
  // per-kernel-class profiling
  bool profile = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<int> prof_cls;
  size_t prof_used = 0;
  double prof_ms[16] = {};
  i64 prof_n[16] = {};
  // current run
  int variant = 0, max_iter = 0, cur_k = 0, path = CGX_PATH_STREAM;
  i64 launches_run = 0;
  double setup_ms = 0.0, loop_ms = 0.0;
};

static int grid_for(const cgx_ctx* c, i64 work_items) {
  i64 g = (work_items + kBlock - 1) / kBlock;
  i64 cap = (i64)c->sm_count * 8;   // 8 resident CTAs of 256 threads per SM
  if (cap > kMaxGrid) cap = kMaxGrid;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static void free_op(cgx_ctx* c) {
  cudaFree(c->d_ptr); cudaFree(c->d_idx); cudaFree(c->d_val);
  c->d_ptr = c->d_idx = nullptr; c->d_val = nullptr;
  c->op_kind = 0;
}
static void free_problem(cgx_ctx* c) {
  if (c->own_problem) { cudaFree(c->d_b); cudaFree(c->d_x0); cudaFree(c->d_xtrue); }
  c->d_b = c->d_x0 = c->d_xtrue = nullptr;
  c->own_problem = false; c->problem_loaded = false; c->has_xtrue = false;
}
static void free_state(cgx_ctx* c) {
  for (int i = 0; i < V_COUNT; ++i) { cudaFree(c->vec[i]); c->vec[i] = nullptr; }
  cudaFree(c->d_hist); c->d_hist = nullptr; c->hist_len = 0;
  c->ran = false;
}
static void reset_size(cgx_ctx* c, i64 n) {
  if (c->n != n) {
    free_state(c); free_problem(c);
    cudaFree(c->d_dinv); c->d_dinv = nullptr;
    c->n = n;
  }
}

extern "C" int cgx_version(void) { return CGX_VERSION; }
extern "C" const char* cgx_last_error(void) { return g_err.c_str(); }
extern "C" int cgx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int cgx_ctx_create(int device, cgx_ctx** out) {
  if (!out) return fail(CGX_ERR_ARG, "cgx_ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  CU(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count)
    return fail(CGX_ERR_CUDA, "cgx_ctx_create: device %d not available (%d CUDA devices); "
                "this library has no CPU fallback", device, count);
  CU(cudaSetDevice(device));
  cgx_ctx* c = new cgx_ctx();
  c->device = device;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& e : c->ev) CU(cudaEventCreate(&e));
  CU(cudaMalloc(&c->d_sc, sizeof(Scal)));
  CU(cudaMemset(c->d_sc, 0, sizeof(Scal)));
  CU(cudaMalloc(&c->d_partials, sizeof(double) * kMaxGrid * kNRed));
  CU(cudaMalloc(&c->d_ticket, sizeof(unsigned)));
  CU(cudaMemset(c->d_ticket, 0, sizeof(unsigned)));
  *out = c;
  return CGX_OK;
}

extern "C" int cgx_ctx_destroy(cgx_ctx* c) {
  if (!c) return CGX_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_state(c); free_problem(c); free_op(c);
  cudaFree(c->d_dinv); cudaFree(c->d_sc); cudaFree(c->d_partials); cudaFree(c->d_ticket);
  for (auto& e : c->ev) cudaEventDestroy(e);
  for (auto& e : c->prof_events) cudaEventDestroy(e);
  cudaStreamDestroy(c->stream);
  delete c;
  return CGX_OK;
}

// ---------------------------------------------------------------------------------------
// operator / preconditioner / problem
// ---------------------------------------------------------------------------------------
extern "C" int cgx_set_csr_host(cgx_ctx* c, int64_t n, int64_t nnz, const int32_t* indptr,
                                const int32_t* indices, const double* data) {
  if (!c || n <= 0 || nnz < 0 || !indptr || (nnz > 0 && (!indices || !data)))
    return fail(CGX_ERR_ARG, "cgx_set_csr_host: bad arguments");
  if (n >= (1ll << 31) || nnz >= (1ll << 31))
    return fail(CGX_ERR_UNSUPPORTED, "cgx_set_csr_host: int32 index range exceeded");
  if (indptr[0] != 0 || indptr[n] != nnz)
    return fail(CGX_ERR_ARG, "cgx_set_csr_host: indptr[0] != 0 or indptr[n] != nnz");
  CU(cudaSetDevice(c->device));
  free_op(c);
  reset_size(c, n);
  c->nnz = nnz;
  CU(cudaMalloc(&c->d_ptr, sizeof(int) * (n + 1)));
  CU(cudaMalloc(&c->d_idx, sizeof(int) * std::max<i64>(nnz, 1)));
  CU(cudaMalloc(&c->d_val, sizeof(double) * std::max<i64>(nnz, 1)));
  CU(cudaMemcpyAsync(c->d_ptr, indptr, sizeof(int) * (n + 1), cudaMemcpyHostToDevice, c->stream));
  if (nnz) {
    CU(cudaMemcpyAsync(c->d_idx, indices, sizeof(int) * nnz, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_val, data, sizeof(double) * nnz, cudaMemcpyHostToDevice, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  c->csr = CsrOp{c->d_ptr, c->d_idx, c->d_val, n};
  c->op_kind = 1;
  return CGX_OK;
}

extern "C" int cgx_set_stencil(cgx_ctx* c, int dim, int64_t nx, int64_t ny, int64_t nz,
                               double diag, double off) {
  if (!c || (dim != 2 && dim != 3) || nx < 1 || ny < 1 || nz < 1 || (dim == 2 && nz != 1))
    return fail(CGX_ERR_ARG, "cgx_set_stencil: bad arguments");
  const i64 n = nx * ny * nz;
  if (n + nx * ny >= (1ll << 31))
    return fail(CGX_ERR_UNSUPPORTED, "cgx_set_stencil: grid too large for int32 indexing");
  CU(cudaSetDevice(c->device));
  free_op(c);
  reset_size(c, n);
  c->sten = StencilOp{(int)nx, (int)ny, (int)nz, diag, off, n};
  c->nnz = 0;
  c->op_kind = 2;
  return CGX_OK;
}

extern "C" int cgx_set_jacobi_host(cgx_ctx* c, const double* dinv, int64_t n) {
  if (!c) return fail(CGX_ERR_ARG, "cgx_set_jacobi_host: ctx is NULL");
  CU(cudaSetDevice(c->device));
  if (!dinv) { cudaFree(c->d_dinv); c->d_dinv = nullptr; c->pm = 0; c->dinv_s = 1.0; return CGX_OK; }
  if (c->op_kind == 0 || n != c->n)
    return fail(CGX_ERR_ARG, "cgx_set_jacobi_host: set the operator first; n must match (%lld vs %lld)",
                (long long)n, (long long)c->n);
  if (!c->d_dinv) CU(cudaMalloc(&c->d_dinv, sizeof(double) * n));
  CU(cudaMemcpyAsync(c->d_dinv, dinv, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  // constant diagonal (every Poisson stencil): the same products with one HBM stream less
  bool constant = true;
  for (i64 i = 1; i < n && constant; ++i) constant = (memcmp(&dinv[i], &dinv[0], sizeof(double)) == 0);
  c->pm = constant ? 2 : 1;
  c->dinv_s = dinv[0];
  return CGX_OK;
}

static int load_problem(cgx_ctx* c, const double* b, const double* x0, const double* xt, i64 n,
                        cudaMemcpyKind kind) {
  if (!c || !b || !x0) return fail(CGX_ERR_ARG, "cgx_load_problem: b and x0 are required");
  if (c->op_kind == 0 || n != c->n)
    return fail(CGX_ERR_ARG, "cgx_load_problem: set the operator first; n must match (%lld vs %lld)",
                (long long)n, (long long)c->n);
  CU(cudaSetDevice(c->device));
  if (!c->own_problem) {
    c->d_b = c->d_x0 = c->d_xtrue = nullptr;
    CU(cudaMalloc(&c->d_b, sizeof(double) * n));
    CU(cudaMalloc(&c->d_x0, sizeof(double) * n));
    CU(cudaMalloc(&c->d_xtrue, sizeof(double) * n));
    c->own_problem = true;
  }
  CU(cudaMemcpyAsync(c->d_b, b, sizeof(double) * n, kind, c->stream));
  CU(cudaMemcpyAsync(c->d_x0, x0, sizeof(double) * n, kind, c->stream));
  if (xt) CU(cudaMemcpyAsync(c->d_xtrue, xt, sizeof(double) * n, kind, c->stream));
  c->has_xtrue = xt != nullptr;
  c->problem_loaded = true;
  return CGX_OK;
}

extern "C" int cgx_load_problem_host(cgx_ctx* c, const double* b, const double* x0,
                                     const double* xt, int64_t n) {
  int rc = load_problem(c, b, x0, xt, n, cudaMemcpyHostToDevice);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->stream));   // pageable host buffers may be reused by the caller
  return CGX_OK;
}
extern "C" int cgx_load_problem_dev(cgx_ctx* c, const double* b, const double* x0,
                                    const double* xt, int64_t n) {
  return load_problem(c, b, x0, xt, n, cudaMemcpyDeviceToDevice);
}

// ---------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------
// Optional per-kernel-class timing (cgx_set_profile): an event pair around every launch of
// the iteration loop, resolved after the stream has drained.  Off in timed runs.
enum { PC_EW0 = 0, PC_SP0 = 7, PC_INSTR = 15, PC_COUNT = 16 };
static const char* kClassNames[PC_COUNT] = {
    "ew_hs1", "ew_hs2", "ew_cg", "ew_gv", "ew_pr", "ew_pipe_r", "ew_pipe_n",
    "sp_plain", "sp_hs", "sp_cg", "sp_gv", "sp_pr", "sp_pipe_r", "sp_pipe_n", "sp_resid",
    "instrument"};

struct ProfScope {
  cgx_ctx* c; int cls; size_t slot = 0; bool on;
  ProfScope(cgx_ctx* c_, int cls_) : c(c_), cls(cls_), on(c_->profile) {
    if (!on) return;
    if (c->prof_used + 2 > c->prof_events.size()) {
      for (int i = 0; i < 2; ++i) { cudaEvent_t e; cudaEventCreate(&e); c->prof_events.push_back(e); }
    }
    slot = c->prof_used; c->prof_used += 2;
    c->prof_cls.push_back(cls);
    cudaEventRecord(c->prof_events[slot], c->stream);
  }
  ~ProfScope() { if (on) cudaEventRecord(c->prof_events[slot + 1], c->stream); }
};
static void prof_resolve(cgx_ctx* c) {
  for (size_t i = 0; i < c->prof_cls.size(); ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->prof_events[2 * i], c->prof_events[2 * i + 1]) == cudaSuccess) {
      c->prof_ms[c->prof_cls[i]] += ms; c->prof_n[c->prof_cls[i]]++;
    }
  }
  c->prof_cls.clear(); c->prof_used = 0;
}

// which state vector a fused SpMV pass reads through TMA (second one for the 2-RHS pass)
template <int MODE> struct SpInput { static constexpr int v0 = -1, v1 = -1; };
template <> struct SpInput<SP_HS> { static constexpr int v0 = 3 /*V_P*/, v1 = -1; };
template <> struct SpInput<SP_PR> { static constexpr int v0 = 3 /*V_P*/, v1 = -1; };
template <> struct SpInput<SP_CG> { static constexpr int v0 = 2 /*V_RT*/, v1 = -1; };
template <> struct SpInput<SP_GV> { static constexpr int v0 = 7 /*V_WT*/, v1 = -1; };
template <> struct SpInput<SP_PIPE_R> { static constexpr int v0 = 5 /*V_ST*/, v1 = 2 /*V_RT*/; };
template <> struct SpInput<SP_PIPE_N> { static constexpr int v0 = 5 /*V_ST*/, v1 = -1; };

static size_t tma_smem_bytes(int nv) { return (size_t)kRing * nv * kPlaneStride * sizeof(double) + 128; }

template <int MODE, int PM, bool MEUR>
static void launch_spmv(cgx_ctx* c, const Args& g, const double* vin, double* vout) {
  ProfScope ps(c, PC_SP0 + MODE);
  constexpr int v0 = SpInput<MODE>::v0, v1 = SpInput<MODE>::v1;
  if constexpr (v0 >= 0) {
    if (c->op_kind == 2 && c->use_tma && c->tmap_ok[v0] && (v1 < 0 || c->tmap_ok[v1 < 0 ? 0 : v1])) {
      constexpr int nv = (v1 >= 0) ? 2 : 1;
      static bool attr_set = false;
      if (!attr_set) {
        cudaFuncSetAttribute(stencil_tma_kernel<MODE, PM, MEUR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)tma_smem_bytes(nv));
        attr_set = true;
      }
      stencil_tma_kernel<MODE, PM, MEUR><<<c->tma_grid[nv - 1], kTmaThreads, tma_smem_bytes(nv), c->stream>>>(
          c->tmap[v0], c->tmap[v1 < 0 ? v0 : v1], c->geom, g);
      c->launches++;
      return;
    }
  }
  const int grid = grid_for(c, c->n);
  if (c->op_kind == 1)
    spmv_kernel<CsrOp, MODE, PM, MEUR><<<grid, kBlock, 0, c->stream>>>(c->csr, g, vin, vout);
  else
    spmv_kernel<StencilOp, MODE, PM, MEUR><<<grid, kBlock, 0, c->stream>>>(c->sten, g, vin, vout);
  c->launches++;
}
template <int KID, int PM, bool MEUR>
static void launch_ew(cgx_ctx* c, const Args& g) {
  const int grid = grid_for(c, (c->n + 1) / 2);
  ProfScope ps(c, PC_EW0 + KID);
  ew_kernel<KID, PM, MEUR><<<grid, kBlock, 0, c->stream>>>(g);
  c->launches++;
}

// ---- TMA stencil path: descriptors and work decomposition -------------------------------
static bool tma_encode(cgx_ctx* c, double* ptr, CUtensorMap* out) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc || !ptr) return false;
  const StencilOp& S = c->sten;
  cuuint64_t gdim[3] = {(cuuint64_t)S.nx, (cuuint64_t)S.ny, (cuuint64_t)S.nz};
  cuuint64_t gstr[2] = {(cuuint64_t)S.nx * 8, (cuuint64_t)S.nx * S.ny * 8};
  cuuint32_t box[3] = {(cuuint32_t)kPX, (cuuint32_t)kPY, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, ptr, gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool tma_prepare_geom(cgx_ctx* c) {
  if (c->op_kind != 2 || c->no_tma) return false;
  const StencilOp& S = c->sten;
  if (S.nx % 2 != 0 || S.nx < 2) return false;         // TMA needs 16-byte global strides
  if (!get_encode_tiled()) return false;
  TmaGeom& G = c->geom;
  G.nx = S.nx; G.ny = S.ny; G.nz = S.nz;
  G.ntx = (S.nx + kTX - 1) / kTX; G.nty = (S.ny + kTY - 1) / kTY;
  G.zoff = 0; G.has_zlo = 0; G.has_zhi = 0;
  G.diag = S.diag; G.off = S.off;
  // resident CTAs: shared memory bound (227 KB/SM), 8 x 256 threads at most
  const int cols = G.ntx * G.nty;
  int per_sm1 = std::min(8, (int)(227 * 1024 / (tma_smem_bytes(1) + 1024)));
  int per_sm2 = std::min(8, (int)(227 * 1024 / (tma_smem_bytes(2) + 1024)));
  const int cap1 = c->sm_count * per_sm1, cap2 = c->sm_count * per_sm2;
  // one CTA per resident slot; the kernel cuts the (column, plane) sequence evenly between
  // them (>= 4 planes per CTA when the problem is large enough to keep the z-halo small)
  G.lz = 0; G.nchunks = 0;
  const i64 total = (i64)cols * S.nz;
  const i64 want = std::max<i64>(1, (total + 3) / 4);
  c->tma_grid[0] = (int)std::min<i64>(want, cap1);
  c->tma_grid[1] = (int)std::min<i64>(want, cap2);
  return true;
}

static int setup_tma(cgx_ctx* c, unsigned need) {
  c->use_tma = false;
  for (auto& ok : c->tmap_ok) ok = false;
  if (!tma_prepare_geom(c)) return CGX_OK;
  for (int i = 0; i < V_COUNT; ++i)
    if ((need & (1u << i)) && c->vec[i]) c->tmap_ok[i] = tma_encode(c, c->vec[i], &c->tmap[i]);
  c->use_tma = true;
  return CGX_OK;
}

static void launch_instrument(cgx_ctx* c, const Args& g) {
  const int grid = grid_for(c, c->n);
  ProfScope ps(c, PC_INSTR);
  if (c->op_kind == 1) {
    if (c->has_xtrue) instrument_kernel<CsrOp, true><<<grid, kBlock, 0, c->stream>>>(c->csr, g);
    else instrument_kernel<CsrOp, false><<<grid, kBlock, 0, c->stream>>>(c->csr, g);
  } else {
    if (c->has_xtrue) instrument_kernel<StencilOp, true><<<grid, kBlock, 0, c->stream>>>(c->sten, g);
    else instrument_kernel<StencilOp, false><<<grid, kBlock, 0, c->stream>>>(c->sten, g);
  }
  c->launches++;
}
static void launch_dot(cgx_ctx* c, const double* u, const double* v, const double* dinv, int slot) {
  dot_kernel<<<grid_for(c, c->n), kBlock, 0, c->stream>>>(u, v, dinv, c->n, c->d_sc, slot,
                                                          c->d_partials, c->d_ticket);
  c->launches++;
}
static void launch_scale(cgx_ctx* c, const double* dinv, const double* v, double* out) {
  scale_kernel<<<grid_for(c, c->n), kBlock, 0, c->stream>>>(dinv, v, out, c->n);
  c->launches++;
}

struct VariantInfo {
  bool meurant, pipe, recompute;
  int cls;                 // init_scalars class
  unsigned need;           // bitmask of state vectors
};
static VariantInfo variant_info(int v, bool prec) {
  auto bit = [](int i) { return 1u << i; };
  const unsigned base = bit(V_X) | bit(V_R) | bit(V_RT) | bit(V_P) | bit(V_S);
  switch (v) {
    case CGX_HS: return {false, false, false, 0, base};
    case CGX_CG: return {false, false, false, 1, base | bit(V_W)};
    case CGX_GV: return {false, false, false, 1, base | bit(V_W) | bit(V_WT) | bit(V_ST) | bit(V_U) | bit(V_T)};
    case CGX_PR: return {false, false, false, 2, base};
    case CGX_M: return {true, false, false, 2, base};
    case CGX_PIPE_PR: return {false, true, true, 2, base | bit(V_ST) | bit(V_W) | bit(V_U)};
    case CGX_PIPE_PR_M: return {true, true, true, 2, base | bit(V_ST) | bit(V_W) | bit(V_U)};
    case CGX_PIPE_P: return {false, true, false, 2, base | bit(V_ST) | bit(V_W) | bit(V_U) | (prec ? bit(V_WT) : 0u)};
    case CGX_PIPE_P_M: return {true, true, false, 2, base | bit(V_ST) | bit(V_W) | bit(V_U) | (prec ? bit(V_WT) : 0u)};
  }
  return {false, false, false, -1, 0};
}

// one iteration of the streaming path --------------------------------------------------
template <int PREC>
static void iterate_stream(cgx_ctx* c, int variant, const VariantInfo& vi, Args& g) {
  switch (variant) {
    case CGX_HS:
      launch_ew<EW_HS1, PREC, false>(c, g);
      launch_ew<EW_HS2, PREC, false>(c, g);
      launch_spmv<SP_HS, PREC, false>(c, g, nullptr, nullptr);
      break;
    case CGX_CG:
      launch_ew<EW_CG, PREC, false>(c, g);
      launch_spmv<SP_CG, PREC, false>(c, g, nullptr, nullptr);
      break;
    case CGX_GV:
      launch_ew<EW_GV, PREC, false>(c, g);
      launch_spmv<SP_GV, PREC, false>(c, g, nullptr, nullptr);
      break;
    case CGX_PR:
      launch_ew<EW_PR, PREC, false>(c, g);
      launch_spmv<SP_PR, PREC, false>(c, g, nullptr, nullptr);
      break;
    case CGX_M:
      launch_ew<EW_PR, PREC, true>(c, g);
      launch_spmv<SP_PR, PREC, true>(c, g, nullptr, nullptr);
      break;
    case CGX_PIPE_PR:
      launch_ew<EW_PIPE_R, PREC, false>(c, g);
      launch_spmv<SP_PIPE_R, PREC, false>(c, g, nullptr, nullptr);
      break;
    case CGX_PIPE_PR_M:
      launch_ew<EW_PIPE_R, PREC, true>(c, g);
      launch_spmv<SP_PIPE_R, PREC, true>(c, g, nullptr, nullptr);
      break;
    case CGX_PIPE_P:
      launch_ew<EW_PIPE_N, PREC, false>(c, g);
      launch_spmv<SP_PIPE_N, PREC, false>(c, g, nullptr, nullptr);
      break;
    case CGX_PIPE_P_M:
      launch_ew<EW_PIPE_N, PREC, true>(c, g);
      launch_spmv<SP_PIPE_N, PREC, true>(c, g, nullptr, nullptr);
      break;
  }
  (void)vi;
}

// initial state: hs_cg.py:83-94, cg_cg.py:90-104, gv_cg.py:105-121, pr_cg.py:106-120,
// pipe_pr_cg.py:122-140.  Built from plain kernels; not part of the timed loop.
static int init_state(cgx_ctx* c, int variant, const VariantInfo& vi, Args& g) {
  const size_t bytes = sizeof(double) * c->n;
  const double* dinv = c->d_dinv;
  double** v = c->vec;
  CU(cudaMemcpyAsync(v[V_X], c->d_x0, bytes, cudaMemcpyDeviceToDevice, c->stream));
  launch_spmv<SP_RESID, 0, false>(c, g, v[V_X], v[V_R]);            // r = b - A x0
  launch_scale(c, dinv, v[V_R], v[V_RT]);                               // rt = M r
  CU(cudaMemcpyAsync(v[V_P], v[V_RT], bytes, cudaMemcpyDeviceToDevice, c->stream));  // p = rt
  launch_dot(c, v[V_R], v[V_RT], nullptr, 0);                           // nu = r.rt
  if (variant == CGX_HS || variant == CGX_PR || variant == CGX_M || vi.pipe) {
    launch_spmv<SP_PLAIN, 0, false>(c, g, v[V_P], v[V_S]);          // s = A p
    launch_dot(c, v[V_P], v[V_S], nullptr, 1);                          // mu = p.s
  }
  if (variant == CGX_CG || variant == CGX_GV) {
    launch_spmv<SP_PLAIN, 0, false>(c, g, v[V_RT], v[V_W]);         // w = A rt
    CU(cudaMemcpyAsync(v[V_S], v[V_W], bytes, cudaMemcpyDeviceToDevice, c->stream));  // s = A p = w
    launch_dot(c, v[V_P], v[V_S], nullptr, 1);                          // mu = p.s
    launch_dot(c, v[V_W], v[V_RT], nullptr, 2);                         // eta = w.rt
  }
  if (variant == CGX_GV) {
    launch_scale(c, dinv, v[V_W], v[V_WT]);                             // wt = M w
    CU(cudaMemcpyAsync(v[V_ST], v[V_WT], bytes, cudaMemcpyDeviceToDevice, c->stream));
    launch_spmv<SP_PLAIN, 0, false>(c, g, v[V_WT], v[V_T]);         // t = A wt
    CU(cudaMemcpyAsync(v[V_U], v[V_T], bytes, cudaMemcpyDeviceToDevice, c->stream));  // u = A wt
  }
  if (vi.cls == 2) {
    launch_dot(c, v[V_R], v[V_S], dinv, 3);                             // delta = r.(M s)
    launch_dot(c, v[V_S], v[V_S], dinv, 4);                             // gamma = (M s).s
  }
  if (vi.pipe) {
    launch_scale(c, dinv, v[V_S], v[V_ST]);                             // st = M s
    CU(cudaMemcpyAsync(v[V_W], v[V_S], bytes, cudaMemcpyDeviceToDevice, c->stream));   // w = s
    if (v[V_WT]) CU(cudaMemcpyAsync(v[V_WT], v[V_ST], bytes, cudaMemcpyDeviceToDevice, c->stream));
    launch_spmv<SP_PLAIN, 0, false>(c, g, v[V_ST], v[V_U]);         // u = A st
  }
  init_scalars_kernel<<<1, 1, 0, c->stream>>>(c->d_sc, vi.cls, vi.meurant ? 1 : 0);
  c->launches++;
  return CGX_OK;
}

static Args make_args(cgx_ctx* c) {
  Args g{};
  g.x = c->vec[V_X]; g.r = c->vec[V_R]; g.rt = c->vec[V_RT]; g.p = c->vec[V_P];
  g.s = c->vec[V_S]; g.st = c->vec[V_ST]; g.w = c->vec[V_W]; g.wt = c->vec[V_WT];
  g.u = c->vec[V_U]; g.t = c->vec[V_T];
  g.dinv = c->d_dinv; g.dinv_s = c->dinv_s; g.b = c->d_b; g.xtrue = c->d_xtrue;
  g.sc = c->d_sc; g.partials = c->d_partials; g.ticket = c->d_ticket;
  g.hist = c->d_hist; g.hist_len = c->hist_len; g.hist_mask = c->hist_mask;
  g.n = c->n; g.k = c->cur_k;
  return g;
}

extern "C" int cgx_begin(cgx_ctx* c, int variant, int max_iter, unsigned hist_mask, int path) {
  if (!c) return fail(CGX_ERR_ARG, "cgx_begin: ctx is NULL");
  if (variant < 0 || variant >= CGX_NUM_VARIANTS) return fail(CGX_ERR_ARG, "cgx_begin: unknown variant %d", variant);
  if (max_iter < 1) return fail(CGX_ERR_ARG, "cgx_begin: max_iter must be >= 1");
  if (c->op_kind == 0 || !c->problem_loaded) return fail(CGX_ERR_ARG, "cgx_begin: operator and problem must be set first");
  if (path == CGX_PATH_PERSISTENT)
    return fail(CGX_ERR_UNSUPPORTED, "cgx_begin: persistent path not built in this version");
  CU(cudaSetDevice(c->device));
  const bool prec = c->d_dinv != nullptr;
  const VariantInfo vi = variant_info(variant, prec);
  hist_mask &= CGX_HIST_ALL;
  if (!c->has_xtrue) hist_mask &= ~(CGX_HIST_ERROR_A_NORM | CGX_HIST_ERROR_2_NORM);

  // state vectors (allocated on demand, kept across runs)
  for (int i = 0; i < V_COUNT; ++i)
    if ((vi.need & (1u << i)) && !c->vec[i]) CU(cudaMalloc(&c->vec[i], sizeof(double) * c->n));
  if (c->hist_len != max_iter) {
    cudaFree(c->d_hist); c->d_hist = nullptr;
    CU(cudaMalloc(&c->d_hist, sizeof(double) * CGX_HIST_ROWS * (size_t)max_iter));
    c->hist_len = max_iter;
  }
  CU(cudaMemsetAsync(c->d_hist, 0, sizeof(double) * CGX_HIST_ROWS * (size_t)max_iter, c->stream));
  c->hist_mask = hist_mask;
  { int trc = setup_tma(c, vi.need); if (trc) return trc; }
  c->variant = variant; c->max_iter = max_iter; c->cur_k = 0;
  c->path = CGX_PATH_STREAM;
  c->launches_run = 0; c->loop_ms = 0.0;

  Args g = make_args(c);
  const i64 launches0 = c->launches;
  CU(cudaEventRecord(c->ev[0], c->stream));
  int rc = init_state(c, variant, vi, g);
  if (rc) return rc;
  if (hist_mask) launch_instrument(c, g);
  CU(cudaEventRecord(c->ev[1], c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  float ms0 = 0.f;
  CU(cudaEventElapsedTime(&ms0, c->ev[0], c->ev[1]));
  c->setup_ms = ms0;
  if (c->profile) { c->prof_cls.clear(); c->prof_used = 0; }   // initialisation is not profiled
  c->launches_run = c->launches - launches0;
  c->ran = true;
  return CGX_OK;
}

extern "C" int cgx_advance(cgx_ctx* c, int niter) {
  if (!c || !c->ran) return fail(CGX_ERR_ARG, "cgx_advance: call cgx_begin first");
  if (niter < 0) return fail(CGX_ERR_ARG, "cgx_advance: niter must be >= 0");
  CU(cudaSetDevice(c->device));
  const bool prec = c->d_dinv != nullptr;
  const VariantInfo vi = variant_info(c->variant, prec);
  Args g = make_args(c);
  const int last = std::min(c->max_iter - 1, c->cur_k + niter);
  const i64 launches0 = c->launches;
  CU(cudaEventRecord(c->ev[1], c->stream));
  for (int k = c->cur_k + 1; k <= last; ++k) {
    g.k = k;
    if (c->pm == 2) iterate_stream<2>(c, c->variant, vi, g);
    else if (c->pm == 1) iterate_stream<1>(c, c->variant, vi, g);
    else iterate_stream<0>(c, c->variant, vi, g);
    if (c->hist_mask) launch_instrument(c, g);
  }
  CU(cudaEventRecord(c->ev[2], c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  float ms1 = 0.f;
  CU(cudaEventElapsedTime(&ms1, c->ev[1], c->ev[2]));
  c->loop_ms += ms1;
  if (c->profile) prof_resolve(c);
  if (c->use_tma) {
    int flag = 0;
    CU(cudaMemcpyFromSymbol(&flag, g_tma_timeout, sizeof(int)));
    if (flag) return fail(CGX_ERR_CUDA, "cgx_advance: a TMA plane copy did not complete within 1 s");
  }
  c->launches_run += c->launches - launches0;
  c->cur_k = std::max(c->cur_k, last);
  return CGX_OK;
}

extern "C" int cgx_get_info(cgx_ctx* c, cgx_info* info) {
  if (!c || !c->ran || !info) return fail(CGX_ERR_ARG, "cgx_get_info: bad arguments");
  CU(cudaSetDevice(c->device));
  Scal h;
  CU(cudaMemcpy(&h, c->d_sc, sizeof(Scal), cudaMemcpyDeviceToHost));
  info->setup_ms = c->setup_ms; info->loop_ms = c->loop_ms;
  info->h2d_bytes = 0; info->d2h_bytes = 0;
  info->kernel_launches = c->launches_run;
  info->iterations = c->cur_k;
  info->breakdown_iter = h.breakdown;
  info->path = c->path;
  info->reserved = 0;
  return CGX_OK;
}

extern "C" int cgx_set_option(cgx_ctx* c, const char* name, int value) {
  if (!c || !name) return fail(CGX_ERR_ARG, "cgx_set_option: bad arguments");
  if (!strcmp(name, "tma")) { c->no_tma = (value == 0); return CGX_OK; }
  return fail(CGX_ERR_ARG, "cgx_set_option: unknown option '%s'", name);
}

extern "C" int cgx_set_profile(cgx_ctx* c, int on) {
  if (!c) return fail(CGX_ERR_ARG, "cgx_set_profile: ctx is NULL");
  c->profile = on != 0;
  for (int i = 0; i < PC_COUNT; ++i) { c->prof_ms[i] = 0.0; c->prof_n[i] = 0; }
  c->prof_cls.clear(); c->prof_used = 0;
  return CGX_OK;
}
extern "C" int cgx_get_profile(cgx_ctx* c, int cls, double* ms, int64_t* launches) {
  if (!c || cls < 0 || cls >= PC_COUNT) return fail(CGX_ERR_ARG, "cgx_get_profile: bad arguments");
  if (ms) *ms = c->prof_ms[cls];
  if (launches) *launches = c->prof_n[cls];
  return CGX_OK;
}
extern "C" const char* cgx_profile_class_name(int cls) {
  return (cls >= 0 && cls < PC_COUNT) ? kClassNames[cls] : "";
}
extern "C" int cgx_profile_class_count(void) { return PC_COUNT; }

// scalars of the recurrences after the last completed iteration:
// out[0..8] = a_k, a_{k-1}, b_k, nu_k, nu_{k-1}, mu_k, eta_k, delta_k, gamma_k
extern "C" int cgx_get_scalars(cgx_ctx* c, double* out9) {
  if (!c || !c->ran || !out9) return fail(CGX_ERR_ARG, "cgx_get_scalars: bad arguments");
  CU(cudaSetDevice(c->device));
  Scal h;
  CU(cudaMemcpy(&h, c->d_sc, sizeof(Scal), cudaMemcpyDeviceToHost));
  const double v[9] = {h.a, h.a1, h.b, h.nu, h.nu1, h.mu, h.eta, h.del, h.gam};
  memcpy(out9, v, sizeof v);
  return CGX_OK;
}

extern "C" int cgx_run(cgx_ctx* c, int variant, int max_iter, unsigned hist_mask, int path,
                       cgx_info* info) {
  int rc = cgx_begin(c, variant, max_iter, hist_mask, path);
  if (rc) return rc;
  rc = cgx_advance(c, max_iter - 1);
  if (rc) return rc;
  cgx_info tmp;
  rc = cgx_get_info(c, &tmp);
  if (rc) return rc;
  if (info) *info = tmp;
  return tmp.breakdown_iter >= 0 ? CGX_ERR_BREAKDOWN : CGX_OK;
}

static int fetch(cgx_ctx* c, double* x, double* hist, cudaMemcpyKind kind) {
  if (!c || !c->ran) return fail(CGX_ERR_ARG, "cgx_fetch: nothing has been run");
  CU(cudaSetDevice(c->device));
  if (x) CU(cudaMemcpyAsync(x, c->vec[V_X], sizeof(double) * c->n, kind, c->stream));
  if (hist) CU(cudaMemcpyAsync(hist, c->d_hist, sizeof(double) * CGX_HIST_ROWS * (size_t)c->hist_len, kind, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return CGX_OK;
}
extern "C" int cgx_fetch_host(cgx_ctx* c, double* x, double* hist) { return fetch(c, x, hist, cudaMemcpyDeviceToHost); }
extern "C" int cgx_fetch_dev(cgx_ctx* c, double* x, double* hist) { return fetch(c, x, hist, cudaMemcpyDeviceToDevice); }

extern "C" int cgx_fetch_vector_host(cgx_ctx* c, const char* name, double* out) {
  if (!c || !c->ran || !name || !out) return fail(CGX_ERR_ARG, "cgx_fetch_vector_host: bad arguments");
  for (int i = 0; i < V_COUNT; ++i)
    if (!strcmp(name, kVecNames[i])) {
      if (!c->vec[i]) return fail(CGX_ERR_ARG, "cgx_fetch_vector_host: vector '%s' is not part of the last variant's state", name);
      CU(cudaSetDevice(c->device));
      CU(cudaMemcpy(out, c->vec[i], sizeof(double) * c->n, cudaMemcpyDeviceToHost));
      return CGX_OK;
    }
  return fail(CGX_ERR_ARG, "cgx_fetch_vector_host: unknown vector '%s'", name);
}

extern "C" int cgx_solve_host(cgx_ctx* c, int variant, const double* b, const double* x0,
                              const double* xt, int64_t n, int max_iter, unsigned hist_mask,
                              int path, double* x, double* hist, cgx_info* info) {
  int rc = load_problem(c, b, x0, xt, n, cudaMemcpyHostToDevice);
  if (rc) return rc;
  int rrc = cgx_run(c, variant, max_iter, hist_mask, path, info);
  if (rrc != CGX_OK && rrc != CGX_ERR_BREAKDOWN) return rrc;
  rc = cgx_fetch_host(c, x, hist);
  if (rc) return rc;
  if (info) {
    info->h2d_bytes = (double)sizeof(double) * n * (xt ? 3 : 2);
    info->d2h_bytes = (double)sizeof(double) * ((x ? n : 0) + (hist ? (i64)CGX_HIST_ROWS * max_iter : 0));
  }
  return rrc;
}

// ---------------------------------------------------------------------------------------
// primitives for unit tests
// ---------------------------------------------------------------------------------------
extern "C" int cgx_spmv_host(cgx_ctx* c, const double* v, double* y, int64_t n) {
  if (!c || !v || !y || c->op_kind == 0 || n != c->n) return fail(CGX_ERR_ARG, "cgx_spmv_host: bad arguments");
  CU(cudaSetDevice(c->device));
  double *dv = nullptr, *dy = nullptr;
  CU(cudaMalloc(&dv, sizeof(double) * n));
  CU(cudaMalloc(&dy, sizeof(double) * n));
  CU(cudaMemcpyAsync(dv, v, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  Args g{};
  g.n = n;
  launch_spmv<SP_PLAIN, 0, false>(c, g, dv, dy);
  CU(cudaMemcpyAsync(y, dy, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  cudaFree(dv); cudaFree(dy);
  return CGX_OK;
}

extern "C" int cgx_dot_host(cgx_ctx* c, const double* u, const double* v, int64_t n, double* out) {
  if (!c || !u || !v || !out || n < 1) return fail(CGX_ERR_ARG, "cgx_dot_host: bad arguments");
  CU(cudaSetDevice(c->device));
  double *du = nullptr, *dv = nullptr;
  CU(cudaMalloc(&du, sizeof(double) * n));
  CU(cudaMalloc(&dv, sizeof(double) * n));
  CU(cudaMemcpyAsync(du, u, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(dv, v, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  const i64 keep = c->n;
  c->n = n;
  launch_dot(c, du, dv, nullptr, 7);
  c->n = keep;
  Scal h;
  CU(cudaMemcpyAsync(&h, c->d_sc, sizeof(Scal), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  *out = h.tmp[7];
  cudaFree(du); cudaFree(dv);
  return CGX_OK;
}
