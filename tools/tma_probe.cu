// Standalone probe of the TMA plane copy used by cgx_stencil_tma.cuh (debug aid).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../new_cg_variants_b200/csrc/cgx_stencil_tma.cuh"
using namespace cgx;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap tm, double* out, int x0, int y0, int z) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* smem = reinterpret_cast<double*>(raw + ((128u - (smem_u32(raw) & 127u)) & 127u));
  __shared__ __align__(8) uint64_t bar[1];
  if (threadIdx.x == 0) {
    printf("smem_raw=%u aligned=%u bar=%u\n", smem_u32(raw), smem_u32(smem), smem_u32(bar));
    mbar_init(&bar[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar[0], kPlane * 8);
    tma_load_3d(smem, &tm, x0, y0, z, &bar[0]);
  }
  mbar_wait(&bar[0], 0);
  for (int i = threadIdx.x; i < kPlane; i += blockDim.x) out[i] = smem[i];
}
int main() {
  int nx = 256, ny = 64, nz = 8;
  size_t n = (size_t)nx * ny * nz;
  std::vector<double> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (double)i;
  double *d, *o;
  cudaMalloc(&d, n * 8); cudaMalloc(&o, kPlane * 8);
  cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  CUtensorMap tm;
  cuuint64_t gdim[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
  cuuint64_t gstr[2] = {(cuuint64_t)nx * 8, (cuuint64_t)nx * ny * 8};
  cuuint32_t box[3] = {(cuuint32_t)kPX, (cuuint32_t)kPY, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  for (int trial = 0; trial < 2; ++trial) {
    int x0 = trial ? -2 : 0, y0 = trial ? -1 : 8, z = trial ? 0 : 3;
    probe<<<1, 256, kPlaneStride * 8 + 128>>>(tm, o, x0, y0, z);
    cudaError_t e = cudaDeviceSynchronize();
    printf("trial %d sync: %s\n", trial, cudaGetErrorString(e));
    std::vector<double> ho(kPlane);
    cudaMemcpy(ho.data(), o, kPlane * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < kPY; ++yy) for (int xx = 0; xx < kPX; ++xx) {
      int gx = x0 + xx, gy = y0 + yy;
      double want = (gx < 0 || gx >= nx || gy < 0 || gy >= ny) ? 0.0 : (double)((size_t)z * nx * ny + (size_t)gy * nx + gx);
      if (ho[yy * kPX + xx] != want) { if (bad < 5) printf("  mismatch (%d,%d): got %g want %g\n", xx, yy, ho[yy * kPX + xx], want); ++bad; }
    }
    int flag = 0; cudaMemcpyFromSymbol(&flag, g_tma_timeout, sizeof(int));
    printf("trial %d: %d mismatches, timeout flag %d\n", trial, bad, flag);
  }
  return 0;
}
