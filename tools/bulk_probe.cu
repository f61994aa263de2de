// cp.async.bulk (global -> shared) throughput as a function of copy size, copies in flight per issuing
// warp and issuing warps per SM: every warp streams its share of a 1 GiB buffer through a private ring.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/bulk_probe tools/bulk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const unsigned char* src, size_t total, int S, int D, unsigned long long* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar[32][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  unsigned char* ring = sm + (size_t)warp * D * S;
  if (lane == 0) {
    for (int d = 0; d < D; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[warp][d])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const size_t GW = (size_t)gridDim.x * nw, gw = (size_t)blockIdx.x * nw + warp;
  const size_t ncopies = total / S;
  unsigned long long acc = 0;
  if (lane == 0) {
    size_t issued = gw, done = gw;
    int si = 0, sd = 0;
    uint32_t ph = 0;
    auto issue = [&]() {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[warp][si])), "r"(S) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(s32(ring + (size_t)si * S)), "l"(src + issued * S), "r"(S), "r"(s32(&bar[warp][si])) : "memory");
      issued += GW; if (++si == D) si = 0;
    };
    for (int d = 0; d < D && issued < ncopies; ++d) issue();
    while (done < ncopies) {
      uint32_t ok = 0;
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                               : "=r"(ok) : "r"(s32(&bar[warp][sd])), "r"(ph) : "memory");
      acc += *reinterpret_cast<unsigned long long*>(ring + (size_t)sd * S);
      done += GW;
      if (++sd == D) { sd = 0; ph ^= 1u; }
      if (issued < ncopies) issue();
    }
    sink[gw] = acc;
  }
}
int main() {
  const size_t total = 1ull << 30;
  unsigned char* src; unsigned long long* sink;
  cudaMalloc(&src, total); cudaMalloc(&sink, 8 * 148 * 64);
  cudaMemset(src, 1, total);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int S : {1024, 2048, 4096, 8192, 16384})
    for (int nw : {1, 4, 8, 16})
      for (int D : {2, 4, 8}) {
        const size_t smem = (size_t)nw * D * S;
        if (smem > 200 * 1024) continue;
        probe<<<148, nw * 32, smem>>>(src, total, S, D, sink);     // warm-up
        cudaEventRecord(e0);
        probe<<<148, nw * 32, smem>>>(src, total, S, D, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("copy %5d B  warps/SM %2d  in flight/warp %d (%3zu KB/SM): %7.1f GB/s  (%.0f ns per copy per SM)\n", S, nw, D, smem >> 10,
               total / (ms * 1e-3) / 1e9, ms * 1e6 / (double(total / S) / 148));
      }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
