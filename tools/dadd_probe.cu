// Dependent-issue latency of the fp64 add / multiply / fma pipes and of a shared-memory load feeding an
// add (one warp, one CTA): cycles per operation of a 4096-long dependent chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/dadd_probe tools/dadd_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, double seed, int lanes) {
  __shared__ double sh[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = seed * i;
  __syncthreads();
  if ((int)threadIdx.x >= lanes) return;
  double a = seed, b = seed * 0.5;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 4096; ++i) a = __dadd_rn(a, b);
  long long t1 = clock64();
#pragma unroll 16
  for (int i = 0; i < 4096; ++i) a = __dmul_rn(a, b);
  long long t2 = clock64();
#pragma unroll 16
  for (int i = 0; i < 4096; ++i) a = fma(a, b, b);
  long long t3 = clock64();
  float f = (float)seed, h = f * 0.5f;
#pragma unroll 16
  for (int i = 0; i < 4096; ++i) f = __fadd_rn(f, h);
  long long t4 = clock64();
  // shared-memory load -> add, batches of 8 loads then 8 dependent adds (the row-sum pattern)
  const double* p = sh + threadIdx.x * 65 % 2048;
  for (int i = 0; i < 2048; i += 8) {
    double t[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) t[q] = p[i + q];
#pragma unroll
    for (int q = 0; q < 8; ++q) a = __dadd_rn(a, t[q]);
  }
  long long t5 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
  out[threadIdx.x] = a + f;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 8);
  for (int warps : {1, 4, 8, 16}) for (int lanes : {1, 32}) {
    const int threads = warps * 32;
    probe<<<1, threads>>>(out, cyc, 1.0000001, lanes == 1 ? 1 : threads);
    long long h[5];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("warps %2d lanes/warp %2d: cycles per dependent op: dadd %.1f  dmul %.1f  dfma %.1f  fadd %.1f | lds+dadd (per element, batches of 8) %.1f\n",
           lanes == 1 ? 1 : warps, lanes, h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0, h[3] / 4096.0, h[4] / 2048.0);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
