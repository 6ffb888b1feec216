// Micro-benchmark: achievable HBM bandwidth for the rollout kernel's access pattern - warps gathering random
// records (512 B or 1536 B) out of a 1.3 GB table, U records in flight per warp.   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int U, int REC4>   // REC4 = float4 per lane per record (1: 512 B, 3: 1536 B)
__global__ void gather(const float4* __restrict__ tab, size_t nrec, int iters, float* out, unsigned seed) {
  const int lane = threadIdx.x & 31;
  unsigned s = seed + (blockIdx.x * blockDim.x + threadIdx.x) / 32 * 2654435761u;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float4 v[U][REC4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s = s * 1664525u + 1013904223u;
      const size_t r = (size_t)(s >> 8) % nrec;
#pragma unroll
      for (int c = 0; c < REC4; ++c) v[u][c] = __ldg(tab + r * 96 + c * 32 + lane);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < REC4; ++c) acc += v[u][c].x + v[u][c].w;
  }
  if (acc == 123.456f) out[0] = acc;
}
template <int U, int REC4>
void run(const float4* tab, size_t nrec, float* out, int warps_per_sm) {
  const int threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
  const int iters = 4000 / U;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  gather<U, REC4><<<blocks, threads>>>(tab, nrec, iters / 4, out, 1u);
  cudaEventRecord(a);
  gather<U, REC4><<<blocks, threads>>>(tab, nrec, iters, out, 7u);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)blocks * threads / 32 * iters * U * REC4 * 512.0;
  printf("record %4d B, %2d in flight/warp, %2d warps/SM: %7.1f GB/s  (%.2f ms)\n", REC4 * 512, U, warps_per_sm, bytes / ms / 1e6, ms);
}
int main() {
  const size_t nrec = 868352;                      // 8192 x 106 node records of 1536 B = 1.33 GB
  float4* tab; float* out;
  cudaMalloc(&tab, nrec * 1536); cudaMalloc(&out, 4); cudaMemset(tab, 0, nrec * 1536);
  for (int w : {16, 32, 64}) {
    run<4, 1>(tab, nrec, out, w); run<8, 1>(tab, nrec, out, w); run<16, 1>(tab, nrec, out, w);
    run<2, 3>(tab, nrec, out, w); run<4, 3>(tab, nrec, out, w);
  }
  return 0;
}
