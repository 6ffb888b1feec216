// Micro-benchmark (B200): tensor-memory read bandwidth (tcgen05.ld) and tcgen05.mma issue rates, alone and together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/tmem_bench tools/tmem_bench.cu -lcuda
// mode bits: 1 = 16 reader warps, 2 = 4 reader warps (one per SM sub-partition), 4 = MMA warp (A from TMEM, N=128),
//            8 = MMA with A from shared memory, 16 = MMA N=16 (A from TMEM), 32 = MMA N=256 (A from TMEM),
//            64 = MMA N=64, A from shared memory, ONE accumulator (the FF_tour shape of the rollout kernel),
//            128 = the same with the 4 MMAs of a round going to 4 different accumulators
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t acc, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t acc, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(576, 1) bench(int mode, int iters, long long* out, uint32_t* sink) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  const long long t0 = clock64();
  if (warp < 16) {
    const bool reader = (mode & 1) || ((mode & 2) && warp < 4);
    if (reader) {
      uint32_t acc = 0;
      const uint32_t ta = tb + ((uint32_t)((warp & 3) * 32) << 16) + 128 + (warp >> 2) * 32;
      for (int i = 0; i < iters; ++i) {
        uint32_t v[32];
        tmem_ld32(ta, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) acc ^= v[e];
      }
      if (acc == 0x12345678u) sink[threadIdx.x] = acc;
      if (threadIdx.x == 0) out[0] = clock64() - t0;
    }
  } else if (warp == 16 && (mode & (4 | 8 | 16 | 32 | 64 | 128))) {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    const int n = (mode & 16) ? 16 : (mode & 32) ? 256 : (mode & (64 | 128)) ? 64 : 128;
    const uint32_t idesc = idesc_f16(128, n);
    const uint64_t db = umma_desc(smem_u32(smem)), da = umma_desc(smem_u32(smem) + 32768);
    const uint32_t d = tb + 256, a = tb;           // D columns 256.. (up to 256 wide), A planes at columns 0..
    if (pred) {
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (mode & 64) mma_ss(d, da + 2 * k, db + 2 * k, 1u, idesc);
          else if (mode & 128) mma_ss(d + 64 * k, da + 2 * k, db + 2 * k, 1u, idesc);
          else if (mode & 8) mma_ss(d, da + 2 * k, db + 2 * k, 1u, idesc);
          else mma_ts(d, a + 8 * k, db + 2 * k, 1u, idesc);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      out[1] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 16); cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  const int modes[] = {1, 2, 4, 8, 16, 32, 64, 128, 1 | 4, 2 | 4, 1 | 8, 1 | 16, 1 | 32};
  for (int mode : modes) {
    long long h[2] = {0, 0};
    cudaMemset(out, 0, 16);
    bench<<<1, 576, 200 * 1024>>>(mode, iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("mode %2d: reader %8.1f cycles per 32x32b.x32 load round (per warp)   mma %7.1f cycles per MMA\n", mode,
           (double)h[0] / iters, (double)h[1] / (iters * 4.0));
  }
  return 0;
}
