// tcgen05 / TMEM / TMA GEMM for the encoder's dense contractions, fp32-faithful.
//
//   C[M,N] = epilogue( A[M,K] @ W[N,K]^T )      A: fp32 activations, row-major;  W: nn.Linear layout [out,in]
//
// Precision: greedy tours flip under TF32-level noise, so every operand is split into two
// TF32 planes with round-to-nearest,  x = hi + lo  (|x - hi - lo| <= 2^-24 |x|), and the product is evaluated as
//   A W^T ~= Ah Wh^T + Ah Wl^T + Al Wh^T            (3 x kind::tf32 MMAs, fp32 accumulation in TMEM)
// The dropped Al Wl^T term is <= 2^-22 relative, i.e. fp32-rounding level.
//
// Kernel: persistent, one CTA per SM, warp-specialised (320 threads):
//   warps 0-3  epilogue      tcgen05.ld TMEM -> registers -> bias / ReLU / residual / folded-BatchNorm -> global
//   warps 4-7  A converters  global fp32 (coalesced float4) -> cvt.rna.tf32 hi/lo -> 128B-swizzled K-major smem
//   warp  8    MMA issuer    one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M=128,N=128,K=8)
//   warp  9    W producer    TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) of the pre-split weight planes
// A CTA owns one 128-column chunk of N and strides over 128-row M tiles.  When K == 128 the whole weight chunk
// (hi+lo, 128 KB) stays resident in shared memory for the CTA's life and only A streams (3-stage ring of 32-column
// K blocks); for K == 512 (FF2) weights stream through the same ring.  Two 128-column TMEM accumulators let the
// epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include "evrp_tc.cuh"

namespace evrp {
namespace tc {

constexpr int BM = 128, BN = 128, BKB = 32;          // tile; BKB fp32 = one 128-byte swizzle row
constexpr int MAX_STAGES = 4;
#ifndef EVRP_GEMM_STAGES
#define EVRP_GEMM_STAGES 4
#endif
#ifndef EVRP_GEMM_RAW_STAGES
#define EVRP_GEMM_RAW_STAGES 6
#endif
constexpr int PLANE = BM * BKB * 4;                  // 16 KB: one [128 x 32] fp32 plane
constexpr int NTHREADS = 320;
constexpr int EPI_BIAS = 1, EPI_RELU = 2, EPI_RESID = 4, EPI_AFFINE = 8, EPI_AUX = 16, EPI_GATE = 32;
// AUX: + aux[m][0..2] . auxw[0..2][n] before the ReLU; GATE: the result is kept where resid[m][n] > 0, else 0 (ReLU backward)

constexpr uint32_t IDESC = make_idesc_tf32(128, 128);

struct Epilogue {
  const float* bias; const float* resid; int ldr; const float* scale; const float* shift;
  const float* aux; const float* auxw; int n_cols;      // EPI_AUX: aux [M,3] row-major, auxw [3][n_cols]
};

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];      // stage filled (4 converter-warp arrivals [+1 producer arrival carrying the TMA bytes])
  uint64_t empty[MAX_STAGES];     // stage consumed (tcgen05.commit)
  uint64_t w_ready;           // resident weights landed
  uint64_t acc_full[2];       // accumulator ready for the epilogue (tcgen05.commit)
  uint64_t acc_empty[2];      // accumulator drained (4 epilogue-warp arrivals)
  uint64_t raw_full[8];       // raw fp32 activation tile landed (TMA bytes)        [A_TMEM mode]
  uint64_t raw_empty[8];      // raw tile consumed by the 4 converter warps          [A_TMEM mode]
  uint32_t tmem_base;
};
constexpr int RAW_STAGES = EVRP_GEMM_RAW_STAGES;      // raw activation staging ring (16 KB each)

// A_TMEM: the converters write the activation planes into tensor memory (tcgen05.st) and the MMAs read A from
// TMEM, so shared memory only carries the weight planes (halves the shared-memory traffic per MMA).
constexpr int ACC_COLS = 2 * BN;                     // two accumulators
constexpr int A_STAGE_COLS = 2 * BKB;                // hi | lo planes of one 32-wide K block

template <bool RESIDENT, int EPI, bool A_TMEM>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm3xtf32_kernel(const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWl,
                  const __grid_constant__ CUtensorMap mapA, const float* __restrict__ A, int lda, float* __restrict__ C, int ldc, int M, int K, int n_chunks,
                  Epilogue ep) {
  constexpr int STAGES = A_TMEM ? EVRP_GEMM_STAGES : 3;   // TMEM A staging: 64 columns per stage next to 2 x 128 accumulator columns
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment (swizzle atoms) by pointer arithmetic on the __shared__ array itself: the compiler keeps the
  // address space, so every access below is an LDS/STS (an integer round trip degrades them to generic LD/ST)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // carve: [resident W: (K/32) x (hi,lo) planes] [ring: STAGES x (A hi, A lo [, W hi, W lo])] [barriers]
  const int KB = K / BKB;
  unsigned char* wres = smem;
  unsigned char* ring = smem + (RESIDENT ? KB * 2 * PLANE : 0);
  // per-stage shared memory: A planes (SS mode only) followed by W planes (streamed mode only)
  constexpr int A_BYTES = A_TMEM ? 0 : 2 * PLANE;
  constexpr int STAGE_BYTES = A_BYTES + (RESIDENT ? 0 : 2 * PLANE);
  // A_TMEM mode: raw fp32 activation tiles [128 x 32] land here by TMA (128B-swizzled), the converters read
  // their row back conflict-free and push the TF32 planes to tensor memory
  unsigned char* raw = ring + STAGES * STAGE_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(raw + (A_TMEM ? RAW_STAGES * PLANE : 0));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % n_chunks;
  const int cta_in_chunk = blockIdx.x / n_chunks, ctas_per_chunk = gridDim.x / n_chunks;
  const int n0 = chunk * BN;
  const int m_tiles = (M + BM - 1) / BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->full[s], RESIDENT ? 4 : 5); mbar_init(&bars->empty[s], 1); }
    mbar_init(&bars->w_ready, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->acc_full[b], 1); mbar_init(&bars->acc_empty[b], 4); }
    for (int r = 0; r < RAW_STAGES; ++r) { mbar_init(&bars->raw_full[r], 1); mbar_init(&bars->raw_empty[r], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {   // TMEM: 2 accumulators x 128 columns
    if (A_TMEM) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base)) : "memory");
    else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&bars->tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 9) {
    // ================= TMA producer: weight planes (+ raw activation tiles in A_TMEM mode) =================
    if (lane == 0) {
      if (RESIDENT) {
        mbar_expect_tx(&bars->w_ready, (uint32_t)(KB * 2 * PLANE));
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_2d(wres + (kb * 2 + 0) * PLANE, &mapWh, &bars->w_ready, kb * BKB, n0);
          tma_load_2d(wres + (kb * 2 + 1) * PLANE, &mapWl, &bars->w_ready, kb * BKB, n0);
        }
      }
      if (!RESIDENT || A_TMEM) {
        uint32_t it = 0;
        for (int tile = cta_in_chunk; tile < m_tiles; tile += ctas_per_chunk) {
          for (int kb = 0; kb < KB; ++kb, ++it) {
            if (A_TMEM) {
              const int rsg = it % RAW_STAGES;
              mbar_wait(&bars->raw_empty[rsg], ((it / RAW_STAGES) & 1) ^ 1);
              mbar_expect_tx(&bars->raw_full[rsg], PLANE);
              tma_load_2d(raw + rsg * PLANE, &mapA, &bars->raw_full[rsg], kb * BKB, tile * BM);   // OOB rows -> 0
            }
            if (!RESIDENT) {
              const int s = it % STAGES;
              mbar_wait(&bars->empty[s], ((it / STAGES) & 1) ^ 1);
              unsigned char* st = ring + s * STAGE_BYTES;
              mbar_expect_tx(&bars->full[s], 2 * PLANE);
              tma_load_2d(st + A_BYTES, &mapWh, &bars->full[s], kb * BKB, n0);
              tma_load_2d(st + A_BYTES + PLANE, &mapWl, &bars->full[s], kb * BKB, n0);
            }
          }
        }
      }
    }
  } else if (warp == 8) {
    // ================= MMA issuer =================
    if (lane == 0) {
      if (RESIDENT) mbar_wait(&bars->w_ready, 0);
      uint32_t it = 0, tcount = 0;
      for (int tile = cta_in_chunk; tile < m_tiles; tile += ctas_per_chunk, ++tcount) {
        const int buf = tcount & 1;
        mbar_wait(&bars->acc_empty[buf], ((tcount >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&bars->full[s], (it / STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          unsigned char* st = ring + s * STAGE_BYTES;
          const uint32_t a_hi = smem_u32(st), a_lo = smem_u32(st + PLANE);
          const uint32_t w_hi = RESIDENT ? smem_u32(wres + (kb * 2) * PLANE) : smem_u32(st + A_BYTES);
          const uint32_t w_lo = w_hi + PLANE;
          const uint32_t ta_hi = tmem_base + ACC_COLS + s * A_STAGE_COLS, ta_lo = ta_hi + BKB;
#pragma unroll
          for (int k = 0; k < BKB / 8; ++k) {            // UMMA_K = 8 tf32 = 32 bytes inside the swizzle row
            const uint64_t dwh = umma_desc(w_hi + k * 32), dwl = umma_desc(w_lo + k * 32);
            if (A_TMEM) {
              umma_tf32_ts(d_tmem, ta_hi + k * 8, dwh, (kb | k) ? 1u : 0u, IDESC);
              umma_tf32_ts(d_tmem, ta_hi + k * 8, dwl, 1u, IDESC);
              umma_tf32_ts(d_tmem, ta_lo + k * 8, dwh, 1u, IDESC);
            } else {
              const uint64_t dah = umma_desc(a_hi + k * 32), dal = umma_desc(a_lo + k * 32);
              umma_tf32(d_tmem, dah, dwh, (kb | k) ? 1u : 0u, IDESC);
              umma_tf32(d_tmem, dah, dwl, 1u, IDESC);
              umma_tf32(d_tmem, dal, dwh, 1u, IDESC);
            }
          }
          umma_commit(&bars->empty[s]);                  // frees the stage when these MMAs retire
        }
        umma_commit(&bars->acc_full[buf]);
      }
    }
  } else if (warp >= 4 && A_TMEM) {
    // ================= A converters (TMEM): thread = row; raw fp32 (smem, TMA-fed) -> (hi, lo) TF32 -> tcgen05.st ====
    const int t = threadIdx.x - 128;                     // 0..127 = row of the tile = TMEM lane
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int my_tiles = (m_tiles - cta_in_chunk + ctas_per_chunk - 1) / ctas_per_chunk;
    const uint32_t total = (uint32_t)my_tiles * KB;
    for (uint32_t it = 0; it < total; ++it) {
      const int rsg = it % RAW_STAGES, s = it % STAGES;
      mbar_wait(&bars->raw_full[rsg], (it / RAW_STAGES) & 1);
      const unsigned char* rp = raw + rsg * PLANE + t * 128;
      float4 x[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) x[c] = *reinterpret_cast<const float4*>(rp + ((c ^ (t & 7)) << 4));   // TMA 128B swizzle
      // the converted values below depend on x[], so the loads have completed before the release further down
      mbar_wait(&bars->empty[s], ((it / STAGES) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem_base + lane_base + ACC_COLS + s * A_STAGE_COLS;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = x[half * 4 + i];
          hi[4 * i + 0] = to_tf32(v.x); hi[4 * i + 1] = to_tf32(v.y); hi[4 * i + 2] = to_tf32(v.z); hi[4 * i + 3] = to_tf32(v.w);
          lo[4 * i + 0] = to_tf32(v.x - __uint_as_float(hi[4 * i + 0])); lo[4 * i + 1] = to_tf32(v.y - __uint_as_float(hi[4 * i + 1]));
          lo[4 * i + 2] = to_tf32(v.z - __uint_as_float(hi[4 * i + 2])); lo[4 * i + 3] = to_tf32(v.w - __uint_as_float(hi[4 * i + 3]));
        }
        tmem_st16(ta + half * 16, hi);
        tmem_st16(ta + BKB + half * 16, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();                                      // one arrival per warp: per-thread arrivals serialise on the barrier
      if (lane == 0) { mbar_arrive(&bars->full[s]); mbar_arrive(&bars->raw_empty[rsg]); }   // raw tile consumed (values already in TMEM)
    }
  } else if (warp >= 4) {
    // ================= A converters: fp32 -> (hi, lo) TF32 planes, swizzled K-major =================
    // Loads run PF k-blocks ahead of the conversion (register ring) so that global latency is off the critical path.
    const int t = threadIdx.x - 128;                     // 0..127
    const int c = t & 7, r0 = t >> 3;                    // 16-byte chunk within the 128-byte row, row within a 16-row slab
    constexpr int PF = 3;
    const int my_tiles = (m_tiles - cta_in_chunk + ctas_per_chunk - 1) / ctas_per_chunk;
    const uint32_t total = (uint32_t)my_tiles * KB;
    float4 v[PF][8];
    auto issue = [&](uint32_t j, float4 (&dst)[8]) {
      const int tile = cta_in_chunk + (int)(j / KB) * ctas_per_chunk, kb = (int)(j % KB);
      const int m0 = tile * BM;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = i * 16 + r0;
        dst[i] = (j < total && m0 + row < M)
                     ? __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + row) * lda + kb * BKB + c * 4))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
#pragma unroll
    for (int p = 0; p < PF; ++p) issue(p, v[p]);
    for (uint32_t it0 = 0; it0 < total; it0 += PF) {
#pragma unroll
      for (int p = 0; p < PF; ++p) {
        const uint32_t it = it0 + p;
        if (it < total) {
          const int s = it % STAGES;
          mbar_wait(&bars->empty[s], ((it / STAGES) & 1) ^ 1);
          unsigned char* st = ring + s * STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = i * 16 + r0;
            const uint32_t off = row * 128 + ((c ^ (row & 7)) << 4);
            const float4 x = v[p][i];
            uint4 hi, lo;
            hi.x = to_tf32(x.x); hi.y = to_tf32(x.y); hi.z = to_tf32(x.z); hi.w = to_tf32(x.w);
            lo.x = to_tf32(x.x - __uint_as_float(hi.x)); lo.y = to_tf32(x.y - __uint_as_float(hi.y));
            lo.z = to_tf32(x.z - __uint_as_float(hi.z)); lo.w = to_tf32(x.w - __uint_as_float(hi.w));
            *reinterpret_cast<uint4*>(st + off) = hi;
            *reinterpret_cast<uint4*>(st + PLANE + off) = lo;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->full[s]);
          issue(it + PF, v[p]);
        }
      }
    }
  } else {
    // ================= epilogue: warp w reads TMEM lanes 32w..32w+31 (row = lane) =================
    uint32_t tcount = 0;
    for (int tile = cta_in_chunk; tile < m_tiles; tile += ctas_per_chunk, ++tcount) {
      const int buf = tcount & 1;
      mbar_wait(&bars->acc_full[buf], (tcount >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // tcgen05.ld.16x256b.x8: 16 TMEM lanes x 64 columns per instruction.  Thread T receives, for i = 0..7,
      //   r[4i], r[4i+1]   = row T/4,     columns 8i + 2(T%4), +1
      //   r[4i+2], r[4i+3] = row T/4 + 8, same columns
      // so every warp-wide 8-byte store writes 8 rows x one full 32-byte sector (no partial-sector traffic).
      const int tq = lane & 3, tr = lane >> 2;
#pragma unroll 1
      for (int part = 0; part < 4; ++part) {
        const int slab = part >> 1, cb = (part & 1) * 64;
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32 + slab * 16) << 16) + buf * BN + cb;
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
            "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int mbase = tile * BM + warp * 32 + slab * 16 + tr;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = n0 + cb + 8 * i + 2 * tq;
          float2 bb = make_float2(0.f, 0.f), sc = make_float2(1.f, 1.f), sh = make_float2(0.f, 0.f);
          if (EPI & EPI_BIAS) bb = __ldg(reinterpret_cast<const float2*>(ep.bias + n));
          if (EPI & EPI_AFFINE) { sc = __ldg(reinterpret_cast<const float2*>(ep.scale + n)); sh = __ldg(reinterpret_cast<const float2*>(ep.shift + n)); }
          float2 aw0 = bb, aw1 = bb, aw2 = bb;
          if (EPI & EPI_AUX) {
            aw0 = __ldg(reinterpret_cast<const float2*>(ep.auxw + n)); aw1 = __ldg(reinterpret_cast<const float2*>(ep.auxw + ep.n_cols + n));
            aw2 = __ldg(reinterpret_cast<const float2*>(ep.auxw + 2 * ep.n_cols + n));
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int m = mbase + 8 * h;
            if (m < M) {
              float v0 = __uint_as_float(r[4 * i + 2 * h]), v1 = __uint_as_float(r[4 * i + 2 * h + 1]);
              if (EPI & EPI_BIAS) { v0 += bb.x; v1 += bb.y; }
              if (EPI & EPI_AUX) {                              // vehicle term of FF_tour.0 (nets/DRLModel.py:215-235)
                const float a0 = __ldg(ep.aux + (size_t)m * 3), a1 = __ldg(ep.aux + (size_t)m * 3 + 1), a2 = __ldg(ep.aux + (size_t)m * 3 + 2);
                v0 += fmaf(a2, aw2.x, fmaf(a1, aw1.x, a0 * aw0.x)); v1 += fmaf(a2, aw2.y, fmaf(a1, aw1.y, a0 * aw0.y));
              }
              if (EPI & EPI_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
              if (EPI & EPI_RESID) {
                const float2 rr = __ldg(reinterpret_cast<const float2*>(ep.resid + (size_t)m * ep.ldr + n));
                v0 += rr.x; v1 += rr.y;
              }
              if (EPI & EPI_GATE) {                             // gradient through a ReLU whose output is `resid`
                const float2 rr = __ldg(reinterpret_cast<const float2*>(ep.resid + (size_t)m * ep.ldr + n));
                v0 = rr.x > 0.f ? v0 : 0.f; v1 = rr.y > 0.f ? v1 : 0.f;
              }
              if (EPI & EPI_AFFINE) { v0 = fmaf(v0, sc.x, sh.x); v1 = fmaf(v1, sc.y, sh.y); }
              *reinterpret_cast<float2*>(C + (size_t)m * ldc + n) = make_float2(v0, v1);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
    }
  }
  __syncthreads();
  if (warp == 8) {
    if (A_TMEM) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

// row-major fp32 matrix [rows][ld] (weights plane [N][K] or activations [M][lda]); box = 32 (K) x box_rows,
// 128-byte swizzle; out-of-bounds rows are zero-filled (ragged last M tile)
int make_weight_map(CUtensorMap* map, const float* W, int N, int K, int ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -EVRP_E_BADARG;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)(ld ? ld : K) * 4};
  cuuint32_t box[2] = {BKB, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)W, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)(2000 + r);
}

int make_weight_map_f16(CUtensorMap* map, const void* W, int N, int K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -EVRP_E_BADARG;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)W, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)(2000 + r);
}

template <bool RESIDENT, int EPI, bool A_TMEM>
static int launch_impl(const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& ma, const float* A, int lda, float* C, int ldc, int M, int N,
                  int K, const Epilogue& ep, int sm_count, cudaStream_t st) {
  const int n_chunks = N / BN;
  const int per_chunk = sm_count / n_chunks > 0 ? sm_count / n_chunks : 1;
  const int m_tiles = (M + BM - 1) / BM;
  const int ctas_per_chunk = per_chunk < m_tiles ? per_chunk : m_tiles;
  const size_t a_bytes = A_TMEM ? 0 : 2 * PLANE;
  constexpr int STAGES = A_TMEM ? EVRP_GEMM_STAGES : 3;
  const size_t smem = (RESIDENT ? (size_t)(K / BKB) * 2 * PLANE + STAGES * a_bytes : (size_t)STAGES * (a_bytes + 2 * PLANE)) +
                      (A_TMEM ? (size_t)RAW_STAGES * PLANE : 0) + sizeof(Barriers) + 1024;
  auto kern = gemm3xtf32_kernel<RESIDENT, EPI, A_TMEM>;
  EVRP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<n_chunks * ctas_per_chunk, NTHREADS, smem, st>>>(mh, ml, ma, A, lda, C, ldc, M, K, n_chunks, ep);
  EVRP_LAUNCH_OK();
  return 0;
}

static bool a_in_tmem() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("EVRP_TC_A_SMEM"); v = (e && e[0] == '1') ? 0 : 1; }
  return v != 0;
}

template <bool RESIDENT, int EPI>
static int launch(const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& ma, const float* A, int lda, float* C,
                  int ldc, int M, int N, int K, const Epilogue& ep, int sm_count, cudaStream_t st) {
  return a_in_tmem() ? launch_impl<RESIDENT, EPI, true>(mh, ml, ma, A, lda, C, ldc, M, N, K, ep, sm_count, st)
                     : launch_impl<RESIDENT, EPI, false>(mh, ml, ma, A, lda, C, ldc, M, N, K, ep, sm_count, st);
}

}  // namespace tc

// C = epi(A @ W^T) with W given as pre-split TF32 planes Wh, Wl ([N][K]).  K in {128, 512}, N % 128 == 0.
int tc_gemm(int epi, const float* A, int lda, const float* Wh, const float* Wl, float* C, int ldc, int M, int N, int K,
            const float* bias, const float* resid, int ldr, const float* scale, const float* shift, int sm_count,
            cudaStream_t st, const float* aux, const float* auxw) {
  using namespace tc;
  if (N % BN || K % BKB || K < BKB || (K != 128 && K != 512 && epi != 0)) return -EVRP_E_BADARG;
  CUtensorMap mh, ml, ma;
  int rc;
  if ((rc = make_weight_map(&mh, Wh, N, K, 0, BN))) return rc;
  if ((rc = make_weight_map(&ml, Wl, N, K, 0, BN))) return rc;
  if ((rc = make_weight_map(&ma, A, M, K, lda, BN))) return rc;
  Epilogue ep{bias, resid, ldr, scale, shift, aux, auxw, N};
  if (K == 128) {
    switch (epi) {
      case 0: return launch<true, 0>(mh, ml, ma, A, lda, C, ldc, M, N, K, ep, sm_count, st);
      case EPI_BIAS | EPI_RELU: return launch<true, EPI_BIAS | EPI_RELU>(mh, ml, ma, A, lda, C, ldc, M, N, K, ep, sm_count, st);
      case EPI_RESID | EPI_AFFINE: return launch<true, EPI_RESID | EPI_AFFINE>(mh, ml, ma, A, lda, C, ldc, M, N, K, ep, sm_count, st);
      case EPI_BIAS | EPI_RELU | EPI_AUX: return launch<true, EPI_BIAS | EPI_RELU | EPI_AUX>(mh, ml, ma, A, lda, C, ldc, M, N, K, ep, sm_count, st);
      case EPI_GATE: return launch<true, EPI_GATE>(mh, ml, ma, A, lda, C, ldc, M, N, K, ep, sm_count, st);
    }
  } else if (epi == (EPI_BIAS | EPI_RESID | EPI_AFFINE)) {
    return launch<false, EPI_BIAS | EPI_RESID | EPI_AFFINE>(mh, ml, ma, A, lda, C, ldc, M, N, K, ep, sm_count, st);
  } else if (epi == 0) {                       // plain product with streamed weights: any K that is a multiple of 32
    return launch<false, 0>(mh, ml, ma, A, lda, C, ldc, M, N, K, ep, sm_count, st);
  }
  return -EVRP_E_BADARG;
}

}  // namespace evrp

// C ABI: the encoder's fp32-faithful tensor-core GEMM on its own (training path: forward products and the
// input-gradient products of the projections / feed-forward layers).
extern "C" int evrp_gemm_3xtf32(int epilogue, const float* A, int lda, const float* W_hi, const float* W_lo, float* C,
                                int ldc, int M, int N, int K, const float* bias, const float* resid, int ldr,
                                const float* scale, const float* shift, const float* aux, const float* aux_w, void* stream) {
  using namespace evrp::tc;
  if (!A || !W_hi || !W_lo || !C || M < 0 || N < 1 || K < 1 || lda < K || ldc < N) return -EVRP_E_BADARG;
  if (((epilogue & EPI_BIAS) && !bias) || ((epilogue & (EPI_RESID | EPI_GATE)) && !resid) || ((epilogue & EPI_AFFINE) && (!scale || !shift)) ||
      ((epilogue & EPI_AUX) && (!aux || !aux_w)))
    return -EVRP_E_BADARG;
  if (M == 0) return 0;
  int dev = 0, sms = 148;
  EVRP_CUDA_OK(cudaGetDevice(&dev));
  EVRP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return evrp::tc_gemm(epilogue, A, lda, W_hi, W_lo, C, ldc, M, N, K, bias, resid, ldr, scale, shift, sms, (cudaStream_t)stream, aux, aux_w);
}
