// Weight-gradient GEMM of the REINFORCE backward (trainer.py:115-127 -> autograd of the nn.Linear /
// W_query / W_key / W_val / W_out products of nets/GraphEncoder.py:55-118,153-170 and FF_tour nets/DRLModel.py:233-237):
//
//   gW[n][k] = sum_m gy[m][n] * x[m][k]          gy [M, N_out], x [M, K_in], M = B*S or B*T rows (1e5 .. 2e5)
//   gaux[i][n] = sum_m aux[m][i] * gy[m][n]      aux column 0 = ones (bias gradient), columns 1..n_aux from `aux`
//
// The contraction runs over the LONG dimension, so both operands arrive "M-major" (the contracted index is the slow
// one).  tcgen05 wants K-major operands; the kernel transposes while it splits:
//   * TMA brings raw fp32 tiles [32 rows of m][128 columns], un-swizzled, into shared memory (loading the columns straight
//     from global memory into registers, one block ahead, measured 1.5 x slower: latency-bound);
//   * converter A (4 warps, thread = column p of the x tile = TMEM lane) reads its column (conflict-free), splits every
//     value into TF32 (hi, lo) with round-to-nearest and writes the planes to TENSOR MEMORY (tcgen05.st): the A operand
//     [128 p x 32 m] never touches shared memory again;
//   * converter B (4 warps, thread = column q of the gy tile) does the same split and writes the planes K-major with the
//     128-byte swizzle ([128 q rows][32 m] = one swizzle row per q), where tcgen05.mma reads them through descriptors;
//   * one lane issues tcgen05.mma.cta_group::1.kind::tf32 (M = N = 128, K = 8): 4 k-steps x 3 products
//     (Ah Bh + Ah Bl + Al Bh, fp32 accumulation in TMEM) per 32-row block - the fp32-faithful scheme of evrp_tc_gemm.cu.
// Split-K: a CTA owns one 128 x 128 output tile and one contiguous slice of the m blocks; its accumulator goes to a
// partial buffer [slice][N_out][K_in] and a second kernel sums the slices in a FIXED order (deterministic, unlike
// floating-point atomics).  CTAs of the same slice run side by side, so the re-read of a raw tile by the other
// output tiles of that slice hits L2 (ncu: DRAM bytes = operand bytes).
// Accumulation length: the tensor core adds into its fp32 accumulator with truncation, a one-sided error that grows with
// the number of accumulations (2e-5 relative after 2,900 rows).  The accumulator is therefore DRAINED every DRAIN blocks
// (512 rows): two TMEM accumulators alternate, four drain warps add the finished one to the partial tile with
// round-to-nearest fp32 adds while the MMAs fill the other.
#include <cuda.h>
#include "evrp_tc.cuh"

namespace evrp {
namespace wg {

using namespace tc;

constexpr int TB = 32;                         // m rows per block = one 128-byte swizzle row of TF32
constexpr int TILE = 128;
constexpr int PLANE = TILE * TB * 4;           // 16 KB
#ifndef EVRP_WG_STAGES
#define EVRP_WG_STAGES 3
#define EVRP_WG_RAW 4
#endif
constexpr int STAGES = EVRP_WG_STAGES;         // operand stages: A planes in TMEM, B planes in shared memory
constexpr int RAW = EVRP_WG_RAW;               // raw fp32 tiles in flight per operand
constexpr int DRAIN = 16;                      // blocks (of 32 rows) accumulated in tensor memory between two drains
constexpr int NTHREADS = 704;                  // 2 x 4 converter-A + 2 x 4 converter-B + MMA + TMA + 4 drain warps
constexpr int ACC_COLS = 256;                  // TMEM: two accumulators, then STAGES x (hi 32 | lo 32) columns of A
constexpr int A_STAGE_COLS = 2 * TB;
constexpr int MAX_AUX = 3;
constexpr uint32_t IDESC = make_idesc_tf32(128, 128);

struct __align__(8) Bars {
  uint64_t full[STAGES];        // 8 converter-warp arrivals: A planes in TMEM + B planes in shared memory
  uint64_t empty[STAGES];       // the MMAs over that stage retired (tcgen05.commit)
  uint64_t rawp_full[RAW], rawp_empty[RAW];
  uint64_t rawq_full[RAW], rawq_empty[RAW];
  uint64_t acc_full[2];         // accumulator holds a finished group of DRAIN blocks (tcgen05.commit)
  uint64_t acc_empty[2];        // drained (4 drain-warp arrivals)
  uint32_t tmem_base;
};

constexpr size_t SMEM_BYTES = (size_t)(2 * RAW + 2 * STAGES) * PLANE + sizeof(Bars) + 1024;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// x = P operand [M, K_in] (tile column block pc), gy = Q operand [M, N_out] (tile column block qc)
__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapQ, int M, int np, int nq,
             int slices, const float* __restrict__ aux, int n_aux, float* __restrict__ partial,
             float* __restrict__ partial_aux) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* rawP = smem;                              // [RAW][32 m][128 p] fp32
  unsigned char* rawQ = smem + RAW * PLANE;                // [RAW][32 m][128 q] fp32
  unsigned char* bst = smem + 2 * RAW * PLANE;             // [STAGES][hi, lo][128 q][32 m] TF32, 128B-swizzled
  Bars* bars = reinterpret_cast<Bars*>(smem + (2 * RAW + 2 * STAGES) * PLANE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = np * nq;
  const int tile = blockIdx.x % tiles, slice = blockIdx.x / tiles;
  const int pc = tile % np, qc = tile / np;
  const int nb = (M + TB - 1) / TB;
  const int b0 = (int)((long long)slice * nb / slices), b1 = (int)((long long)(slice + 1) * nb / slices);
  const int cnt = b1 - b0;
  const int ldo = np * TILE, n_out = nq * TILE;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->full[s], 8); mbar_init(&bars->empty[s], 1); }
    for (int r = 0; r < RAW; ++r) {
      mbar_init(&bars->rawp_full[r], 1); mbar_init(&bars->rawp_empty[r], 4);
      mbar_init(&bars->rawq_full[r], 1); mbar_init(&bars->rawq_empty[r], 4);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->acc_full[b], 1); mbar_init(&bars->acc_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 17) {
    // ================= TMA producer: raw fp32 tiles of both operands =================
    if (lane == 0) {
      for (int i = 0; i < cnt; ++i) {
        const int r = i % RAW;
        const uint32_t ph = ((i / RAW) & 1) ^ 1;
        mbar_wait(&bars->rawp_empty[r], ph);
        mbar_expect_tx(&bars->rawp_full[r], PLANE);
        tma_load_2d(rawP + r * PLANE, &mapP, &bars->rawp_full[r], pc * TILE, (b0 + i) * TB);     // rows >= M arrive as zeros
        mbar_wait(&bars->rawq_empty[r], ph);
        mbar_expect_tx(&bars->rawq_full[r], PLANE);
        tma_load_2d(rawQ + r * PLANE, &mapQ, &bars->rawq_full[r], qc * TILE, (b0 + i) * TB);
      }
    }
  } else if (warp == 16) {
    // ================= MMA issuer =================
    if (lane == 0) {
      for (int i = 0; i < cnt; ++i) {
        const int s = i % STAGES, grp = i / DRAIN, buf = grp & 1, ig = i % DRAIN;
        if (ig == 0) {                                       // new group: its accumulator must have been drained
          mbar_wait(&bars->acc_empty[buf], ((grp >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        mbar_wait(&bars->full[s], (i / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + buf * TILE;
        const uint32_t b_hi = smem_u32(bst + s * 2 * PLANE), b_lo = b_hi + PLANE;
        const uint32_t ta_hi = tmem_base + ACC_COLS + s * A_STAGE_COLS, ta_lo = ta_hi + TB;
#pragma unroll
        for (int k = 0; k < TB / 8; ++k) {
          const uint64_t dbh = umma_desc(b_hi + k * 32), dbl = umma_desc(b_lo + k * 32);
          umma_tf32_ts(d_tmem, ta_hi + k * 8, dbh, (ig | k) ? 1u : 0u, IDESC);
          umma_tf32_ts(d_tmem, ta_hi + k * 8, dbl, 1u, IDESC);
          umma_tf32_ts(d_tmem, ta_lo + k * 8, dbh, 1u, IDESC);
        }
        umma_commit(&bars->empty[s]);
        if (ig == DRAIN - 1 || i == cnt - 1) umma_commit(&bars->acc_full[buf]);
      }
    }
  } else if (warp >= 18) {
    // ================= drain warps: finished accumulator -> partial tile (fp32 adds with round-to-nearest) =================
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int tl = (warp & 3) * 32 + lane;
    float* out = partial + (size_t)slice * n_out * ldo + (size_t)(qc * TILE) * ldo + pc * TILE + tl;
    const int groups = (cnt + DRAIN - 1) / DRAIN;
    if (groups == 0) {
      for (int q = 0; q < TILE; ++q) out[(size_t)q * ldo] = 0.f;
    }
    for (int grp = 0; grp < groups; ++grp) {
      const int buf = grp & 1;
      mbar_wait(&bars->acc_full[buf], (grp >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cb = 0; cb < 8; ++cb) {
        uint32_t v[16];
        float o[16];
        if (grp > 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = out[(size_t)(cb * 16 + j) * ldo];        // L2-resident, this thread's own words
        }
        tmem_ld16(tmem_base + lane_base + buf * TILE + cb * 16, v);
#pragma unroll
        for (int j = 0; j < 16; ++j)                                                   // 32 lanes = 128 B of output row q
          out[(size_t)(cb * 16 + j) * ldo] = grp > 0 ? o[j] + __uint_as_float(v[j]) : __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty[buf]);
    }
  } else if (warp < 8) {
    // ================= converter A (two groups of 4 warps, alternate blocks): thread = column p of the x tile = TMEM lane ====
    // One group alone is latency-bound (wait, column reads, split, tcgen05.st, wait::st, arrive: ~1,600 cycles per block
    // against ~800 of MMA time); two groups work on consecutive blocks at the same time.
    const int t = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int i = warp >> 2; i < cnt; i += 2) {
      const int r = i % RAW, s = i % STAGES;
      mbar_wait(&bars->rawp_full[r], (i / RAW) & 1);
      const float* rp = reinterpret_cast<const float*>(rawP + r * PLANE) + t;
      float x[TB];
#pragma unroll
      for (int m = 0; m < TB; ++m) x[m] = rp[m * TILE];
      mbar_wait(&bars->empty[s], ((i / STAGES) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem_base + lane_base + ACC_COLS + s * A_STAGE_COLS;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v = x[half * 16 + j];
          hi[j] = to_tf32(v);
          lo[j] = to_tf32(v - __uint_as_float(hi[j]));
        }
        tmem_st16(ta + half * 16, hi);
        tmem_st16(ta + TB + half * 16, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();                                          // one arrival per warp (per-thread arrivals serialise on the barrier)
      if (lane == 0) { mbar_arrive(&bars->full[s]); mbar_arrive(&bars->rawp_empty[r]); }
    }
  } else if (warp < 16) {
    // ================= converter B (two groups, alternate blocks): thread = column q of the gy tile -> swizzled TF32 planes ====
    const int t = (warp & 3) * 32 + lane;
    const bool do_aux = (pc == 0) && (partial_aux != nullptr);
    float asum0 = 0.f, asum1 = 0.f, asum2 = 0.f, asum3 = 0.f;
    for (int i = (warp >> 2) & 1; i < cnt; i += 2) {
      const int r = i % RAW, s = i % STAGES;
      mbar_wait(&bars->rawq_full[r], (i / RAW) & 1);
      const float* rq = reinterpret_cast<const float*>(rawQ + r * PLANE) + t;
      float g[TB];
#pragma unroll
      for (int m = 0; m < TB; ++m) g[m] = rq[m * TILE];
      if (do_aux) {
#pragma unroll
        for (int m = 0; m < TB; ++m) asum0 += g[m];
        if (n_aux > 0) {
          const int m0 = (b0 + i) * TB;
          for (int m = 0; m < TB; ++m) {
            const float* ap = aux + (size_t)min(m0 + m, M - 1) * n_aux;      // gy rows >= M are zero: any finite aux value will do
            asum1 = fmaf(__ldg(ap), g[m], asum1);
            if (n_aux > 1) asum2 = fmaf(__ldg(ap + 1), g[m], asum2);
            if (n_aux > 2) asum3 = fmaf(__ldg(ap + 2), g[m], asum3);
          }
        }
      }
      mbar_wait(&bars->empty[s], ((i / STAGES) & 1) ^ 1);
      unsigned char* st = bst + s * 2 * PLANE + t * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 hi, lo;
        hi.x = to_tf32(g[4 * c + 0]); hi.y = to_tf32(g[4 * c + 1]); hi.z = to_tf32(g[4 * c + 2]); hi.w = to_tf32(g[4 * c + 3]);
        lo.x = to_tf32(g[4 * c + 0] - __uint_as_float(hi.x)); lo.y = to_tf32(g[4 * c + 1] - __uint_as_float(hi.y));
        lo.z = to_tf32(g[4 * c + 2] - __uint_as_float(hi.z)); lo.w = to_tf32(g[4 * c + 3] - __uint_as_float(hi.w));
        const uint32_t off = (uint32_t)((c ^ (t & 7)) << 4);
        *reinterpret_cast<uint4*>(st + off) = hi;
        *reinterpret_cast<uint4*>(st + PLANE + off) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to tcgen05
      __syncwarp();
      if (lane == 0) { mbar_arrive(&bars->full[s]); mbar_arrive(&bars->rawq_empty[r]); }
    }
    if (do_aux) {
      // the two groups hold the sums of the even / odd blocks: two strips of partial sums per slice
      float* pa = partial_aux + ((size_t)(slice * 2 + ((warp >> 2) & 1)) * (1 + MAX_AUX)) * n_out + qc * TILE + t;
      pa[0] = asum0; pa[n_out] = asum1; pa[2 * n_out] = asum2; pa[3 * n_out] = asum3;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// gW[i] = sum over slices (fixed order); gaux[a][n] likewise.  accumulate != 0: added to what gW / gaux already hold.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_aux, int slices,
                                    int n_elems, int n_out, int n_auxrows, float* __restrict__ gW,
                                    float* __restrict__ gaux, int accumulate) {
  const int i4 = blockIdx.x * blockDim.x + threadIdx.x;
  const int n4 = n_elems >> 2;
  if (i4 < n4) {
    float4 acc = accumulate ? reinterpret_cast<const float4*>(gW)[i4] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < slices; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(partial + (size_t)s * n_elems) + i4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(gW)[i4] = acc;
  } else if (gaux) {
    const int j = i4 - n4;                                  // one thread per (aux row, output column)
    if (j < n_auxrows * n_out) {
      const int a = j / n_out, n = j % n_out;
      float acc = accumulate ? gaux[j] : 0.f;
      for (int s = 0; s < 2 * slices; ++s) acc += __ldg(partial_aux + ((size_t)s * (1 + MAX_AUX) + a) * n_out + n);
      gaux[j] = acc;
    }
  }
}

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

// row-major fp32 matrix [M][ld]: box = 128 columns x 32 rows, no swizzle (the converters read columns), OOB rows -> 0
static int make_raw_map(CUtensorMap* map, const float* X, int M, int cols, int ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -EVRP_E_BADARG;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {TILE, TB};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)(2000 + r);
}

}  // namespace wg
}  // namespace evrp

extern "C" size_t evrp_wgrad_workspace_bytes(void) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  // slices * tiles <= SM count: one 128 x 128 fp32 partial tile + one strip of auxiliary sums per CTA
  return (size_t)sms * (128 * 128 + 2 * (1 + evrp::wg::MAX_AUX) * 128) * sizeof(float);
}

// gW [N_out, K_in] = gy^T x;  gaux [(1 + n_aux), N_out]: row 0 = column sums of gy (bias gradient), row 1 + i =
// sum_m aux[m][i] * gy[m][:]  (gaux may be NULL).  N_out, K_in multiples of 128; ldg >= N_out, ldx >= K_in, multiples of 4.
extern "C" int evrp_wgrad_3xtf32(const float* gy, int ldg, int N_out, const float* x, int ldx, int K_in, int M,
                                 const float* aux, int n_aux, float* gW, float* gaux, int accumulate, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  using namespace evrp::wg;
  if (!gy || !x || !gW || !workspace || M < 1 || N_out < TILE || K_in < TILE || N_out % TILE || K_in % TILE || ldg < N_out ||
      ldx < K_in || (ldg & 3) || (ldx & 3) || n_aux < 0 || n_aux > MAX_AUX || (n_aux > 0 && (!aux || !gaux)))
    return -EVRP_E_BADARG;
  int dev = 0, sms = 148;
  EVRP_CUDA_OK(cudaGetDevice(&dev));
  EVRP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int np = K_in / TILE, nq = N_out / TILE, tiles = np * nq;
  if (tiles > sms) return -EVRP_E_BADARG;
  const int nb = (M + TB - 1) / TB;
  int slices = sms / tiles;
  if (slices > nb) slices = nb;
  const size_t need = ((size_t)slices * N_out * K_in + (size_t)2 * slices * (1 + MAX_AUX) * N_out) * sizeof(float);
  if (workspace_bytes < need) return -EVRP_E_WORKSPACE;
  float* partial = reinterpret_cast<float*>(workspace);
  float* partial_aux = gaux ? partial + (size_t)slices * N_out * K_in : nullptr;
  CUtensorMap mp, mq;
  int rc;
  if ((rc = make_raw_map(&mp, x, M, K_in, ldx))) return rc;
  if ((rc = make_raw_map(&mq, gy, M, N_out, ldg))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  EVRP_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  wgrad_kernel<<<tiles * slices, NTHREADS, SMEM_BYTES, st>>>(mp, mq, M, np, nq, slices, aux, n_aux, partial, partial_aux);
  EVRP_LAUNCH_OK();
  const int n_elems = N_out * K_in;
  const int threads = (n_elems >> 2) + (gaux ? (1 + n_aux) * N_out : 0);
  wgrad_reduce_kernel<<<(threads + 255) / 256, 256, 0, st>>>(partial, partial_aux, slices, n_elems, N_out, 1 + n_aux, gW, gaux,
                                                             accumulate);
  EVRP_LAUNCH_OK();
  return 0;
}
