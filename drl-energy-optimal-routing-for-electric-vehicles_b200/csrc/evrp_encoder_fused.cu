// Fused encoder: init embedding + every MultiHeadAttentionLayer + the decoder's node projection in ONE persistent
// tcgen05 / TMEM / TMA kernel:
//   AttentionModel._init_embed        nets/DRLModel.py:192-200
//   MultiHeadAttentionLayer x L       nets/GraphEncoder.py:153-179  (MHA :55-118, BatchNorm eval :142-150, FF :172-176)
//   project_node_embeddings           nets/DRLModel.py:351-352      (project_out folded into the logit key)
//
// One CTA per SM owns one 128-row TILE of the flattened [B*S] node rows at a time: one instance when S > 64, else
// floor(128 / S) whole instances (their rows are contiguous; the attention masks the keys of the other instances).
// The activations of the tile never leave the SM between the input (3 floats per node) and the outputs (emb, kvl):
// they live in tensor memory as fp16 (hi, lo) operand planes.  Only the weights stream: 16 KB TMA blocks from L2
// through a shared-memory ring, in a fixed order that the host packs once per parameter change.
//
// Precision: greedy tours flip under TF32-level noise.  Every operand is split into two fp16 planes
// with round-to-nearest, x = hi + lo (22 significant bits = a TF32 pair at half the bytes, at the kind::f16 rate), and
// every product is Ah Bh + Al Bh + Ah Bl with fp32 accumulation in TMEM.  Weights are pre-scaled by a power of two so
// that their lo plane stays normal; the accumulators are scaled back in the epilogues.
//
// Tensor memory (512 columns x 128 lanes, lane = tile row):
//   X  [  0,128)  operand planes of the layer input / residual stream: hi 64 | lo 64 columns (2 channels per column)
//   S  [128,256)  accumulator 0: Q / V chunk, scores of a head, W_out product, FF1 chunk, node-projection chunk
//   P  [256,384)  accumulator 1 (K chunk, node chunk 1);  softmax probabilities / FF hidden planes (A operands)
//   QO [384,512)  Q planes per head (hi 8 | lo 8 columns), overwritten head by head by the P V accumulators, then (in
//                 place) the operand planes of the concatenated heads; FF2 accumulator
// Shared memory: K planes 64 KB + V^T planes 64 KB (B operands of the attention MMAs, written by the epilogue of the
// QKV product) + weight ring 5 x 16 KB + softmax exchange 8 KB.
//
// Warp roles (576 threads): warps 0-15 workers: thread = (tile row = TMEM lane, column quarter q): every epilogue,
// the softmax, the operand splits; warp 16: one lane issues every tcgen05.mma; warp 17: one lane drives TMA.
// All three roles run the same static per-tile program and meet on mbarriers; MMAs of one stream overlap the
// epilogue of the other (QKV chunk c+1 | chunk c, scores h+1 | softmax h | P V h-1, FF1 j+1 | ReLU j | FF2 j-1).
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include "evrp_tc.cuh"

namespace evrp {
namespace fenc {

using namespace tc;

constexpr int NWORK = 16;
constexpr int NT = (NWORK + 2) * 32;
constexpr int BLK = 128 * 128;                 // bytes of one [128 rows x 64 fp16] operand block (128-byte swizzle rows)
constexpr int NSLOT = 5;
constexpr int TM_X = 0, TM_S = 128, TM_P = 256, TM_QO = 384;
constexpr uint32_t IDESC_G = make_idesc_f16(128, 128);
constexpr uint32_t IDESC_PV = make_idesc_f16(128, 16);
constexpr int LAYER_BLOCKS = 48;               // weight blocks per layer: QKV 12, W_out 4, FF1 16, FF2 16
// per-layer fp32 parameter vector (host: policy_math.pack_fused_encoder)
constexpr int LP = 1280;
constexpr int P_BN1S = 0, P_BN1B = 128, P_B1 = 256, P_B2 = 768, P_BN2S = 896, P_BN2B = 1024, P_INV = 1152;

struct __align__(8) Bars {
  uint64_t full[NSLOT], empty[NSLOT];
  uint64_t xp_full;        // X planes written by the workers (16 warp arrivals)
  uint64_t acc_full[2];    // accumulator complete (tcgen05.commit)
  uint64_t acc_empty[2];   // accumulator read into registers (16)
  uint64_t hp_full;        // P region planes written (16)
  uint64_t hp_empty;       // MMAs that read the P region retired (commit)
  uint64_t y_full;         // QO accumulators complete: all heads' P V / the FF2 product (commit)
  uint64_t qo_ready;       // QO operand planes written: Q planes / concatenated heads (16)
  uint64_t kv_ready;       // K and V^T planes in shared memory (16)
  uint32_t tmem_base;
};

constexpr size_t SMEM_BYTES = 8 * (size_t)BLK /*K, V^T planes*/ + NSLOT * (size_t)BLK + 2 * 2 * 4 * 128 * 4 /*pmax, psum*/ +
                              sizeof(Bars) + 1024;

// ---- diagnostics: a wait that does not complete within ~2 s records where it was and releases every other wait ----
__device__ int g_abort;
__device__ int g_dbg[8];

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity, int code) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (!ok && (spin & 255u) == 255u) {
      if (*(volatile int*)&g_abort) return;
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {
        if (atomicCAS(&g_abort, 0, 1) == 0) { g_dbg[0] = code; g_dbg[1] = blockIdx.x; g_dbg[2] = threadIdx.x; g_dbg[3] = (int)parity; }
        return;
      }
    }
  }
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
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// A operand from TMEM (lane = row, one 32-bit column = two consecutive fp16 K elements), kind::f16
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// one warp = one arrival: every lane has fenced its own tensor-memory / shared-memory writes before the __syncwarp
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void row_group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

// 32 fp32 values -> 16 + 16 packed fp16 (hi, lo) pairs, element 2e in the low half (= the lower K index)
__device__ __forceinline__ void split32(const float (&v)[32], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
    hi[e] = *reinterpret_cast<const uint32_t*>(&h);
    lo[e] = *reinterpret_cast<const uint32_t*>(&l);
  }
}
__device__ __forceinline__ void join32(const uint32_t (&hi)[16], const uint32_t (&lo)[16], float (&x)[32]) {
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi[e]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lo[e]));
    x[2 * e] = a.x + b.x; x[2 * e + 1] = a.y + b.y;
  }
}

struct Args {
  const float* params;                 // [n_layers][LP] + [1] inverse scale of the node projection
  const float *depot_w, *depot_b, *station_w, *station_b, *cust_w, *cust_b;
  const float* xy;                     // [B,2,S]
  const float* dyn;                    // [B,4,S]
  int B, S, F, n_layers, n_node_chunks;
  float* emb;                          // [B*S,128]
  float* kvl;                          // [B*S,384]
  float* step_tab;                     // [B*S,128] (node chunk 3) or NULL
};

__global__ void __launch_bounds__(NT, 1)
encoder_fused_kernel(const __grid_constant__ CUtensorMap mapW, const Args a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* KP = smem;                          // [channel block 0..1][plane hi, lo][128 keys x 128 B]
  unsigned char* VT = smem + 4 * BLK;                // [key block 0..1][plane hi, lo][128 channels x 128 B]
  unsigned char* ring = smem + 8 * BLK;
  float* pmax = reinterpret_cast<float*>(ring + NSLOT * BLK);     // [parity][q][row]
  float* psum = pmax + 2 * 4 * 128;                               // [parity][q][row]
  Bars* bars = reinterpret_cast<Bars*>(psum + 2 * 4 * 128);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int S = a.S;
  const int inst_per_tile = S > 64 ? 1 : 128 / S;
  const int rows_per_tile = inst_per_tile * S;
  const int M = a.B * S;
  const int n_tiles = (a.B + inst_per_tile - 1) / inst_per_tile;
  const int tile_blocks = a.n_layers * LAYER_BLOCKS + a.n_node_chunks * 4;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    mbar_init(&bars->xp_full, NWORK);
    for (int b = 0; b < 2; ++b) { mbar_init(&bars->acc_full[b], 1); mbar_init(&bars->acc_empty[b], NWORK); }
    mbar_init(&bars->hp_full, NWORK); mbar_init(&bars->hp_empty, 1);
    mbar_init(&bars->y_full, 1); mbar_init(&bars->qo_ready, NWORK); mbar_init(&bars->kv_ready, NWORK);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NWORK) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == NWORK + 1) {
    // =========================== TMA producer: the tile's weight blocks, in consumption order ===========================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int i = 0; i < tile_blocks; ++i, ++it) {
          const int s = it % NSLOT;
          wait_bar(&bars->empty[s], ((it / NSLOT) & 1) ^ 1, 100 + s);
          mbar_expect_tx(&bars->full[s], BLK);
          tma_load_2d(ring + s * BLK, &mapW, &bars->full[s], 0, i * 128);
        }
      }
    }
  } else if (warp == NWORK) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      uint32_t it = 0;                       // ring position
      uint32_t n_xp = 0, n_qo = 0, n_kv = 0, n_hp = 0;
      uint32_t n_acc[2] = {0, 0};            // products written to each accumulator so far
      const int nks = (rows_per_tile + 15) >> 4;          // 16-key steps of the P V product
      // D[128 x 128] (+)= A[128 x 128] W^T: A planes in TMEM (hi at a_tmem, lo at a_tmem + 64), four ring blocks
      // (K block 0 hi, lo, K block 1 hi, lo)
      auto gemm_k128 = [&](uint32_t d_tmem, uint32_t a_tmem, bool accumulate) {
#pragma unroll 1
        for (int kb = 0; kb < 2; ++kb) {
          const int s0 = it % NSLOT, s1 = (it + 1) % NSLOT;
          wait_bar(&bars->full[s0], (it / NSLOT) & 1, 200 + s0);
          wait_bar(&bars->full[s1], ((it + 1) / NSLOT) & 1, 210 + s1);
          tc_fence_after();
          const uint32_t wh = smem_u32(ring + s0 * BLK), wl = smem_u32(ring + s1 * BLK);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ah = a_tmem + (kb * 4 + k) * 8, al = ah + 64;
            const uint64_t dwh = umma_desc(wh + k * 32), dwl = umma_desc(wl + k * 32);
            umma_f16_ts(d_tmem, ah, dwh, (accumulate || kb || k) ? 1u : 0u, IDESC_G);
            umma_f16_ts(d_tmem, al, dwh, 1u, IDESC_G);
            umma_f16_ts(d_tmem, ah, dwl, 1u, IDESC_G);
          }
          umma_commit(&bars->empty[s0]);
          umma_commit(&bars->empty[s1]);
          it += 2;
        }
      };
      auto acquire_acc = [&](int b) {          // the previous product in this accumulator has been read
        if (n_acc[b] > 0) { wait_bar(&bars->acc_empty[b], (n_acc[b] - 1) & 1, 220 + b); tc_fence_after(); }
        ++n_acc[b];
      };
      const uint32_t tX = tmem_base + TM_X, tS = tmem_base + TM_S, tP = tmem_base + TM_P, tQO = tmem_base + TM_QO;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        wait_bar(&bars->xp_full, n_xp++ & 1, 230); tc_fence_after();
        for (int l = 0; l < a.n_layers; ++l) {
          // ---- Q | K | V = x W_qkv^T, three 128-column chunks into accumulators 0, 1, 0 ----
          for (int c = 0; c < 3; ++c) {
            const int b = c & 1;
            acquire_acc(b);
            gemm_k128(b ? tP : tS, tX, false);
            umma_commit(&bars->acc_full[b]);
          }
          // ---- attention: scores of head h + 1 are issued before the P V product of head h ----
          wait_bar(&bars->qo_ready, n_qo++ & 1, 231);
          wait_bar(&bars->kv_ready, n_kv++ & 1, 232);
          tc_fence_after();
          auto issue_pv = [&](int h) {
            wait_bar(&bars->hp_full, n_hp++ & 1, 233); tc_fence_after();
#pragma unroll 1
            for (int ks = 0; ks < nks; ++ks) {
              const uint32_t vb = smem_u32(VT + (ks >> 2) * (2 * BLK)) + h * 2048 + (ks & 3) * 32;
              const uint64_t dvh = umma_desc(vb), dvl = umma_desc(vb + BLK);
              const uint32_t ph = tP + ks * 8, pl = ph + 64;
              umma_f16_ts(tQO + h * 16, ph, dvh, ks ? 1u : 0u, IDESC_PV);
              umma_f16_ts(tQO + h * 16, pl, dvh, 1u, IDESC_PV);
              umma_f16_ts(tQO + h * 16, ph, dvl, 1u, IDESC_PV);
            }
            umma_commit(&bars->hp_empty);
          };
          for (int h = 0; h < H; ++h) {
            acquire_acc(0);
            const uint32_t kb = smem_u32(KP + (h >> 2) * (2 * BLK)) + (h & 3) * 32;
            const uint64_t dkh = umma_desc(kb), dkl = umma_desc(kb + BLK);
            const uint32_t qh = tQO + h * 16, ql = qh + 8;
            umma_f16_ts(tS, qh, dkh, 0u, IDESC_G);
            umma_f16_ts(tS, ql, dkh, 1u, IDESC_G);
            umma_f16_ts(tS, qh, dkl, 1u, IDESC_G);
            umma_commit(&bars->acc_full[0]);
            if (h > 0) issue_pv(h - 1);
          }
          issue_pv(H - 1);
          umma_commit(&bars->y_full);
          // ---- heads W_out^T ----
          wait_bar(&bars->qo_ready, n_qo++ & 1, 234); tc_fence_after();
          acquire_acc(0);
          gemm_k128(tS, tQO, false);
          umma_commit(&bars->acc_full[0]);
          // ---- FF: hidden chunk j+1 (FF1) is issued before the FF2 partial product over chunk j ----
          wait_bar(&bars->xp_full, n_xp++ & 1, 235); tc_fence_after();
          for (int j = 0; j < 4; ++j) {
            acquire_acc(0);
            gemm_k128(tS, tX, false);
            umma_commit(&bars->acc_full[0]);
            if (j > 0) {
              wait_bar(&bars->hp_full, n_hp++ & 1, 236); tc_fence_after();
              gemm_k128(tQO, tP, j > 1);
              umma_commit(&bars->hp_empty);
            }
          }
          wait_bar(&bars->hp_full, n_hp++ & 1, 237); tc_fence_after();
          gemm_k128(tQO, tP, true);
          umma_commit(&bars->hp_empty);
          umma_commit(&bars->y_full);
          wait_bar(&bars->xp_full, n_xp++ & 1, 238); tc_fence_after();      // next layer's input / the embeddings
        }
        // ---- node projection (glimpse K | glimpse V | folded logit K [| AM step table]) ----
        for (int c = 0; c < a.n_node_chunks; ++c) {
          const int b = c & 1;
          acquire_acc(b);
          gemm_k128(b ? tP : tS, tX, false);
          umma_commit(&bars->acc_full[b]);
        }
      }
    }
  } else {
    // =========================== workers: thread = (row r of the tile = TMEM lane, column quarter q) ===========================
    const int g = warp & 3, q = warp >> 2;
    const int r = g * 32 + lane;
    const uint32_t lane_base = (uint32_t)(g * 32) << 16;
    const uint32_t tX = tmem_base + lane_base + TM_X, tS = tmem_base + lane_base + TM_S,
                   tP = tmem_base + lane_base + TM_P, tQO = tmem_base + lane_base + TM_QO;
    const float c2 = 0.25f * 1.4426950408889634f;         // norm_factor 1/sqrt(16) (nets/GraphEncoder.py:38) and log2(e)
    uint32_t n_accf[2] = {0, 0}, n_hp = 0, n_y = 0;
    // keys this thread's row attends to: the rows of its own instance
    uint32_t kmask = 0;
    {
      int gi = r / S;
      if (gi >= inst_per_tile) gi = inst_per_tile - 1;    // padding rows: any valid key range (their outputs are dropped)
      const int klo = gi * S, khi = klo + S;
#pragma unroll
      for (int e = 0; e < 32; ++e) kmask |= ((32 * q + e >= klo && 32 * q + e < khi) ? 1u : 0u) << e;
    }
    auto wait_acc = [&](int b, int code) { wait_bar(&bars->acc_full[b], n_accf[b]++ & 1, code); tc_fence_after(); };
    auto release_acc = [&](int b) { tmem_ld_wait(); tc_fence_before(); warp_arrive(&bars->acc_empty[b], lane); };
    auto wait_p_free = [&](int code) {       // the MMAs that read the previous contents of the P region have retired
      if (n_hp > 0) { wait_bar(&bars->hp_empty, (n_hp - 1) & 1, code); tc_fence_after(); }
      ++n_hp;
    };
    auto store_x_planes = [&](const float (&v)[32]) {
      uint32_t hi[16], lo[16];
      split32(v, hi, lo);
      tmem_st16(tX + 16 * q, hi);
      tmem_st16(tX + 64 + 16 * q, lo);
      tmem_st_wait(); tc_fence_before();
      warp_arrive(&bars->xp_full, lane);
    };

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int grow = tile * rows_per_tile + r;
      const bool valid = r < rows_per_tile && grow < M;
      // ---- init embedding (nets/DRLModel.py:192-200): depot | stations | customers (x/100, y/100[, demand]) ----
      {
        float v[32];
        if (valid) {
          const int b = grow / S, j = grow - b * S;
          const float x = __fdiv_rn(__ldg(a.xy + ((size_t)b * 2 + 0) * S + j), 100.f);
          const float y = __fdiv_rn(__ldg(a.xy + ((size_t)b * 2 + 1) * S + j), 100.f);
          if (j <= a.F) {
            const float* w = (j == 0 ? a.depot_w : a.station_w) + 64 * q;
            const float* bb = (j == 0 ? a.depot_b : a.station_b) + 32 * q;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float4 ww = __ldg(reinterpret_cast<const float4*>(w + 2 * i));
              const float2 b2 = __ldg(reinterpret_cast<const float2*>(bb + i));
              v[i] = fmaf(ww.y, y, fmaf(ww.x, x, 0.f)) + b2.x;
              v[i + 1] = fmaf(ww.w, y, fmaf(ww.z, x, 0.f)) + b2.y;
            }
          } else {
            const float dem = __ldg(a.dyn + ((size_t)b * 4 + 1) * S + j);
            const float* w = a.cust_w + 96 * q;
            const float* bb = a.cust_b + 32 * q;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              v[i] = fmaf(__ldg(w + 3 * i + 2), dem, fmaf(__ldg(w + 3 * i + 1), y, fmaf(__ldg(w + 3 * i), x, 0.f))) + __ldg(bb + i);
          }
          if (a.n_layers == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(a.emb + (size_t)grow * D + 32 * q + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        store_x_planes(v);
      }

      for (int l = 0; l < a.n_layers; ++l) {
        const float* prm = a.params + (size_t)l * LP;
        const float inv_qkv = __ldg(prm + P_INV), inv_wo = __ldg(prm + P_INV + 1), inv_w1 = __ldg(prm + P_INV + 2),
                    inv_w2 = __ldg(prm + P_INV + 3);
        // ---- Q chunk -> operand planes per head in QO: head 2q at columns 32q (hi 8 | lo 8), head 2q+1 at 32q + 16 ----
        {
          uint32_t u[32];
          wait_acc(0, 300);
          tmem_ld32(tS + 32 * q, u);
          release_acc(0);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(u[i]) * inv_qkv;
          uint32_t hi[16], lo[16], o[32];
          split32(v, hi, lo);
#pragma unroll
          for (int i = 0; i < 8; ++i) { o[i] = hi[i]; o[8 + i] = lo[i]; o[16 + i] = hi[8 + i]; o[24 + i] = lo[8 + i]; }
          tmem_st32(tQO + 32 * q, o);
          tmem_st_wait(); tc_fence_before();
          warp_arrive(&bars->qo_ready, lane);
        }
        // ---- K chunk -> K planes in shared memory: row = key r, channels 32q..32q+31 = 4 swizzled 16-byte chunks ----
        {
          uint32_t u[32];
          wait_acc(1, 301);
          tmem_ld32(tP + 32 * q, u);
          release_acc(1);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(u[i]) * inv_qkv;
          uint32_t hi[16], lo[16];
          split32(v, hi, lo);
          unsigned char* kp = KP + (q >> 1) * (2 * BLK) + r * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int off = ((((q & 1) * 4 + i) ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(kp + off) = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
            *reinterpret_cast<uint4*>(kp + BLK + off) = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
          }
        }
        // ---- V chunk -> V^T planes: element (channel d, key r) -> key block r / 64, row d, fp16 column r % 64 ----
        {
          uint32_t u[32];
          wait_acc(0, 302);
          tmem_ld32(tS + 32 * q, u);
          release_acc(0);
          unsigned char* vb = VT + (r >> 6) * (2 * BLK) + (32 * q) * 128 + ((r & 7) << 1);
          const int kc = (r & 63) >> 3;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = __uint_as_float(u[i]) * inv_qkv;
            const __half hi = __float2half_rn(x);
            const __half lo = __float2half_rn(x - __half2float(hi));
            unsigned char* p = vb + i * 128 + ((kc ^ (i & 7)) << 4);
            *reinterpret_cast<__half*>(p) = hi;
            *reinterpret_cast<__half*>(p + BLK) = lo;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // K and V^T writes -> visible to tcgen05
          warp_arrive(&bars->kv_ready, lane);
        }
        // ---- softmax of every head: p = 2^(c2 (s - max)) unnormalised, masked keys 0; planes to the P region ----
        float tot0 = 1.f, tot1 = 1.f;                     // softmax denominators of this thread's heads 2q, 2q + 1
        for (int h = 0; h < H; ++h) {
          const int par = h & 1;
          uint32_t u[32];
          wait_acc(0, 310 + h);
          tmem_ld32(tS + 32 * q, u);
          release_acc(0);
          float mx = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; ++e) mx = fmaxf(mx, ((kmask >> e) & 1u) ? __uint_as_float(u[e]) : -INFINITY);
          pmax[(par * 4 + q) * 128 + r] = mx;
          row_group_sync(g);
          mx = fmaxf(fmaxf(pmax[(par * 4 + 0) * 128 + r], pmax[(par * 4 + 1) * 128 + r]),
                     fmaxf(pmax[(par * 4 + 2) * 128 + r], pmax[(par * 4 + 3) * 128 + r]));
          if (h > 0) {                                    // the previous head's denominator (written after its barrier)
            const float* ps = psum + ((par ^ 1) * 4) * 128 + r;
            const float t = (ps[0] + ps[128]) + (ps[256] + ps[384]);
            if (((h - 1) >> 1) == q) { if ((h - 1) & 1) tot1 = t; else tot0 = t; }
          }
          const float mx2 = mx * c2;
          float p[32];
          float sum = 0.f;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float x = ((kmask >> e) & 1u) ? fmaf(__uint_as_float(u[e]), c2, -mx2) : -INFINITY;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[e]) : "f"(x));
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) sum += p[e];
          uint32_t hi[16], lo[16];
          split32(p, hi, lo);
          wait_p_free(320 + h);
          tmem_st16(tP + 16 * q, hi);
          tmem_st16(tP + 64 + 16 * q, lo);
          tmem_st_wait(); tc_fence_before();
          warp_arrive(&bars->hp_full, lane);
          psum[(par * 4 + q) * 128 + r] = sum;
        }
        // ---- concatenated heads: O / denominator -> operand planes, in place in QO ----
        {
          row_group_sync(g);                              // the last head's partial denominators are visible
          {
            const float* ps = psum + (((H - 1) & 1) * 4) * 128 + r;
            const float t = (ps[0] + ps[128]) + (ps[256] + ps[384]);
            if (((H - 1) >> 1) == q) tot1 = t;
          }
          wait_bar(&bars->y_full, n_y++ & 1, 330); tc_fence_after();
          uint32_t u[32];
          tmem_ld32(tQO + 32 * q, u);
          tmem_ld_wait();
          const float i0 = 1.f / tot0, i1 = 1.f / tot1;
          float v[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] = __uint_as_float(u[i]) * i0; v[16 + i] = __uint_as_float(u[16 + i]) * i1; }
          uint32_t hi[16], lo[16];
          split32(v, hi, lo);
          tc_fence_before();
          row_group_sync(g);                              // the four threads of a row have read their columns: layout change
          tc_fence_after();
          tmem_st16(tQO + 16 * q, hi);
          tmem_st16(tQO + 64 + 16 * q, lo);
          tmem_st_wait(); tc_fence_before();
          warp_arrive(&bars->qo_ready, lane);
        }
        // ---- x1 = BN1(x + heads W_out^T)  (nets/GraphEncoder.py:162-169): planes rewritten in place ----
        {
          uint32_t u[32], xh[16], xl[16];
          wait_acc(0, 340);
          tmem_ld32(tS + 32 * q, u);
          tmem_ld16(tX + 16 * q, xh);
          tmem_ld16(tX + 64 + 16 * q, xl);
          release_acc(0);
          float x[32], v[32];
          join32(xh, xl, x);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(prm + P_BN1S + 32 * q + i));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(prm + P_BN1B + 32 * q + i));
            v[i] = fmaf(fmaf(__uint_as_float(u[i]), inv_wo, x[i]), sc.x, sh.x);
            v[i + 1] = fmaf(fmaf(__uint_as_float(u[i + 1]), inv_wo, x[i + 1]), sc.y, sh.y);
            v[i + 2] = fmaf(fmaf(__uint_as_float(u[i + 2]), inv_wo, x[i + 2]), sc.z, sh.z);
            v[i + 3] = fmaf(fmaf(__uint_as_float(u[i + 3]), inv_wo, x[i + 3]), sc.w, sh.w);
          }
          store_x_planes(v);
        }
        // ---- hidden chunk j = relu(x1 W1_j^T + b1_j) -> operand planes in the P region ----
        for (int j = 0; j < 4; ++j) {
          uint32_t u[32];
          wait_acc(0, 350 + j);
          tmem_ld32(tS + 32 * q, u);
          release_acc(0);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(prm + P_B1 + 128 * j + 32 * q + i));
            v[i] = fmaxf(fmaf(__uint_as_float(u[i]), inv_w1, bb.x), 0.f);
            v[i + 1] = fmaxf(fmaf(__uint_as_float(u[i + 1]), inv_w1, bb.y), 0.f);
            v[i + 2] = fmaxf(fmaf(__uint_as_float(u[i + 2]), inv_w1, bb.z), 0.f);
            v[i + 3] = fmaxf(fmaf(__uint_as_float(u[i + 3]), inv_w1, bb.w), 0.f);
          }
          uint32_t hi[16], lo[16];
          split32(v, hi, lo);
          wait_p_free(360 + j);
          tmem_st16(tP + 16 * q, hi);
          tmem_st16(tP + 64 + 16 * q, lo);
          tmem_st_wait(); tc_fence_before();
          warp_arrive(&bars->hp_full, lane);
        }
        // ---- x2 = BN2(x1 + hidden W2^T + b2): the next layer's input; the last layer's is the node embedding ----
        {
          uint32_t u[32], xh[16], xl[16];
          wait_bar(&bars->y_full, n_y++ & 1, 370); tc_fence_after();
          tmem_ld32(tQO + 32 * q, u);
          tmem_ld16(tX + 16 * q, xh);
          tmem_ld16(tX + 64 + 16 * q, xl);
          tmem_ld_wait();
          float x[32], v[32];
          join32(xh, xl, x);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(prm + P_B2 + 32 * q + i));
            const float4 sc = __ldg(reinterpret_cast<const float4*>(prm + P_BN2S + 32 * q + i));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(prm + P_BN2B + 32 * q + i));
            v[i] = fmaf(fmaf(__uint_as_float(u[i]), inv_w2, bb.x) + x[i], sc.x, sh.x);
            v[i + 1] = fmaf(fmaf(__uint_as_float(u[i + 1]), inv_w2, bb.y) + x[i + 1], sc.y, sh.y);
            v[i + 2] = fmaf(fmaf(__uint_as_float(u[i + 2]), inv_w2, bb.z) + x[i + 2], sc.z, sh.z);
            v[i + 3] = fmaf(fmaf(__uint_as_float(u[i + 3]), inv_w2, bb.w) + x[i + 3], sc.w, sh.w);
          }
          if (l == a.n_layers - 1 && valid) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(a.emb + (size_t)grow * D + 32 * q + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
          store_x_planes(v);
        }
      }
      // ---- node projection chunks -> kvl (chunks 0..2) / step table (chunk 3) ----
      {
        const float inv_node = __ldg(a.params + (size_t)a.n_layers * LP);
        for (int c = 0; c < a.n_node_chunks; ++c) {
          const int b = c & 1;
          uint32_t u[32];
          wait_acc(b, 380 + c);
          tmem_ld32((b ? tP : tS) + 32 * q, u);
          release_acc(b);
          if (valid) {
            float* o = c < 3 ? a.kvl + (size_t)grow * 384 + 128 * c + 32 * q : a.step_tab + (size_t)grow * D + 32 * q;
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(o + i) = make_float4(__uint_as_float(u[i]) * inv_node, __uint_as_float(u[i + 1]) * inv_node,
                                                               __uint_as_float(u[i + 2]) * inv_node, __uint_as_float(u[i + 3]) * inv_node);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NWORK) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace fenc

// weights: fp16 [n_blocks][128][64] operand blocks in consumption order (policy_math.pack_fused_encoder)
int launch_encoder_fused(const evrp_encoder_weights* w, const float* xy, const float* dyn, int B, int S, int F, float* emb,
                         float* kvl, float* step_tab, int sm_count, cudaStream_t st) {
  using namespace fenc;
  if (!w->fused_w || !w->fused_params) return -EVRP_E_BADARG;
  const int n_node_chunks = step_tab ? 4 : 3;
  if (step_tab && w->fused_node_chunks < 4) return -EVRP_E_BADARG;
  // the node projection is packed with fused_node_chunks chunks; a 3-chunk call on a 4-chunk pack reads the first three
  const int pack_blocks = w->n_layers * LAYER_BLOCKS + w->fused_node_chunks * 4;
  CUtensorMap map;
  int rc;
  if ((rc = tc::make_weight_map_f16(&map, w->fused_w, pack_blocks * 128, 64, 128))) return rc;
  Args a;
  a.params = w->fused_params;
  a.depot_w = w->depot_w; a.depot_b = w->depot_b; a.station_w = w->station_w; a.station_b = w->station_b;
  a.cust_w = w->cust_w; a.cust_b = w->cust_b;
  a.xy = xy; a.dyn = dyn; a.B = B; a.S = S; a.F = F; a.n_layers = w->n_layers; a.n_node_chunks = n_node_chunks;
  a.emb = emb; a.kvl = kvl; a.step_tab = step_tab;
  const int ipt = S > 64 ? 1 : 128 / S;
  const int n_tiles = (B + ipt - 1) / ipt;
  EVRP_CUDA_OK(cudaFuncSetAttribute(encoder_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  encoder_fused_kernel<<<grid, NT, SMEM_BYTES, st>>>(map, a);
  EVRP_LAUNCH_OK();
  return 0;
}

}  // namespace evrp

// diagnostics of the fused encoder kernel: out8[0] = code of the first wait that timed out (0: none), [1] CTA, [2] thread
extern "C" int evrp_encoder_fused_debug(int* out8_host, int reset) {
  int ab = 0;
  if (cudaMemcpyFromSymbol(&ab, evrp::fenc::g_abort, sizeof(int)) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(out8_host, evrp::fenc::g_dbg, sizeof(int) * 8) != cudaSuccess) return -1;
  if (!ab) out8_host[0] = 0;
  if (reset) {
    const int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(evrp::fenc::g_abort, z, sizeof(int));
    cudaMemcpyToSymbol(evrp::fenc::g_dbg, z, sizeof(z));
  }
  return ab;
}
