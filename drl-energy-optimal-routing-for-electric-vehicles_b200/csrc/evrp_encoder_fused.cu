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

#ifdef EVRP_FENC_PROF
// wait-time profile of CTA 0: cycles spent in every wait site (index = code % 100) of the producer lane, the MMA lane
// and worker thread 0, plus their total run time in slot 99
__device__ unsigned long long g_prof[3][100];
#define PROF_DECL long long prof_acc[100]; for (int _i = 0; _i < 100; ++_i) prof_acc[_i] = 0; const long long prof_t0 = clock64(); long long prof_last = prof_t0;
#define PHASE(i) do { const long long _n = clock64(); prof_acc[i] += _n - prof_last; prof_last = _n; } while (0)
#define PROF_DUMP(role) do { if (blockIdx.x == 0) { prof_acc[99] = clock64() - prof_t0; for (int _i = 0; _i < 100; ++_i) g_prof[role][_i] = (unsigned long long)prof_acc[_i]; } } while (0)
#define WAIT(bar, parity, code) do { const long long _t = clock64(); wait_bar(bar, parity, code); prof_acc[(code) % 100] += clock64() - _t; } while (0)
#else
#define PROF_DECL
#define PHASE(i) do { } while (0)
#define PROF_DUMP(role) do { } while (0)
#define WAIT(bar, parity, code) wait_bar(bar, parity, code)
#endif

__device__ __forceinline__ uint32_t try_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok;
}
// slow path, out of line (one copy): keeps the ~40 wait sites small -- the kernel's three roles run three different
// instruction streams, so code size is instruction-cache pressure
__device__ __noinline__ void wait_slow(uint32_t addr, uint32_t parity, int code) {
  long long t0 = 0;
  for (uint32_t spin = 1;; ++spin) {
    if (try_wait(addr, parity)) return;
    if ((spin & 255u) == 0u) {
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
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity, int code) {
  const uint32_t addr = smem_u32(bar);
  if (try_wait(addr, parity)) return;
  wait_slow(addr, parity, code);
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void col_group_sync(int q) { asm volatile("bar.sync %0, 128;" ::"r"(5 + q) : "memory"); }
// bulk tensor store shared -> global (128-byte swizzled [rows x 32 fp32] box), tracked by the issuing thread's bulk group
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void row_group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

// packed fp32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100: one issue slot for two IEEE fp32 operations)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { return ((uint64_t)__float_as_uint(hi) << 32) | __float_as_uint(lo); }
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ float lo_of(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_of(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// one fp32 pair -> packed fp16 (hi, lo) planes: x = hi + lo to 22 significant bits (5 instructions per pair)
__device__ __forceinline__ void split_pair(uint64_t x, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(lo_of(x), hi_of(x));
  const float2 hf = __half22float2(h);
  const uint64_t d = sub2(x, pk2(hf.x, hf.y));
  const __half2 l = __floats2half2_rn(lo_of(d), hi_of(d));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 16 fp32 pairs -> 16 + 16 packed fp16 (hi, lo) words, element 2e in the low half (= the lower K index)
__device__ __forceinline__ void split32(const uint64_t (&v)[16], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int e = 0; e < 16; ++e) split_pair(v[e], hi[e], lo[e]);
}
__device__ __forceinline__ uint64_t join_pair(uint32_t hi, uint32_t lo) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lo));
  return add2(pk2(a.x, a.y), pk2(b.x, b.y));
}

template <int V> struct IC { static constexpr int value = V; };

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
encoder_fused_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapEmb,
                     const __grid_constant__ CUtensorMap mapKvl, const __grid_constant__ CUtensorMap mapTab, const Args a) {
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
      PROF_DECL
      uint32_t slot = 0, phase = 1;            // a fresh barrier passes a wait on the phase before its first
#pragma unroll 1
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int i = 0; i < tile_blocks; ++i) {
          WAIT(&bars->empty[slot], phase, 100);
          mbar_expect_tx(&bars->full[slot], BLK);
          tma_load_2d(ring + slot * BLK, &mapW, &bars->full[slot], 0, i * 128);
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
        }
      }
      PROF_DUMP(0);
    }
  } else if (warp == NWORK) {
    // =========================== MMA issuer ===========================
    // The whole warp runs the (warp-uniform) program so that descriptors and addresses live in uniform registers; one
    // elected lane issues the tcgen05 instructions (a divergent `if (lane == 0)` costs an election loop per MMA, and the
    // single issuing warp is instruction-latency bound: every descriptor below is one add away from a base).
    {
      const bool leader = elect_one();
      PROF_DECL
      uint32_t slot = 0, phase = 0;            // ring position
      uint32_t n_xp = 0, n_qo = 0, n_kv = 0, n_hp = 0;
      uint32_t n_acc0 = 0, n_acc1 = 0;         // products written to each accumulator so far
      const int nks = (rows_per_tile + 15) >> 4;          // 16-key steps of the P V product
      const uint64_t ring_desc = umma_desc(smem_u32(ring));               // + 1024 per slot, + 2 per 16-element K step
      const uint64_t kp_desc = umma_desc(smem_u32(KP)), vt_desc = umma_desc(smem_u32(VT));
      constexpr uint64_t DBLK = BLK >> 4;
      // D[128 x 128] (+)= A[128 x 128] W^T: A planes in TMEM (hi at a_tmem, lo at a_tmem + 64), four ring blocks
      // (K block 0 hi, lo, K block 1 hi, lo)
      auto gemm_k128 = [&](uint32_t d_tmem, uint32_t a_tmem, bool accumulate) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t s0 = slot, p0 = phase;
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
          const uint32_t s1 = slot, p1 = phase;
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
          WAIT(&bars->full[s0], p0, 200);
          WAIT(&bars->full[s1], p1, 210);
          tc_fence_after();
          const uint64_t dwh = ring_desc + s0 * DBLK, dwl = ring_desc + s1 * DBLK;
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t ah = a_tmem + (kb * 4 + k) * 8, al = ah + 64;
              umma_f16_ts(d_tmem, ah, dwh + 2 * k, (accumulate || kb || k) ? 1u : 0u, IDESC_G);
              umma_f16_ts(d_tmem, al, dwh + 2 * k, 1u, IDESC_G);
              umma_f16_ts(d_tmem, ah, dwl + 2 * k, 1u, IDESC_G);
            }
            umma_commit(&bars->empty[s0]);
            umma_commit(&bars->empty[s1]);
          }
        }
      };
      // the previous product in an accumulator has been read by the workers
      auto acquire_acc0 = [&]() { if (n_acc0 > 0) { WAIT(&bars->acc_empty[0], (n_acc0 - 1) & 1, 220); tc_fence_after(); } ++n_acc0; };
      auto acquire_acc1 = [&]() { if (n_acc1 > 0) { WAIT(&bars->acc_empty[1], (n_acc1 - 1) & 1, 221); tc_fence_after(); } ++n_acc1; };
      const uint32_t tX = tmem_base + TM_X, tS = tmem_base + TM_S, tP = tmem_base + TM_P, tQO = tmem_base + TM_QO;
#pragma unroll 1
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        WAIT(&bars->xp_full, n_xp++ & 1, 230); tc_fence_after();
        PHASE(45);
#pragma unroll 1
        for (int l = 0; l < a.n_layers; ++l) {
          // ---- Q | K | V = x W_qkv^T, three 128-column chunks into accumulators 0, 1, 0 ----
          acquire_acc0(); gemm_k128(tS, tX, false); if (leader) umma_commit(&bars->acc_full[0]);
          acquire_acc1(); gemm_k128(tP, tX, false); if (leader) umma_commit(&bars->acc_full[1]);
          acquire_acc0(); gemm_k128(tS, tX, false); if (leader) umma_commit(&bars->acc_full[0]);
          PHASE(40);
          // ---- attention: scores of head h + 1 are issued before the P V product of head h ----
          WAIT(&bars->qo_ready, n_qo++ & 1, 231);
          WAIT(&bars->kv_ready, n_kv++ & 1, 232);
          tc_fence_after();
          auto issue_pv = [&](int h) {
            WAIT(&bars->hp_full, n_hp++ & 1, 233); tc_fence_after();
            if (leader) {
              const uint64_t dv = vt_desc + (uint32_t)h * 128u;               // V^T rows 16h.. : 16 x 128 B
              const uint32_t d_o = tQO + h * 16;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                if (ks < nks) {
                  const uint64_t dvh = dv + (ks >> 2) * (2 * DBLK) + (ks & 3) * 2, dvl = dvh + DBLK;
                  const uint32_t ph = tP + ks * 8, pl = ph + 64;
                  umma_f16_ts(d_o, ph, dvh, ks ? 1u : 0u, IDESC_PV);
                  umma_f16_ts(d_o, pl, dvh, 1u, IDESC_PV);
                  umma_f16_ts(d_o, ph, dvl, 1u, IDESC_PV);
                }
              }
              umma_commit(&bars->hp_empty);
            }
          };
#pragma unroll 1
          for (int h = 0; h < H; ++h) {
            acquire_acc0();
            if (leader) {
              const uint64_t dkh = kp_desc + (uint32_t)(h >> 2) * (2 * DBLK) + (uint32_t)(h & 3) * 2, dkl = dkh + DBLK;
              const uint32_t qh = tQO + h * 16, ql = qh + 8;
              umma_f16_ts(tS, qh, dkh, 0u, IDESC_G);
              umma_f16_ts(tS, ql, dkh, 1u, IDESC_G);
              umma_f16_ts(tS, qh, dkl, 1u, IDESC_G);
              umma_commit(&bars->acc_full[0]);
            }
            if (h > 0) issue_pv(h - 1);
          }
          issue_pv(H - 1);
          if (leader) umma_commit(&bars->y_full);
          PHASE(41);
          // ---- heads W_out^T ----
          WAIT(&bars->qo_ready, n_qo++ & 1, 234); tc_fence_after();
          acquire_acc0();
          gemm_k128(tS, tQO, false);
          if (leader) umma_commit(&bars->acc_full[0]);
          PHASE(42);
          // ---- FF: hidden chunk j+1 (FF1) is issued before the FF2 partial product over chunk j ----
          WAIT(&bars->xp_full, n_xp++ & 1, 235); tc_fence_after();
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            acquire_acc0();
            gemm_k128(tS, tX, false);
            if (leader) umma_commit(&bars->acc_full[0]);
            if (j > 0) {
              WAIT(&bars->hp_full, n_hp++ & 1, 236); tc_fence_after();
              gemm_k128(tQO, tP, j > 1);
              if (leader) umma_commit(&bars->hp_empty);
            }
          }
          WAIT(&bars->hp_full, n_hp++ & 1, 237); tc_fence_after();
          gemm_k128(tQO, tP, true);
          if (leader) { umma_commit(&bars->hp_empty); umma_commit(&bars->y_full); }
          WAIT(&bars->xp_full, n_xp++ & 1, 238); tc_fence_after();      // next layer's input / the embeddings
          PHASE(43);
        }
        // ---- node projection (glimpse K | glimpse V | folded logit K [| AM step table]) ----
#pragma unroll 1
        for (int c = 0; c < a.n_node_chunks; ++c) {
          if (c & 1) { acquire_acc1(); gemm_k128(tP, tX, false); if (leader) umma_commit(&bars->acc_full[1]); }
          else { acquire_acc0(); gemm_k128(tS, tX, false); if (leader) umma_commit(&bars->acc_full[0]); }
        }
        PHASE(44);
      }
      if (leader) PROF_DUMP(1);
    }
  } else {
    // =========================== workers: thread = (row r of the tile = TMEM lane, column quarter q) ===========================
    const int g = warp & 3, q = warp >> 2;
    const int r = g * 32 + lane;
    const uint32_t lane_base = (uint32_t)(g * 32) << 16;
    const uint32_t tX = tmem_base + lane_base + TM_X, tS = tmem_base + lane_base + TM_S,
                   tP = tmem_base + lane_base + TM_P, tQO = tmem_base + lane_base + TM_QO;
    const float c2 = 0.25f * 1.4426950408889634f;         // norm_factor 1/sqrt(16) (nets/GraphEncoder.py:38) and log2(e)
    uint32_t n_accf0 = 0, n_accf1 = 0, n_hp = 0, n_y = 0;
    PROF_DECL
    // keys this thread's row attends to: the rows of its own instance
    uint32_t kmask = 0;
    {
      int gi = r / S;
      if (gi >= inst_per_tile) gi = inst_per_tile - 1;    // padding rows: any valid key range (their outputs are dropped)
      const int klo = gi * S, khi = klo + S;
#pragma unroll
      for (int e = 0; e < 32; ++e) kmask |= ((32 * q + e >= klo && 32 * q + e < khi) ? 1u : 0u) << e;
    }
    const bool all_keys = __all_sync(0xffffffffu, kmask == 0xffffffffu);      // warp-uniform: no masking needed
    // blocks of 16 keys (32q.., 32q + 16..) that some row of this warp attends to (warp-uniform): n_live of them, the
    // first at block blk0; kmask_live = the thread's key mask from that block on
    const uint32_t half_live = (__any_sync(0xffffffffu, (kmask & 0xffffu) != 0u) ? 1u : 0u) |
                               (__any_sync(0xffffffffu, (kmask >> 16) != 0u) ? 2u : 0u);
    const int n_live = (half_live & 1u) + (half_live >> 1), blk0 = half_live == 2u ? 1 : 0;
    const uint32_t kmask_live = kmask >> (16 * blk0);
    auto wait_acc0 = [&](int code) { WAIT(&bars->acc_full[0], n_accf0++ & 1, code); tc_fence_after(); };
    auto wait_acc1 = [&](int code) { WAIT(&bars->acc_full[1], n_accf1++ & 1, code); tc_fence_after(); };
    auto release_acc = [&](int b) { tmem_ld_wait(); tc_fence_before(); warp_arrive(&bars->acc_empty[b], lane); };
    auto wait_p_free = [&](int code) {       // the MMAs that read the previous contents of the P region have retired
      if (n_hp > 0) { WAIT(&bars->hp_empty, (n_hp - 1) & 1, code); tc_fence_after(); }
      ++n_hp;
    };
    // 16 fp32 pairs -> (hi, lo) planes at packed columns 16q.. / 64 + 16q.. of a 128-column operand region
    auto store_planes = [&](uint32_t region, const uint64_t (&v)[16]) {
      uint32_t hi[16], lo[16];
      split32(v, hi, lo);
      tmem_st16(region + 16 * q, hi);
      tmem_st16(region + 64 + 16 * q, lo);
      tmem_st_wait(); tc_fence_before();
    };

    // Output tiles leave through shared memory: the K / V^T plane regions (dead outside the attention phase) alternate
    // as staging buffers of four [rows x 32 fp32] 128-byte-swizzled boxes; the four warps that own a column quarter
    // fill box q and one of their threads hands it to TMA (a direct store would touch 32 separate lines per warp
    // instruction).  A buffer is rewritten only after the bulk store issued two outputs earlier has read it.
    uint32_t n_out = 0;
    const bool out_leader = (g == 0 && lane == 0);
    auto stage_out = [&](const CUtensorMap* map, int col0, int row0, const uint32_t (&u)[32], float scale) {
      unsigned char* box = (n_out & 1 ? VT : KP) + q * BLK;
      ++n_out;
      if (out_leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      col_group_sync(q);
      unsigned char* row = box + r * 128;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(row + ((i ^ (r & 7)) << 4)) =
            make_float4(__uint_as_float(u[4 * i]) * scale, __uint_as_float(u[4 * i + 1]) * scale,
                        __uint_as_float(u[4 * i + 2]) * scale, __uint_as_float(u[4 * i + 3]) * scale);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      col_group_sync(q);
      if (out_leader) tma_store_2d(map, box, col0 + 32 * q, row0);
    };

#pragma unroll 1
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int grow = tile * rows_per_tile + r;
      const bool valid = r < rows_per_tile && grow < M;
      // ---- init embedding (nets/DRLModel.py:192-200): depot | stations | customers (x/100, y/100[, demand]) ----
      {
        uint64_t v[16];
        if (valid) {
          const int b = grow / S, j = grow - b * S;
          const float x = __fdiv_rn(__ldg(a.xy + ((size_t)b * 2 + 0) * S + j), 100.f);
          const float y = __fdiv_rn(__ldg(a.xy + ((size_t)b * 2 + 1) * S + j), 100.f);
          if (j <= a.F) {
            const float* w = (j == 0 ? a.depot_w : a.station_w) + 64 * q;
            const float* bb = (j == 0 ? a.depot_b : a.station_b) + 32 * q;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float4 ww = __ldg(reinterpret_cast<const float4*>(w + 4 * i));
              const float2 b2 = __ldg(reinterpret_cast<const float2*>(bb + 2 * i));
              v[i] = pk2(fmaf(ww.y, y, fmaf(ww.x, x, 0.f)) + b2.x, fmaf(ww.w, y, fmaf(ww.z, x, 0.f)) + b2.y);
            }
          } else {
            const float dem = __ldg(a.dyn + ((size_t)b * 4 + 1) * S + j);
            const float* w = a.cust_w + 96 * q;
            const float* bb = a.cust_b + 32 * q;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 w0 = __ldg(reinterpret_cast<const float2*>(w + 6 * i)), w1 = __ldg(reinterpret_cast<const float2*>(w + 6 * i + 2)),
                           w2 = __ldg(reinterpret_cast<const float2*>(w + 6 * i + 4));
              const float2 b2 = __ldg(reinterpret_cast<const float2*>(bb + 2 * i));
              v[i] = pk2(fmaf(w1.x, dem, fmaf(w0.y, y, fmaf(w0.x, x, 0.f))) + b2.x, fmaf(w2.y, dem, fmaf(w2.x, y, fmaf(w1.y, x, 0.f))) + b2.y);
            }
          }
          if (a.n_layers == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 2)
              *reinterpret_cast<float4*>(a.emb + (size_t)grow * D + 32 * q + 2 * i) = make_float4(lo_of(v[i]), hi_of(v[i]), lo_of(v[i + 1]), hi_of(v[i + 1]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0ull;
        }
        store_planes(tX, v);
        if (out_leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // K / V^T regions free again
        warp_arrive(&bars->xp_full, lane);
      }
      PHASE(90);

#pragma unroll 1
      for (int l = 0; l < a.n_layers; ++l) {
        const float* prm = a.params + (size_t)l * LP;
        const float inv_qkv = __ldg(prm + P_INV), inv_wo = __ldg(prm + P_INV + 1), inv_w1 = __ldg(prm + P_INV + 2),
                    inv_w2 = __ldg(prm + P_INV + 3);
        // ---- Q chunk -> operand planes per head in QO: head 2q at columns 32q (hi 8 | lo 8), head 2q+1 at 32q + 16 ----
        {
          uint32_t u[32];
          wait_acc0(300);
          tmem_ld32(tS + 32 * q, u);
          release_acc(0);
          const uint64_t sc = pk2(inv_qkv, inv_qkv);
          uint32_t o[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            uint32_t hi, lo;
            split_pair(mul2(pk2u(u[2 * i], u[2 * i + 1]), sc), hi, lo);
            o[(i >> 3) * 16 + (i & 7)] = hi;
            o[(i >> 3) * 16 + 8 + (i & 7)] = lo;
          }
          tmem_st32(tQO + 32 * q, o);
          tmem_st_wait(); tc_fence_before();
          warp_arrive(&bars->qo_ready, lane);
        }
        // ---- K chunk -> K planes in shared memory: row = key r, channels 32q..32q+31 = 4 swizzled 16-byte chunks ----
        {
          uint32_t u[32];
          wait_acc1(301);
          tmem_ld32(tP + 32 * q, u);
          release_acc(1);
          const uint64_t sc = pk2(inv_qkv, inv_qkv);
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) split_pair(mul2(pk2u(u[2 * i], u[2 * i + 1]), sc), hi[i], lo[i]);
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
          wait_acc0(302);
          tmem_ld32(tS + 32 * q, u);
          release_acc(0);
          const uint64_t sc = pk2(inv_qkv, inv_qkv);
          unsigned char* vb = VT + (r >> 6) * (2 * BLK) + (32 * q) * 128 + ((r & 7) << 1);
          const int kc = (r & 63) >> 3;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            uint32_t hi, lo;
            split_pair(mul2(pk2u(u[2 * i], u[2 * i + 1]), sc), hi, lo);
            unsigned char* p0 = vb + (2 * i) * 128 + ((kc ^ ((2 * i) & 7)) << 4);
            unsigned char* p1 = vb + (2 * i + 1) * 128 + ((kc ^ ((2 * i + 1) & 7)) << 4);
            *reinterpret_cast<unsigned short*>(p0) = (unsigned short)(hi & 0xffffu);
            *reinterpret_cast<unsigned short*>(p1) = (unsigned short)(hi >> 16);
            *reinterpret_cast<unsigned short*>(p0 + BLK) = (unsigned short)(lo & 0xffffu);
            *reinterpret_cast<unsigned short*>(p1 + BLK) = (unsigned short)(lo >> 16);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // K and V^T writes -> visible to tcgen05
          warp_arrive(&bars->kv_ready, lane);
        }
        PHASE(91);
        // ---- softmax of every head: p = 2^(c2 (s - max)) unnormalised, masked keys 0; planes to the P region ----
        // Of the thread's two blocks of 16 keys only those that some row of the warp attends to are processed (keys
        // beyond the graph and the other instances of a packed tile cost nothing): NL = 2, 1 or 0 live blocks.
        float tot0 = 1.f, tot1 = 1.f;                     // softmax denominators of this thread's heads 2q, 2q + 1
        auto softmax_head = [&](int h, auto nl_tag) {
          constexpr int NL = decltype(nl_tag)::value;
          constexpr int NE = NL == 2 ? 32 : 16;
          const int par = h & 1;
          uint32_t u[NE];
          wait_acc0(310);
          if (NL == 2) tmem_ld32(tS + 32 * q, *reinterpret_cast<uint32_t (*)[32]>(&u));
          else if (NL == 1) tmem_ld16(tS + 32 * q + 16 * blk0, *reinterpret_cast<uint32_t (*)[16]>(&u));
          release_acc(0);
          float mx = -INFINITY;
          if (NL > 0) {
            if (!all_keys) {
#pragma unroll
              for (int e = 0; e < NE; ++e) u[e] = ((kmask_live >> e) & 1u) ? u[e] : 0xff800000u;      // -inf: p = 0
            }
            float m0 = fmaxf(__uint_as_float(u[0]), __uint_as_float(u[1])), m1 = fmaxf(__uint_as_float(u[2]), __uint_as_float(u[3]));
#pragma unroll
            for (int e = 4; e < NE; e += 4) {
              m0 = fmaxf(m0, fmaxf(__uint_as_float(u[e]), __uint_as_float(u[e + 1])));
              m1 = fmaxf(m1, fmaxf(__uint_as_float(u[e + 2]), __uint_as_float(u[e + 3])));
            }
            mx = fmaxf(m0, m1);
          }
          pmax[(par * 4 + q) * 128 + r] = mx;
          row_group_sync(g);
          mx = fmaxf(fmaxf(pmax[(par * 4 + 0) * 128 + r], pmax[(par * 4 + 1) * 128 + r]),
                     fmaxf(pmax[(par * 4 + 2) * 128 + r], pmax[(par * 4 + 3) * 128 + r]));
          if (h > 0) {                                    // the previous head's denominator (written after its barrier)
            const float* ps = psum + ((par ^ 1) * 4) * 128 + r;
            const float t = (ps[0] + ps[128]) + (ps[256] + ps[384]);
            if (((h - 1) >> 1) == q) { if ((h - 1) & 1) tot1 = t; else tot0 = t; }
          }
          const float nmx2 = -mx * c2;
          const uint64_t cc = pk2(c2, c2), mm = pk2(nmx2, nmx2);
          uint32_t hi[NE / 2], lo[NE / 2];
          uint64_t sum0 = 0ull, sum1 = 0ull;
          if (NL > 0) {
#pragma unroll
            for (int hb = 0; hb < NE / 16; ++hb) {          // per block of 16: the exponentials (MUFU), then the split (FMA / ALU)
              uint64_t p[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint64_t x = fma2(pk2u(u[16 * hb + 2 * e], u[16 * hb + 2 * e + 1]), cc, mm);
                float a0, a1;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a0) : "f"(lo_of(x)));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(a1) : "f"(hi_of(x)));
                p[e] = pk2(a0, a1);
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                if (e & 1) sum1 = add2(sum1, p[e]); else sum0 = add2(sum0, p[e]);
                split_pair(p[e], hi[8 * hb + e], lo[8 * hb + e]);
              }
            }
          }
          const uint64_t sm = add2(sum0, sum1);
          const float sum = lo_of(sm) + hi_of(sm);
          wait_p_free(320);
          if (NL == 2) {
            tmem_st16(tP + 16 * q, *reinterpret_cast<const uint32_t (*)[16]>(&hi));
            tmem_st16(tP + 64 + 16 * q, *reinterpret_cast<const uint32_t (*)[16]>(&lo));
          } else {
            const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if (NL == 1) {
              tmem_st8(tP + 16 * q + 8 * blk0, *reinterpret_cast<const uint32_t (*)[8]>(&hi));
              tmem_st8(tP + 64 + 16 * q + 8 * blk0, *reinterpret_cast<const uint32_t (*)[8]>(&lo));
              tmem_st8(tP + 16 * q + 8 * (blk0 ^ 1), z);
              tmem_st8(tP + 64 + 16 * q + 8 * (blk0 ^ 1), z);
            } else {
              tmem_st8(tP + 16 * q, z); tmem_st8(tP + 16 * q + 8, z);
              tmem_st8(tP + 64 + 16 * q, z); tmem_st8(tP + 64 + 16 * q + 8, z);
            }
          }
          tmem_st_wait(); tc_fence_before();
          warp_arrive(&bars->hp_full, lane);
          psum[(par * 4 + q) * 128 + r] = sum;
        };
#pragma unroll 1
        for (int h = 0; h < H; ++h) {
          if (n_live == 2) softmax_head(h, IC<2>{});
          else if (n_live == 1) softmax_head(h, IC<1>{});
          else softmax_head(h, IC<0>{});
        }
        PHASE(92);
        // ---- concatenated heads: O / denominator -> operand planes, in place in QO ----
        {
          row_group_sync(g);                              // the last head's partial denominators are visible
          {
            const float* ps = psum + (((H - 1) & 1) * 4) * 128 + r;
            const float t = (ps[0] + ps[128]) + (ps[256] + ps[384]);
            if (((H - 1) >> 1) == q) tot1 = t;
          }
          WAIT(&bars->y_full, n_y++ & 1, 330); tc_fence_after();
          uint32_t u[32];
          tmem_ld32(tQO + 32 * q, u);
          tmem_ld_wait();
          const float i0 = 1.f / tot0, i1 = 1.f / tot1;
          const uint64_t s0 = pk2(i0, i0), s1 = pk2(i1, i1);
          uint64_t v[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) { v[i] = mul2(pk2u(u[2 * i], u[2 * i + 1]), s0); v[8 + i] = mul2(pk2u(u[16 + 2 * i], u[17 + 2 * i]), s1); }
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
        PHASE(93);
        // ---- x1 = BN1(x + heads W_out^T)  (nets/GraphEncoder.py:162-169): planes rewritten in place ----
        {
          uint32_t u[32], xh[16], xl[16];
          wait_acc0(340);
          tmem_ld32(tS + 32 * q, u);
          tmem_ld16(tX + 16 * q, xh);
          tmem_ld16(tX + 64 + 16 * q, xl);
          release_acc(0);
          const uint64_t iw = pk2(inv_wo, inv_wo);
          uint64_t v[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(prm + P_BN1S + 32 * q + 2 * i));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(prm + P_BN1B + 32 * q + 2 * i));
            v[i] = fma2(fma2(pk2u(u[2 * i], u[2 * i + 1]), iw, join_pair(xh[i], xl[i])), pk2(sc.x, sc.y), pk2(sh.x, sh.y));
            v[i + 1] = fma2(fma2(pk2u(u[2 * i + 2], u[2 * i + 3]), iw, join_pair(xh[i + 1], xl[i + 1])), pk2(sc.z, sc.w), pk2(sh.z, sh.w));
          }
          store_planes(tX, v);
          warp_arrive(&bars->xp_full, lane);
        }
        PHASE(94);
        // ---- hidden chunk j = relu(x1 W1_j^T + b1_j) -> operand planes in the P region ----
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          uint32_t u[32];
          wait_acc0(350);
          tmem_ld32(tS + 32 * q, u);
          release_acc(0);
          const uint64_t iw = pk2(inv_w1, inv_w1);
          uint64_t v[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(prm + P_B1 + 128 * j + 32 * q + 2 * i));
            const uint64_t t0 = fma2(pk2u(u[2 * i], u[2 * i + 1]), iw, pk2(bb.x, bb.y));
            const uint64_t t1 = fma2(pk2u(u[2 * i + 2], u[2 * i + 3]), iw, pk2(bb.z, bb.w));
            v[i] = pk2(fmaxf(lo_of(t0), 0.f), fmaxf(hi_of(t0), 0.f));
            v[i + 1] = pk2(fmaxf(lo_of(t1), 0.f), fmaxf(hi_of(t1), 0.f));
          }
          uint32_t hi[16], lo[16];
          split32(v, hi, lo);
          wait_p_free(360);
          tmem_st16(tP + 16 * q, hi);
          tmem_st16(tP + 64 + 16 * q, lo);
          tmem_st_wait(); tc_fence_before();
          warp_arrive(&bars->hp_full, lane);
        }
        PHASE(95);
        // ---- x2 = BN2(x1 + hidden W2^T + b2): the next layer's input; the last layer's is the node embedding ----
        {
          uint32_t u[32], xh[16], xl[16];
          WAIT(&bars->y_full, n_y++ & 1, 370); tc_fence_after();
          tmem_ld32(tQO + 32 * q, u);
          tmem_ld16(tX + 16 * q, xh);
          tmem_ld16(tX + 64 + 16 * q, xl);
          tmem_ld_wait();
          const uint64_t iw = pk2(inv_w2, inv_w2);
          uint64_t v[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(prm + P_B2 + 32 * q + 2 * i));
            const float4 sc = __ldg(reinterpret_cast<const float4*>(prm + P_BN2S + 32 * q + 2 * i));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(prm + P_BN2B + 32 * q + 2 * i));
            v[i] = fma2(add2(fma2(pk2u(u[2 * i], u[2 * i + 1]), iw, pk2(bb.x, bb.y)), join_pair(xh[i], xl[i])), pk2(sc.x, sc.y), pk2(sh.x, sh.y));
            v[i + 1] = fma2(add2(fma2(pk2u(u[2 * i + 2], u[2 * i + 3]), iw, pk2(bb.z, bb.w)), join_pair(xh[i + 1], xl[i + 1])), pk2(sc.z, sc.w), pk2(sh.z, sh.w));
          }
          store_planes(tX, v);
          warp_arrive(&bars->xp_full, lane);
          if (l == a.n_layers - 1) {                      // the node embeddings (after the hand-off: the MMAs run meanwhile)
            uint32_t w[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) { w[2 * i] = (uint32_t)v[i]; w[2 * i + 1] = (uint32_t)(v[i] >> 32); }
            stage_out(&mapEmb, 0, tile * rows_per_tile, w, 1.f);
          }
        }
        PHASE(96);
      }
      // ---- node projection chunks -> kvl (chunks 0..2) / step table (chunk 3) ----
      {
        const float inv_node = __ldg(a.params + (size_t)a.n_layers * LP);
#pragma unroll 1
        for (int c = 0; c < a.n_node_chunks; ++c) {
          const int b = c & 1;
          uint32_t u[32];
          if (b) wait_acc1(381); else wait_acc0(380);
          tmem_ld32((b ? tP : tS) + 32 * q, u);
          release_acc(b);
          if (c < 3) stage_out(&mapKvl, 128 * c, tile * rows_per_tile, u, inv_node);
          else stage_out(&mapTab, 0, tile * rows_per_tile, u, inv_node);
        }
      }
      PHASE(97);
    }
    if (out_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef EVRP_FENC_PROF
    if (threadIdx.x == 0) PROF_DUMP(2);
#endif
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
  CUtensorMap map, mapEmb, mapKvl, mapTab;
  int rc;
  if ((rc = tc::make_weight_map_f16(&map, w->fused_w, pack_blocks * 128, 64, 128))) return rc;
  const int ipt0 = S > 64 ? 1 : 128 / S;
  // output maps: box = [32 fp32 x rows of one tile], 128-byte swizzle; rows beyond B*S are clipped by the TMA unit
  if ((rc = tc::make_weight_map(&mapEmb, emb, B * S, 128, 128, ipt0 * S))) return rc;
  if ((rc = tc::make_weight_map(&mapKvl, kvl, B * S, 384, 384, ipt0 * S))) return rc;
  if ((rc = tc::make_weight_map(&mapTab, step_tab ? step_tab : emb, B * S, 128, 128, ipt0 * S))) return rc;
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
  encoder_fused_kernel<<<grid, NT, SMEM_BYTES, st>>>(map, mapEmb, mapKvl, mapTab, a);
  EVRP_LAUNCH_OK();
  return 0;
}

}  // namespace evrp

#ifdef EVRP_FENC_PROF
extern "C" int evrp_encoder_fused_prof(unsigned long long* out300_host) {
  return cudaMemcpyFromSymbol(out300_host, evrp::fenc::g_prof, sizeof(unsigned long long) * 300) == cudaSuccess ? 0 : -1;
}
#endif

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
