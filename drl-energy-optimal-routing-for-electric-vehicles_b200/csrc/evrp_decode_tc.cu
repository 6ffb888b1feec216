// Persistent rollout kernel, tensor-core variant: ONE launch decodes whole tours and the
// per-step route-context MLP (FF_tour, nets/DRLModel.py:203-237) runs on tcgen05 tensor cores inside it.
//
//   AttentionModel._inner            nets/DRLModel.py:240-280   (the step loop lives inside the kernel)
//   route_feature                    nets/DRLModel.py:203-237   (running max + FF_tour)
//   _get_log_p / _one_to_many_logits nets/DRLModel.py:379-452   (8-head masked glimpse, pointer logits)
//   _select_node                     nets/DRLModel.py:316-342   (greedy first-index argmax / sampling)
//   _calc_log_likelihood             nets/DRLModel.py:182-190
//   update_dynamic / update_mask     problems/EVRP.py:105-280   (fused, bit-exact recipe of evrp_common.cuh)
//
// Layout: ONE CTA per SM (640 threads) owns up to RPC = 64 rollout rows for their whole life.
//   warps 0-15   row warps: per-row decode phase; in the FF phase they are the TMEM epilogue / operand converters
//   warp  16     TMA producer: streams the pre-split fp16 weight planes of FF_tour from L2 (3-slot ring, 32 KB stages)
//   warp  17     MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = 64, K = 16)
//   warps 18-19  walk rows only (idle in the FF phase): 56 rows per CTA at batch 8192 = at most 3 rows per warp
// Per step
//   1. FF_tour, transposed so that the ROWS are the MMA N dimension (any row count <= 64 costs N/256 of a full tile):
//        h1^T [512 x rows] = W1max [512 x 128] . m^T      4 M-tiles, A = weight tile (smem, TMA), B = running-max planes
//        ctx^T [128 x rows] = W2 [128 x 512] . relu(h1 + b1 + W1veh veh)^T     B = hidden planes written by the epilogue
//      every operand is split into two fp16 planes, x = hi + lo (22 significant bits, like a TF32 pair at half the
//      bytes; weights pre-scaled by a power of two so that lo stays normal) and every product is the compensated sum
//      Ah Bh + Al Bh + Ah Bl with fp32 accumulation in TMEM - the fp32-faithful scheme of the encoder GEMMs
//      (evrp_tc_gemm.cu) at half the weight stream and half the MMA count; accumulators: 4 x 64 + 64 TMEM columns.
//   2. per row (warp): glimpse attention / logits over the FEASIBLE nodes only, warp-level masked log-softmax,
//      first-index argmax / inverse-CDF sample, state transition + look-ahead feasibility mask (bit-exact recipe).
//      The running max of visited embeddings is kept exact in global memory (L2-resident) and as fp16 (hi, lo)
//      planes in the 128B-swizzled K-major layout the MMA reads.
// Rows retire individually; the host reads back only t_max and the status word.
#include <math.h>
#include <stdlib.h>
#include <type_traits>
#include <cuda_fp16.h>
#include "evrp_tc.cuh"

namespace evrp {
namespace dtc {

using namespace tc;

#ifndef EVRP_RPC
#define EVRP_RPC 64          // rollout rows per CTA = MMA N
#define EVRP_ROW_WARPS 16
#define EVRP_NB2 2           // hidden-plane staging buffers
#define EVRP_CTAS 1          // CTAs per SM
#define EVRP_TMEM_COLS 512
#endif
#ifndef EVRP_EXTRA_WALKERS
#define EVRP_EXTRA_WALKERS 2 // warps that only walk rows (idle in the FF phase): 56 rows per CTA over 20 walkers = at most 3 rows per warp
#endif
constexpr int RPC = EVRP_RPC;
constexpr int ROW_WARPS = EVRP_ROW_WARPS;
constexpr int NB2 = EVRP_NB2;
static_assert(RPC == 4 * ROW_WARPS, "epilogue mapping: 16 rows per (quadrant, column group) warp");
constexpr int XW = EVRP_EXTRA_WALKERS;
constexpr int NWALK = ROW_WARPS + 2 + XW;    // row warps + the TMA / MMA warps + walker-only warps
constexpr int NT = NWALK * 32;               // 640
constexpr int SP = EVRP_MAX_S;               // padded node count (128)
constexpr int PSTR = SP + 4;                 // head stride of the compat buffer
constexpr int KBLK = RPC * 128;              // 8 KB: [64 rows x 64 fp16], 128B-swizzled, one plane of one K block
constexpr int ACT_BYTES = 2 * 2 * KBLK;      // 32 KB: 2 K blocks (128 channels) x (hi, lo)
constexpr int WPLANE = 128 * 128;            // 16 KB: [128 x 64 fp16] weight tile plane
constexpr int WSTAGE = 2 * WPLANE;           // hi + lo
#ifndef EVRP_WSTAGES
#define EVRP_WSTAGES 3
#endif
constexpr int WSTAGES = EVRP_WSTAGES;
constexpr int NSTAGE_STEP = 16;              // stages per decode step: 8 of W1max (4 tiles x 2 K blocks), 8 of W2
constexpr uint32_t IDESC = make_idesc_f16(128, RPC);
constexpr int TM_D2 = 4 * RPC;               // TMEM column of the ctx accumulator (after the 4 hidden tiles)

struct RowState {
  float load, soc, tim, d4;
  int now, done, inst, valid;
  uint32_t mk[4];
};

struct WarpScratch {
  float pc[H][PSTR];             // compat -> attention probabilities, by list position
  float logit[SP];               // clipped logits / temp, by list position
  float heads[D];                // glimpse heads (concat)
  unsigned char list[SP];        // feasible node indices, ascending
};

constexpr int SCRATCH_BYTES = NWALK * (int)sizeof(WarpScratch);   // the two FF-phase warps and the extra warps walk rows as well
constexpr int RING_BYTES = WSTAGES * WSTAGE;
constexpr int RS_BYTES = ((SCRATCH_BYTES > RING_BYTES ? SCRATCH_BYTES : RING_BYTES) + 1023) / 1024 * 1024;

struct __align__(16) Ctl {
  RowState rs[RPC];
  float veh[RPC][4];             // load, soc/Start_SOC, time/t_limit          (:215-222)
  uint64_t full[WSTAGES], empty[WSTAGES];
  uint64_t d1_full[4];           // hidden tile j accumulated (tcgen05.commit)
  uint64_t b2_full[2];           // hidden tile planes (double-buffered) written by the 16 epilogue warps
  uint64_t b2_empty[2];          // GEMM2 partial over that buffer retired (tcgen05.commit)
  uint64_t d2_full;              // ctx accumulated
  uint32_t tmem_base;
  int next_row;
};

constexpr size_t SMEM_BYTES = (1 + NB2) * (size_t)ACT_BYTES + RS_BYTES + sizeof(Ctl) + 1024;

#ifdef EVRP_PHASE_TIMING
// debug build only (tools/phase_timing.py): cycles of warp 0 in [FF hidden tiles, FF ctx + sync, row phase busy, barrier wait]
__device__ unsigned long long g_phase_cycles_tc[4];
__device__ unsigned long long g_row_cycles_tc[8];   // per-pass cycles of warp 0 inside the row phase
__device__ unsigned long long g_ff_cycles_tc[12];   // FF phase of warp 0: [2j] wait for hidden tile j, [2j+1] convert + store, [8] ctx wait, [9] ctx store, [10] sync
#define FT_MARK(i) do { if (tid == 0) { const long long _n = clock64(); ft_acc[i] += _n - ft_t; ft_t = _n; } } while (0)
#define RT_MARK(i) do { if (tid == 0) { const long long _n = clock64(); rt_acc[i] += _n - rt_t; rt_t = _n; } } while (0)
#define PT_MARK(i) do { if (tid == 0) { const long long _n = clock64(); pt_acc[i] += _n - pt_t; pt_t = _n; } } while (0)
#else
#define PT_MARK(i) do { } while (0)
#define RT_MARK(i) do { } while (0)
#define FT_MARK(i) do { } while (0)
#endif

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// Record loads of the row phase carry an L2::evict_first policy: every K / V / logit-K record is consumed exactly once
// per decode step, so marking the line for early eviction keeps L2 for the records that the just-in-time prefetches
// have requested but the passes have not read yet (25.2 -> 22.1 ms at B = 8192, N = 100; profiles/r1_l2_policy.md).
__device__ __forceinline__ float4 ldg4_pol(const float* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// the two warps of a row team (TEAM == 2) meet here; orders their shared-memory traffic
__device__ __forceinline__ void team_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <typename T>
__device__ __forceinline__ T sel4(const T (&a)[4], int i) { return i == 0 ? a[0] : i == 1 ? a[1] : i == 2 ? a[2] : a[3]; }

struct Maps { CUtensorMap w1h, w1l, w2h, w2l; };

// MODE: 0 greedy, 1 sample, 2 teacher-forced; COORDS: geometry recomputed from coordinate / elevation rows;
// TEAM: warps per row in the row phase.  2 (small batches, at most 10 rows per CTA): an owner warp and a helper warp split
// the feasible nodes of the three record passes (K, V, logit-K) by batch parity and meet on a named barrier; one row then
// costs ~2/3 of the single-warp latency.  The partial sums of the even and the odd batches are kept apart and added once
// in BOTH modes, so the results do not depend on the mode (a batch and its shards decode bit-identically).
template <int MODE, bool COORDS, int TEAM>
__global__ void __launch_bounds__(NT, EVRP_CTAS)
rollout_tc_kernel(const __grid_constant__ Maps maps, evrp_phys phys, evrp_decoder_weights w, evrp_rollout_cfg cfg,
                  const float* __restrict__ emb, const float* __restrict__ kvl, const float* __restrict__ fixed_ctx,
                  const float* __restrict__ dynamic0, const float* __restrict__ Dm, const float* __restrict__ Gm,
                  const float* __restrict__ xyp, const float* __restrict__ elevp,
                  float* station_ws, float* dem_ws, float* m_ws, int B, int S, int F, int rows_total, int rows_per_cta,
                  int pf_lines, int walkers, const int64_t* __restrict__ forced, const float* __restrict__ uniforms,
                  int64_t* __restrict__ pi, float* __restrict__ ll, float* __restrict__ reward,
                  float* __restrict__ energy, uint32_t* __restrict__ saved_mask, float* __restrict__ saved_veh,
                  int32_t* __restrict__ steps_used, int32_t* __restrict__ t_max, int32_t* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment (swizzle atoms) by pointer arithmetic on the __shared__ array itself: the compiler keeps the
  // address space, so every access below is an LDS/STS (an integer round trip degrades them to generic LD/ST)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* B1 = smem;                                  // running-max planes   [kb][plane][row][128 B]
  unsigned char* B2 = smem + ACT_BYTES;                      // hidden tile planes   [buf][kb][plane][row][128 B]
  float (*ctx)[D] = reinterpret_cast<float (*)[D]>(B2);      // FF_tour output (aliases B2 once GEMM2 has retired)
  unsigned char* ring = smem + (1 + NB2) * ACT_BYTES;                // weight stages (FF phase) / row scratch (row phase)
  Ctl& sm = *reinterpret_cast<Ctl*>(smem + (1 + NB2) * ACT_BYTES + RS_BYTES);
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int Tcap = cfg.max_steps;
  const int nblk = (S + 31) >> 5;
  const float sqrt_d = sqrtf((float)D);
  const bool fastdiv = (phys.div_mode == 0) && (phys.velocity == 50.f);   // call-free look-ahead arithmetic is exact
  const float w1_inv_scale = w.inv_scales ? __ldg(w.inv_scales) : w.w1_inv_scale;      // device-side packing keeps the scales on the device
  const float w2_inv_scale = w.inv_scales ? __ldg(w.inv_scales + 1) : w.w2_inv_scale;

  if (tid == 0) {
    for (int s = 0; s < WSTAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int j = 0; j < 4; ++j) mbar_init(&sm.d1_full[j], 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&sm.b2_full[b], ROW_WARPS); mbar_init(&sm.b2_empty[b], 1); }
    mbar_init(&sm.d2_full, 1);
    sm.next_row = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == ROW_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(EVRP_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  // ---------------- prologue: per-row state ----------------
  if (wid < ROW_WARPS) {
    for (int lr = wid; lr < RPC; lr += ROW_WARPS) {
      const int row = blockIdx.x * rows_per_cta + lr;
      const bool valid = (lr < rows_per_cta) && (row < rows_total);
      const int inst = valid ? row % B : 0;
      const Geo<COORDS> geo = make_geo<COORDS>(Dm, Gm, xyp, elevp, inst, S);
      const float* dy = dynamic0 + (size_t)inst * 4 * S;
      const float load = dy[0], soc = dy[2 * S], tim = dy[3 * S];
      bool nonuni = false;
      float d4 = 0.f;
      uint32_t mk[4];
#pragma unroll
      for (int jb = 0; jb < 4; ++jb) {
        const int j = jb * 32 + lane;
        float dmv = 0.f;
        if (j < S) {
          dmv = dy[S + j];
          nonuni |= (dy[j] != load) || (dy[2 * S + j] != soc) || (dy[3 * S + j] != tim);
          int k = 0; float best, gk, tmp;                     // rule-7 constants (problems/EVRP.py:196-202)
          geo.at(1, j, best, gk);
          for (int i = 1; i < F; ++i) { float v, gv; geo.at(1 + i, j, v, gv); if (v < best) { best = v; gk = gv; k = i; } }
          float d0v, g0v;
          geo.at(0, j, d0v, g0v);
          if (valid) {
            station_ws[((size_t)inst * 4 + 0) * SP + j] = (j == 0) ? 0.f : best;
            station_ws[((size_t)inst * 4 + 1) * SP + j] = gk;
            station_ws[((size_t)inst * 4 + 2) * SP + j] = d0v;       // D[0,:], G[0,:]: instance constants of the look-ahead
            station_ws[((size_t)inst * 4 + 3) * SP + j] = g0v;
          }
          if (j == 0) geo.at(1 + k, 0, d4, tmp);
        }
        if (valid) __stcg(dem_ws + (size_t)row * SP + j, dmv);
        mk[jb] = __ballot_sync(FULL, j < S);                     // mask = ones (:256)
      }
      d4 = __shfl_sync(FULL, d4, 0);
      if (valid && __any_sync(FULL, nonuni) && lane == 0) atomicOr(status, EVRP_ST_NONUNIFORM_DYNAMIC);
      // zeros before the first pick (:224): exact copy and both operand planes
      if (valid) __stcg(reinterpret_cast<float4*>(m_ws + (size_t)row * D) + lane, make_float4(0.f, 0.f, 0.f, 0.f));
      {
        unsigned char* p = B1 + (lane >> 4) * (2 * KBLK) + lr * 128 + ((lane & 15) << 3);
        *reinterpret_cast<uint2*>(p) = make_uint2(0u, 0u);
        *reinterpret_cast<uint2*>(p + KBLK) = make_uint2(0u, 0u);
      }
      if (lane == 0) {
        RowState& rs = sm.rs[lr];
        rs.load = load; rs.soc = soc; rs.tim = tim; rs.d4 = d4;
        rs.now = 0; rs.done = valid ? 0 : 1; rs.inst = inst; rs.valid = valid ? 1 : 0;
        rs.mk[0] = mk[0]; rs.mk[1] = mk[1]; rs.mk[2] = mk[2]; rs.mk[3] = mk[3];
        const float v1 = __fdiv_rn(soc, phys.start_soc), v2 = __fdiv_rn(tim, phys.t_limit);
        sm.veh[lr][0] = valid ? load : 0.f; sm.veh[lr][1] = valid ? v1 : 0.f; sm.veh[lr][2] = valid ? v2 : 0.f; sm.veh[lr][3] = 0.f;
        if (valid) {
          if (saved_veh) { float* sv = saved_veh + ((size_t)row * Tcap) * 3; sv[0] = load; sv[1] = v1; sv[2] = v2; }
          if (saved_mask) { uint32_t* smk = saved_mask + ((size_t)row * Tcap) * 4; smk[0] = mk[0]; smk[1] = mk[1]; smk[2] = mk[2]; smk[3] = mk[3]; }
        }
      }
    }
    fence_async_smem();                                          // plane writes -> visible to tcgen05.mma
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = sm.tmem_base;
  uint64_t pol;                                   // record loads: consumed once per step (knob bit 16 off: evict_normal)
  if (pf_lines & 0x10000) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  int t = 0;
  uint32_t it = 0;          // weight-ring counter (producer and MMA issuer keep their own copy in step)
  uint32_t n_be = 0;        // b2_empty phases consumed by the converters (single staging buffer only)
  (void)n_be;
#ifdef EVRP_PHASE_TIMING
  long long pt_acc[4] = {0, 0, 0, 0};
  long long pt_t = clock64();
  long long rt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long rt_t = 0;
  long long ft_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long ft_t = 0;
#endif

  // ---------------- step loop ----------------
  while (true) {
    int my_done = 1;
    if (wid < ROW_WARPS) {
#pragma unroll
      for (int r = 0; r < RPC / ROW_WARPS; ++r) my_done &= sm.rs[r * ROW_WARPS + wid].done;
    }
    if (__syncthreads_and(my_done)) break;
    PT_MARK(3);

    // =========================== FF_tour on the tensor cores ===========================
    if (wid == ROW_WARPS) {
      // ---- TMA producer: 8 stages of W1max (tile j, K block kb of 64), then 8 stages of W2 (K block of 64) ----
      if (lane == 0) {
        if (t == 0) sm.next_row = 0;
#pragma unroll 1
        for (int i = 0; i < NSTAGE_STEP; ++i, ++it) {
          const int s = it % WSTAGES;
          mbar_wait(&sm.empty[s], ((it / WSTAGES) & 1) ^ 1);
          unsigned char* st = ring + s * WSTAGE;
          mbar_expect_tx(&sm.full[s], WSTAGE);
          if (w.ff_stream) {
            // pre-swizzled stream: the stage is ONE contiguous 32 KB bulk copy (as 2 x 128 strided 128-byte rows the
            // tensor copy took ~0.65 us per stage and set the length of the whole FF phase)
            bulk_load(st, reinterpret_cast<const unsigned char*>(w.ff_stream) + (size_t)i * WSTAGE, WSTAGE, &sm.full[s]);
          } else if (i < 8) {
            tma_load_2d(st, &maps.w1h, &sm.full[s], (i & 1) * 64, (i >> 1) * 128);
            tma_load_2d(st + WPLANE, &maps.w1l, &sm.full[s], (i & 1) * 64, (i >> 1) * 128);
          } else {
            tma_load_2d(st, &maps.w2h, &sm.full[s], (i - 8) * 64, 0);
            tma_load_2d(st + WPLANE, &maps.w2l, &sm.full[s], (i - 8) * 64, 0);
          }
        }
      }
    } else if (wid == ROW_WARPS + 1) {
      // ---- MMA issuer ----
      if (lane == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {                        // GEMM1: hidden tile j (features 128j..128j+127)
          const uint32_t d_tmem = tmem_base + j * RPC;
#pragma unroll 1
          for (int kb = 0; kb < 2; ++kb, ++it) {
            const int s = it % WSTAGES;
            mbar_wait(&sm.full[s], (it / WSTAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(ring + s * WSTAGE), a_lo = a_hi + WPLANE;
            const uint32_t b_hi = smem_u32(B1 + kb * (2 * KBLK)), b_lo = b_hi + KBLK;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t dah = umma_desc(a_hi + k * 32), dal = umma_desc(a_lo + k * 32);
              const uint64_t dbh = umma_desc(b_hi + k * 32), dbl = umma_desc(b_lo + k * 32);
              umma_f16(d_tmem, dah, dbh, (kb | k) ? 1u : 0u, IDESC);
              umma_f16(d_tmem, dal, dbh, 1u, IDESC);
              umma_f16(d_tmem, dah, dbl, 1u, IDESC);
            }
            umma_commit(&sm.empty[s]);
          }
          umma_commit(&sm.d1_full[j]);
        }
        const uint32_t d2 = tmem_base + TM_D2;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {                         // GEMM2: K blocks 2j, 2j+1 (of 64) = hidden tile j
          mbar_wait(&sm.b2_full[j % NB2], NB2 == 2 ? ((j >> 1) & 1) : (j & 1));   // completions per step: 4 / NB2
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const unsigned char* hb = B2 + (j % NB2) * ACT_BYTES;
#pragma unroll 1
          for (int kb = 0; kb < 2; ++kb, ++it) {
            const int s = it % WSTAGES;
            mbar_wait(&sm.full[s], (it / WSTAGES) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(ring + s * WSTAGE), a_lo = a_hi + WPLANE;
            const uint32_t b_hi = smem_u32(hb + kb * (2 * KBLK)), b_lo = b_hi + KBLK;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t dah = umma_desc(a_hi + k * 32), dal = umma_desc(a_lo + k * 32);
              const uint64_t dbh = umma_desc(b_hi + k * 32), dbl = umma_desc(b_lo + k * 32);
              umma_f16(d2, dah, dbh, (j | kb | k) ? 1u : 0u, IDESC);
              umma_f16(d2, dal, dbh, 1u, IDESC);
              umma_f16(d2, dah, dbl, 1u, IDESC);
            }
            umma_commit(&sm.empty[s]);
          }
          if (j < 4 - NB2) umma_commit(&sm.b2_empty[j % NB2]);  // the converters of tile j + NB2 reuse this buffer
        }
        umma_commit(&sm.d2_full);
      }
    } else if (wid < ROW_WARPS) {
      // ---- epilogue / converters: thread = (TMEM lane = feature 32q + lane, 16 rows 16cg..16cg+15) ----
      const int q = wid & 3, cg = wid >> 2;
      const uint32_t lane_base = (uint32_t)(q * 32) << 16;
#ifdef EVRP_PHASE_TIMING
      if (tid == 0) ft_t = clock64();
#endif
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const int f = j * 128 + q * 32 + lane;
        const float bb = __ldg(w.b1 + f), wv0 = __ldg(w.w1_veh + f), wv1 = __ldg(w.w1_veh + FF + f), wv2 = __ldg(w.w1_veh + 2 * FF + f);
        mbar_wait(&sm.d1_full[j], t & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        FT_MARK(2 * j);
        uint32_t v[16];
        tmem_ld16(tmem_base + lane_base + j * RPC + cg * 16, v);
        if (j >= NB2) {                                         // GEMM2 over tile j - NB2 has drained this buffer
          if (NB2 == 2) mbar_wait(&sm.b2_empty[j & 1], t & 1);
          else { mbar_wait(&sm.b2_empty[0], n_be & 1); ++n_be; }
        }
        // feature 32q + lane of the tile = element (q & 1) * 32 + lane of K block q >> 1: 2 bytes at k * 2
        unsigned char* pb = B2 + (j % NB2) * ACT_BYTES + (q >> 1) * (2 * KBLK) + ((lane & 7) << 1);
        const int chunk = (q & 1) * 4 + (lane >> 3);
        // rows beyond this CTA's row count are never read back: their operand rows keep whatever they hold (each output
        // column of the MMA depends on its own operand row only).  At 7 rows per CTA this is 9 of 10 conversions.
        const int n_i = min(16, rows_per_cta - cg * 16);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (i >= n_i) break;
          const int r = cg * 16 + i;
          const float4 vh = *reinterpret_cast<const float4*>(&sm.veh[r][0]);
          float x = fmaf(__uint_as_float(v[i]), w1_inv_scale, fmaf(wv2, vh.z, fmaf(wv1, vh.y, fmaf(wv0, vh.x, bb))));
          x = fmaxf(x, 0.f);
          const __half hi = __float2half_rn(x);
          const __half lo = __float2half_rn(x - __half2float(hi));
          unsigned char* p = pb + r * 128 + ((chunk ^ (r & 7)) << 4);
          *reinterpret_cast<__half*>(p) = hi;
          *reinterpret_cast<__half*>(p + KBLK) = lo;
        }
        fence_async_smem();
        __syncwarp();                                           // one arrival per warp (512 per-thread arrivals serialise)
        if (lane == 0) mbar_arrive(&sm.b2_full[j % NB2]);
        FT_MARK(2 * j + 1);
      }
      PT_MARK(0);
      mbar_wait(&sm.d2_full, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      FT_MARK(8);
      {
        uint32_t v[16];
        tmem_ld16(tmem_base + lane_base + TM_D2 + cg * 16, v);
        const int n_i = min(16, rows_per_cta - cg * 16);
#pragma unroll
        for (int i = 0; i < 16; ++i) if (i < n_i) ctx[cg * 16 + i][q * 32 + lane] = __uint_as_float(v[i]) * w2_inv_scale;
      }
      if (tid == 0) sm.next_row = 0;
      FT_MARK(9);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    FT_MARK(10);
    PT_MARK(1);

    // =========================== per-row phase: each row warp walks rows ===========================
    if (wid < walkers) {                                   // (the TMA / MMA warps are idle in this phase and walk rows too)
      constexpr bool TEAMED = (TEAM == 2);
      const int tw = TEAMED ? (wid & 1) : 0;               // 0: the row's owner, 1: its helper
      WarpScratch* scr = reinterpret_cast<WarpScratch*>(ring);
      WarpScratch& ws = scr[TEAMED ? (wid & ~1) : wid];    // the owner's scratch, shared by the team
      WarpScratch& wx = scr[wid | 1];                      // (TEAM == 2) the helper's scratch: exchange area of the team
      const int team_bar = 1 + (wid >> 1);                 // named barrier of the team (0 is __syncthreads)
      int* team_slot = reinterpret_cast<int*>(&wx.list[0]);    // [0], [1]: claimed row (ping-pong), [2]: feasible count
      int iter = 0;
      (void)iter; (void)team_slot; (void)team_bar;
      const int hd = lane >> 2, sub = lane & 3;
#pragma unroll 1
      while (true) {
        int lr = 0;
        if (tw == 0) {
          if (lane == 0) lr = atomicAdd(&sm.next_row, 1);
          lr = __shfl_sync(FULL, lr, 0);
        }
        if (TEAMED) {
          if (tw == 0 && lane == 0) team_slot[iter & 1] = lr;
          team_sync(team_bar);
          lr = *reinterpret_cast<volatile int*>(&team_slot[iter & 1]);
          ++iter;
        }
        if (lr >= rows_per_cta) break;
        RowState& rs = sm.rs[lr];
        if (rs.done) continue;                             // uniform over the warp and the team
#ifdef EVRP_PHASE_TIMING
        if (tid == 0) rt_t = clock64();
#endif
        const int row = blockIdx.x * rows_per_cta + lr;
        const int inst = rs.inst;
        const int now = rs.now;
        const Geo<COORDS> geo = make_geo<COORDS>(Dm, Gm, xyp, elevp, inst, S);
        const float* embb = emb + (size_t)inst * S * D;
        const float* kvlb = kvl + (size_t)inst * S * 384;
        uint32_t mk[4] = {rs.mk[0], rs.mk[1], rs.mk[2], rs.mk[3]};
        // demand row and exact running max of this row (L2; written by this CTA only): issued early
        float dmj[4] = {0.f, 0.f, 0.f, 0.f};
        float4 mm = make_float4(0.f, 0.f, 0.f, 0.f);
        int nf = 0;
        float q[4] = {0.f, 0.f, 0.f, 0.f};
        if (tw == 0) {
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) dmj[jb] = __ldcg(dem_ws + (size_t)row * SP + jb * 32 + lane);
        mm = __ldcg(reinterpret_cast<const float4*>(m_ws + (size_t)row * D) + lane);
        // ---- feasible list (ascending node index) ----
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) {
          const uint32_t bits = mk[jb];
          if ((bits >> lane) & 1u) ws.list[nf + __popc(bits & ((1u << lane) - 1u))] = (unsigned char)(jb * 32 + lane);
          nf += __popc(bits);
        }
        __syncwarp();
        // pad the list to a multiple of 16 with its last node: the 8- / 16-wide passes below then need no bounds checks
        // (a duplicate changes neither the running maximum nor, with a zero weight, the weighted sum)
        if (lane < 16 && nf + lane < ((nf + 15) & ~15)) ws.list[nf + lane] = ws.list[nf - 1];
        __syncwarp();
        // Just-in-time L2 prefetch (fire-and-forget, no registers): every pass requests ALL of its records one pass
        // ahead of the demand loads - K here, V when the softmax starts, logit-K when the V pass starts - so the 8-deep
        // demand batches find their lines in L2.  Requesting everything up front, or the next row's records, measured
        // slower (the requests queue ahead of the critical demand loads): profiles/r1_prefetch_sweep.md.
        if (pf_lines & 1) {
          for (int i = lane; i < nf; i += 32) {
            const float* p = kvlb + (size_t)ws.list[i] * 384;
#pragma unroll
            for (int c = 0; c < 4; ++c) prefetch_l2(p + c * 32);
          }
        }
        // ---- query (:394): (fixed_ctx + ctx) + emb[now] ; lane holds 4 channels of head lane/4 ----
        {
          const float4 fc = ldg4(fixed_ctx + (size_t)inst * D + lane * 4);
          const float4 en = ldg4(embb + (size_t)now * D + lane * 4);
          const float4 cx = *reinterpret_cast<const float4*>(&ctx[lr][lane * 4]);
          const float4 b2 = ldg4(w.b2 + lane * 4);
          q[0] = (fc.x + (cx.x + b2.x)) + en.x;
          q[1] = (fc.y + (cx.y + b2.y)) + en.y;
          q[2] = (fc.z + (cx.z + b2.z)) + en.z;
          q[3] = (fc.w + (cx.w + b2.w)) + en.w;
        }
        if (TEAMED) {                                      // hand the query and the list length to the helper
          *reinterpret_cast<float4*>(&ws.heads[lane * 4]) = make_float4(q[0], q[1], q[2], q[3]);
          if (lane == 0) team_slot[2] = nf;
        }
        }                                                  // (tw == 0)
        if (TEAMED) {
          team_sync(team_bar);
          if (tw == 1) {
            nf = *reinterpret_cast<volatile int*>(&team_slot[2]);
            const float4 qq = *reinterpret_cast<const float4*>(&ws.heads[lane * 4]);
            q[0] = qq.x; q[1] = qq.y; q[2] = qq.z; q[3] = qq.w;
          }
        }
        RT_MARK(0);
        // ---- compat[h][i] = q_h . K[j_i, h] / 4 ; CH rows (CH x 512 B) in flight per warp ----
        constexpr int CH = 8;
        float cmx = -INFINITY;                             // per-head maximum, tracked by all four lanes of the head
        for (int i0 = tw * CH; i0 < nf; i0 += CH * TEAM) {
          float4 kv[CH];
#pragma unroll
          for (int a = 0; a < CH; ++a) kv[a] = ldg4_pol(kvlb + (size_t)ws.list[i0 + a] * 384 + lane * 4, pol);
#pragma unroll
          for (int a = 0; a < CH; ++a) {
            float p = fmaf(q[3], kv[a].w, fmaf(q[2], kv[a].z, fmaf(q[1], kv[a].y, q[0] * kv[a].x)));
            p += __shfl_xor_sync(FULL, p, 1);
            p += __shfl_xor_sync(FULL, p, 2);
            p *= 0.25f;
            cmx = fmaxf(cmx, p);
            if (sub == 0) ws.pc[hd][i0 + a] = p;
          }
        }
        if (TEAMED) {                                      // the maximum is over both halves of the list
          wx.logit[tw * 32 + lane] = cmx;
          team_sync(team_bar);
          cmx = fmaxf(cmx, wx.logit[(tw ^ 1) * 32 + lane]);
        } else {
          __syncwarp();
        }
        RT_MARK(1);
        // ---- per-head softmax over the list; lane (hd, sub) strides i = sub, sub+4, ... ----
        float sum = 0.f, sum_b = 0.f;                      // partial sums of the even / odd batches (see TEAM above)
        {
          if ((pf_lines & 4) && tw == 0) {
            for (int i = lane; i < nf; i += 32) {
              const float* p = kvlb + (size_t)ws.list[i] * 384 + 128;
#pragma unroll
              for (int c = 0; c < 4; ++c) prefetch_l2(p + c * 32);
            }
          }
          for (int i0 = tw * CH; i0 < nf; i0 += CH * TEAM) {   // this warp's batches; padded entries weigh nothing
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int i = i0 + sub + 4 * k;
              const float e = (i < nf) ? expf(ws.pc[hd][i] - cmx) : 0.f;
              ws.pc[hd][i] = e;
              sum += e;
            }
            if (!TEAMED) { const float tmp = sum; sum = sum_b; sum_b = tmp; }      // one warp: alternate the two partial sums
          }
          sum += __shfl_xor_sync(FULL, sum, 1);
          sum += __shfl_xor_sync(FULL, sum, 2);               // the division by the sum is applied to the 4 head outputs below
          if (!TEAMED) {
            sum_b += __shfl_xor_sync(FULL, sum_b, 1);
            sum_b += __shfl_xor_sync(FULL, sum_b, 2);
          }
        }
        __syncwarp();
        RT_MARK(2);
        // ---- heads = attn @ V ; lane owns 4 channels ----
        {
          if ((pf_lines & 8) && tw == 0) {
            for (int i = lane; i < nf; i += 32) {
              const float* p = kvlb + (size_t)ws.list[i] * 384 + 256;
#pragma unroll
              for (int c = 0; c < 4; ++c) prefetch_l2(p + c * 32);
            }
          }
          float hv[4] = {0.f, 0.f, 0.f, 0.f}, hb[4] = {0.f, 0.f, 0.f, 0.f};     // even / odd batches
          for (int i0 = tw * CH; i0 < nf; i0 += CH * TEAM) {
            float4 vv[CH]; float pr[CH];
#pragma unroll
            for (int a = 0; a < CH; ++a) {
              vv[a] = ldg4_pol(kvlb + (size_t)ws.list[i0 + a] * 384 + 128 + lane * 4, pol);
              pr[a] = ws.pc[hd][i0 + a];
            }
#pragma unroll
            for (int a = 0; a < CH; ++a) {
              hv[0] = fmaf(pr[a], vv[a].x, hv[0]); hv[1] = fmaf(pr[a], vv[a].y, hv[1]);
              hv[2] = fmaf(pr[a], vv[a].z, hv[2]); hv[3] = fmaf(pr[a], vv[a].w, hv[3]);
            }
            if (!TEAMED) {                                   // one warp: alternate the two accumulators
#pragma unroll
              for (int c = 0; c < 4; ++c) { const float tmp = hv[c]; hv[c] = hb[c]; hb[c] = tmp; }
            }
          }
          if (TEAMED) {                                      // the helper's partial sums reach the owner through its scratch
            if (tw == 1) {
              *reinterpret_cast<float4*>(&wx.heads[lane * 4]) = make_float4(hv[0], hv[1], hv[2], hv[3]);
              wx.logit[64 + lane] = sum;
            }
            team_sync(team_bar);
            if (tw == 0) {
              const float4 o = *reinterpret_cast<const float4*>(&wx.heads[lane * 4]);
              hb[0] = o.x; hb[1] = o.y; hb[2] = o.z; hb[3] = o.w;
              sum_b = wx.logit[64 + lane];
            }
          }
          if (tw == 0) {
            // x + y is commutative, so which accumulator held the even batches does not matter
            const float inv = div_generic(1.f, sum + sum_b);   // one IEEE division per head, then 4 multiplies
            *reinterpret_cast<float4*>(&ws.heads[lane * 4]) =
                make_float4((hv[0] + hb[0]) * inv, (hv[1] + hb[1]) * inv, (hv[2] + hb[2]) * inv, (hv[3] + hb[3]) * inv);
          }
        }
        if (TEAMED) team_sync(team_bar); else __syncwarp();
        RT_MARK(3);
        // ---- logits[i] = heads . lK[j_i] / sqrt(128) -> tanh clip -> / temp  (project_out folded into lK).
        //      8 lanes per node: lane (grp, s8) covers channels 4*s8 + 32*c, c = 0..3; 4 nodes per pass ----
        {
          const int grp = lane >> 3, s8 = lane & 7;
          float4 hq[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) hq[c] = *reinterpret_cast<const float4*>(&ws.heads[s8 * 4 + c * 32]);
          for (int i0 = tw * 16; i0 < nf; i0 += 16 * TEAM) {   // 4 nodes per sub-pass, 4 sub-passes in flight
            float4 lv[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float* p = kvlb + (size_t)ws.list[i0 + u * 4 + grp] * 384 + 256 + s8 * 4;
#pragma unroll
              for (int c = 0; c < 4; ++c) lv[u][c] = ldg4_pol(p + c * 32, pol);
            }
            float au[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float acc = 0.f;
#pragma unroll
              for (int c = 0; c < 4; ++c)
                acc = fmaf(hq[c].w, lv[u][c].w, fmaf(hq[c].z, lv[u][c].z, fmaf(hq[c].y, lv[u][c].y, fmaf(hq[c].x, lv[u][c].x, acc))));
              acc += __shfl_xor_sync(FULL, acc, 1);
              acc += __shfl_xor_sync(FULL, acc, 2);
              acc += __shfl_xor_sync(FULL, acc, 4);
              au[u] = acc;                                   // every lane of the group holds the node's full dot product
            }
            // one copy of the divide / tanh / divide tail: lane (grp, s8 < 4) finishes node i0 + 4*s8 + grp
            const int i = i0 + s8 * 4 + grp;
            if (s8 < 4) {                                    // (padded positions compute a duplicate nobody reads)
              float z = div_generic(sel4(au, s8), sqrt_d);
              if (cfg.tanh_clipping > 0.f) z = tanhf(z) * cfg.tanh_clipping;
              ws.logit[i] = div_generic(z, cfg.temp);
            }
          }
        }
        if (TEAMED) { team_sync(team_bar); if (tw == 1) continue; }     // the helper is done with this row
        else __syncwarp();
        RT_MARK(4);
        // ---- log_softmax over the list (:403) ----
        float zl[4];
        float mx = -INFINITY;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int i = rr * 32 + lane;
          zl[rr] = (i < nf) ? ws.logit[i] : -INFINITY;
          mx = fmaxf(mx, zl[rr]);
        }
        mx = warp_max(mx);
        float se = 0.f;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) if (rr * 32 + lane < nf) se += expf(zl[rr] - mx);
        se = warp_sum(se);
        const float lgs = logf(se);
        const float lse = lgs + mx;
        float lp[4], pr[4];
        bool bad = false;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          lp[rr] = (zl[rr] - mx) - lgs;                          // x - max - log(sum(exp(x - max)))
          pr[rr] = (rr * 32 + lane < nf) ? expf(lp[rr]) : -1.f;  // probs = log_p.exp() (:265)
          bad |= (rr * 32 + lane < nf) && !(lp[rr] == lp[rr]);
        }
        if (__any_sync(FULL, bad) && lane == 0) atomicOr(status, EVRP_ST_NAN_LOGIT);
        // ---- pick ----
        int sel_i = 0;
        if (MODE == 2) {
          const int j = (int)forced[(size_t)row * Tcap + t];
          const bool ok = (j >= 0 && j < S) && ((sel4(mk, j >> 5) >> (j & 31)) & 1u);
          if (!ok) { if (lane == 0) atomicOr(status, EVRP_ST_INFEASIBLE_PICK); sel_i = -1 - max(0, min(j, S - 1)); }
          else {
            int pos = 0;
#pragma unroll
            for (int jb = 0; jb < 4; ++jb) {
              if (jb < (j >> 5)) pos += __popc(mk[jb]);
              else if (jb == (j >> 5)) pos += __popc(mk[jb] & ((1u << (j & 31)) - 1u));
            }
            sel_i = pos;
          }
        } else if (MODE == 0) {
          // first-index argmax of probs (torch.max(1), :327)
          float bv = pr[0]; int bi = lane;
#pragma unroll
          for (int rr = 1; rr < 4; ++rr) if (pr[rr] > bv) { bv = pr[rr]; bi = rr * 32 + lane; }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, bv, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
          }
          sel_i = bi;
        } else {
          // inverse-CDF sample of probs over the list (multinomial(1), :333); lane owns positions 4*lane..4*lane+3
          const float u = uniforms ? uniforms[(size_t)row * Tcap + t] : uniform01(cfg.seed, (uint64_t)row, (uint32_t)t);
          float c[4]; float s = 0.f;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int i = lane * 4 + a;
            const float pv = (i < nf) ? expf(ws.logit[i] - lse) : 0.f;
            s += pv; c[a] = s;
          }
          float incl = s;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
          const float total = __shfl_sync(FULL, incl, 31);
          const float excl = incl - s;
          const float target = u * total;
          int cand = 1 << 30;
#pragma unroll
          for (int a = 3; a >= 0; --a) if (lane * 4 + a < nf && excl + c[a] > target) cand = lane * 4 + a;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(FULL, cand, o));
          sel_i = (cand < nf) ? cand : nf - 1;
        }
        int c_node; float lp_sel;
        if (sel_i >= 0) {
          c_node = ws.list[sel_i];
          lp_sel = __shfl_sync(FULL, sel4(lp, sel_i >> 5), sel_i & 31);
        } else { c_node = -1 - sel_i; lp_sel = -INFINITY; }

        RT_MARK(5);
        // ---- loads for the transition and the new mask, issued together ----
        float dist, slp;
        geo.at(now, c_node, dist, slp);
        float xc = 0.f, yc = 0.f, ec_elev = 0.f;                  // coordinate mode: the chosen node, row D[c,:] / G[c,:] on the fly
        if (geo.coords()) { xc = __ldg(geo.x + c_node); yc = __ldg(geo.y + c_node); ec_elev = __ldg(geo.e + c_node); }
        const float4 ec = ldg4(embb + (size_t)c_node * D + lane * 4);
        float dcj[4], gcj[4], d0j[4], g0j[4], d3j[4], s3j[4];
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) {
          const int j = jb * 32 + lane;
          const bool in = (jb < nblk) && (j < S);
          dcj[jb] = gcj[jb] = 0.f;
          if (in) {
            if (geo.coords()) geo.pair(xc, yc, ec_elev, j, dcj[jb], gcj[jb]);
            else { dcj[jb] = __ldg(geo.D + (size_t)c_node * S + j); gcj[jb] = __ldg(geo.G + (size_t)c_node * S + j); }
          }
          d3j[jb] = in ? station_ws[((size_t)inst * 4 + 0) * SP + j] : 0.f;
          s3j[jb] = in ? station_ws[((size_t)inst * 4 + 1) * SP + j] : 0.f;
          d0j[jb] = in ? station_ws[((size_t)inst * 4 + 2) * SP + j] : 0.f;
          g0j[jb] = in ? station_ws[((size_t)inst * 4 + 3) * SP + j] : 0.f;
        }
        // ---- state transition (problems/EVRP.py:226-280) on warp-uniform scalars ----
        float load = rs.load, soc = rs.soc, tim = rs.tim;
        const float d4 = rs.d4;
        const float tc = travel_time(phys, dist);
        const float e = leg_energy(phys, vehicle_mass(phys, load), slp, tc);
        const bool is_depot = (c_node == 0), is_station = (c_node > 0 && c_node <= F), is_cust = (c_node > F);
        tim = __fsub_rn(tim, tc);
        if (is_depot) tim = phys.t_limit;
        if (is_station) tim = __fsub_rn(tim, phys.charging_time);
        if (is_cust) tim = __fsub_rn(tim, phys.serve_time);
        soc = __fsub_rn(soc, e);
        if (c_node <= F) soc = phys.start_soc;
        if (is_cust) {
          const float dmc = __shfl_sync(FULL, sel4(dmj, c_node >> 5), c_node & 31);
          const float nl = fmaxf(__fsub_rn(load, dmc), 0.f);
          const float nd = fmaxf(__fsub_rn(dmc, load), 0.f);
          if (lane == (c_node & 31)) {
            const int jb = c_node >> 5;
            if (jb == 0) dmj[0] = nd; else if (jb == 1) dmj[1] = nd; else if (jb == 2) dmj[2] = nd; else dmj[3] = nd;
            __stcg(dem_ws + (size_t)row * SP + c_node, nd);
          }
          if (lane == 0) { dmj[0] = __fadd_rn(-1.f, nl); __stcg(dem_ws + (size_t)row * SP, dmj[0]); }
          load = nl;
        }
        if (is_depot) { load = 1.f; if (lane == 0) { dmj[0] = 0.f; __stcg(dem_ws + (size_t)row * SP, 0.f); } }
        if (lane == 0) {
          const size_t o = (size_t)row * Tcap + t;
          pi[o] = c_node; ll[o] = lp_sel;
          reward[o] = phys.objective ? dist : e;
          if (energy) energy[o] = e;
        }
        // ---- running max of visited embeddings (route_feature :207-214): exact copy + fp16 (hi, lo) operand planes ----
        {
          if (t == 0) mm = ec;
          else { mm.x = fmaxf(mm.x, ec.x); mm.y = fmaxf(mm.y, ec.y); mm.z = fmaxf(mm.z, ec.z); mm.w = fmaxf(mm.w, ec.w); }
          __stcg(reinterpret_cast<float4*>(m_ws + (size_t)row * D) + lane, mm);
          const __half2 h01 = __floats2half2_rn(mm.x, mm.y), h23 = __floats2half2_rn(mm.z, mm.w);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn(mm.x - f01.x, mm.y - f01.y), l23 = __floats2half2_rn(mm.z - f23.x, mm.w - f23.y);
          unsigned char* p = B1 + (lane >> 4) * (2 * KBLK) + lr * 128 + ((((lane & 15) >> 1) ^ (lr & 7)) << 4) + ((lane & 1) << 3);
          *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
          *reinterpret_cast<uint2*>(p + KBLK) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
        }
        RT_MARK(6);
        // ---- new mask (problems/EVRP.py:105-223) ----
        bool any_dem = false, cust_dem = false;
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) {
          const int j = jb * 32 + lane;
          any_dem |= (j < S) && (dmj[jb] != 0.f);
          cust_dem |= (j < S) && (j >= F) && (dmj[jb] != 0.f);
        }
        any_dem = __any_sync(FULL, any_dem);
        cust_dem = __any_sync(FULL, cust_dem);
        bool done = false;
        if (!any_dem) {
          done = true;                                     // this row's tour is complete (:127)
        } else {
          const bool combined = (load == 0.f) || !cust_dem;
          bool cust_ok = false;
          uint32_t nm[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
          for (int jb = 0; jb < nblk; ++jb) {              // rolled: one instance of the look-ahead code (i-cache)
            const int j = jb * 32 + lane;
            bool mbit = false;
            const float dm = sel4(dmj, jb);
            if (j < S) {
              mbit = (dm != 0.f) && (dm < load);
              if (j == c_node) mbit = false;
              if (c_node <= F && j >= 1 && j <= F) mbit = false;
              if (is_station && j == 0) mbit = true;
              if (is_cust && j <= F) mbit = true;
              if (is_depot && j == 0) mbit = false;
              if (combined) mbit = (j == 0);
            }
            if (__any_sync(FULL, mbit)) {                  // look-ahead rules 5-7 only where something is still open
              const float a_dc = sel4(dcj, jb), a_gc = sel4(gcj, jb), a_d0 = sel4(d0j, jb), a_g0 = sel4(g0j, jb),
                          a_d3 = sel4(d3j, jb), a_s3 = sel4(s3j, jb);
              bool bad = !fastdiv;
              bool ok = reachable_fast(phys, j, F, load, load, dm, soc, tim, a_dc, a_gc, a_d0, a_g0, a_d3, a_s3, d4, bad);
              if (__any_sync(FULL, bad && mbit))           // out-of-range operand or another divisor: generic IEEE division
                ok = reachable_generic(phys, j, F, load, load, dm, soc, tim, a_dc, a_gc, a_d0, a_g0, a_d3, a_s3, d4);
              if (!ok) mbit = false;
            }
            if (j > F && mbit) cust_ok = true;
            const uint32_t bal = __ballot_sync(FULL, mbit);
            if (jb == 0) nm[0] = bal; else if (jb == 1) nm[1] = bal; else if (jb == 2) nm[2] = bal; else nm[3] = bal;
          }
          cust_ok = __any_sync(FULL, cust_ok);
          if (!cust_ok) nm[0] |= 1u;                       // rule 8 (:220-221)
          mk[0] = nm[0]; mk[1] = nm[1]; mk[2] = nm[2]; mk[3] = nm[3];
          if ((mk[0] | mk[1] | mk[2] | mk[3]) == 0u) {
            if (lane == 0) atomicOr(status, EVRP_ST_NO_FEASIBLE);
            done = true;
          } else if (is_depot && mk[0] == 1u && (mk[1] | mk[2] | mk[3]) == 0u) {
            // fixed point: back at the depot (state reset to load 1 / Start_SOC / t_limit, :254-276) with open demand
            // that no move can reach, so rule 8 leaves the depot as the only choice -- every further step repeats
            // (pi 0, log-prob 0, reward 0) until the reference's step cap (nets/DRLModel.py:255).  The outputs are
            // already zero: retire the row now and let the host widen T to max_steps.
            if (lane == 0) atomicOr(status, EVRP_ST_STUCK_AT_DEPOT);
            done = true;
          } else if (t + 1 >= Tcap) {
            if (lane == 0) atomicOr(status, EVRP_ST_STEP_CAP);
            done = true;
          }
        }
        if (lane == 0) {
          rs.load = load; rs.soc = soc; rs.tim = tim; rs.now = c_node;
          rs.mk[0] = mk[0]; rs.mk[1] = mk[1]; rs.mk[2] = mk[2]; rs.mk[3] = mk[3];
          if (done) { rs.done = 1; steps_used[row] = t + 1; atomicMax(t_max, t + 1); }
          else {
            // vehicle scalars of the next step (:215-222) + trace
            const float v1 = div_generic(soc, phys.start_soc), v2 = div_generic(tim, phys.t_limit);
            sm.veh[lr][0] = load; sm.veh[lr][1] = v1; sm.veh[lr][2] = v2;
            const size_t o = (size_t)row * Tcap + t + 1;
            if (saved_veh) { float* sv = saved_veh + o * 3; sv[0] = load; sv[1] = v1; sv[2] = v2; }
            if (saved_mask) { uint32_t* smk = saved_mask + o * 4; smk[0] = mk[0]; smk[1] = mk[1]; smk[2] = mk[2]; smk[3] = mk[3]; }
          }
        }
        __syncwarp();
        RT_MARK(7);
      }
      PT_MARK(2);
      fence_async_smem();          // B1 plane writes (and the scratch the next TMA overwrites) -> async proxy
    }
    ++t;
  }
#ifdef EVRP_PHASE_TIMING
  if (tid == 0) for (int i = 0; i < 4; ++i) atomicAdd(&g_phase_cycles_tc[i], (unsigned long long)pt_acc[i]);
  if (tid == 0) for (int i = 0; i < 8; ++i) atomicAdd(&g_row_cycles_tc[i], (unsigned long long)rt_acc[i]);
  if (tid == 0) for (int i = 0; i < 12; ++i) atomicAdd(&g_ff_cycles_tc[i], (unsigned long long)ft_acc[i]);
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wid == ROW_WARPS + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(EVRP_TMEM_COLS) : "memory");
}

}  // namespace dtc

#ifdef EVRP_PHASE_TIMING
int debug_phase_cycles_tc(unsigned long long* out4, int reset) {
  if (cudaMemcpyFromSymbol(out4, dtc::g_phase_cycles_tc, sizeof(unsigned long long) * 4) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(dtc::g_phase_cycles_tc, z, sizeof(z)); }
  return 0;
}
int debug_ff_cycles_tc(unsigned long long* out12, int reset) {
  if (cudaMemcpyFromSymbol(out12, dtc::g_ff_cycles_tc, sizeof(unsigned long long) * 12) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[12] = {0}; cudaMemcpyToSymbol(dtc::g_ff_cycles_tc, z, sizeof(z)); }
  return 0;
}
int debug_row_cycles_tc(unsigned long long* out8, int reset) {
  if (cudaMemcpyFromSymbol(out8, dtc::g_row_cycles_tc, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(dtc::g_row_cycles_tc, z, sizeof(z)); }
  return 0;
}
#endif

// host launcher (called from evrp_rollout in evrp_decode.cu when the decoder weights carry ff_mode == 0)
int launch_rollout_tc(const evrp_phys* phys, const evrp_decoder_weights* w, const evrp_rollout_cfg* cfg, const float* emb,
                      const float* kvl, const float* fixed_ctx, const float* dynamic0, const float* distances,
                      const float* slope, const float* static_xy, const float* elevations, int B, int S, int F,
                      const int64_t* forced_actions, const float* uniforms,
                      int64_t* pi, float* ll, float* reward, float* energy, uint32_t* saved_mask, float* saved_veh,
                      int32_t* steps_used, int32_t* t_max, int32_t* status, float* station_ws, float* dem_ws, float* m_ws,
                      cudaStream_t st) {
  using namespace dtc;
  if (!w->w1max_hi || !w->w1max_lo || !w->w2_hi || !w->w2_lo) return -EVRP_E_BADARG;
  Maps maps;
  int rc;
  if ((rc = tc::make_weight_map_f16(&maps.w1h, w->w1max_hi, FF, D, 128))) return rc;
  if ((rc = tc::make_weight_map_f16(&maps.w1l, w->w1max_lo, FF, D, 128))) return rc;
  if ((rc = tc::make_weight_map_f16(&maps.w2h, w->w2_hi, D, FF, 128))) return rc;
  if ((rc = tc::make_weight_map_f16(&maps.w2l, w->w2_lo, D, FF, 128))) return rc;
  const int rows = B * cfg->n_replicas;
  const int mode = forced_actions ? 2 : (cfg->decode_type == EVRP_DECODE_SAMPLE ? 1 : 0);
  const bool coords = (distances == nullptr);
  int dev = 0, sms = 148;
  EVRP_CUDA_OK(cudaGetDevice(&dev));
  EVRP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // balance: whole waves of one CTA per SM, rows spread evenly (<= RPC rows per CTA)
  const int slots = EVRP_CTAS * sms;
  const int min_ctas = (rows + RPC - 1) / RPC;
  const int waves = (min_ctas + slots - 1) / slots;
  int grid = waves * slots;
  if (grid > rows) grid = rows;
  int rows_per_cta = (rows + grid - 1) / grid;
  // Small batches: every CTA streams the whole FF_tour weight set (512 KB) from L2 every step whatever its row count, and
  // one warp walks one row, so up to `walkers` rows per CTA cost one row latency: fewer, fuller CTAs shorten the FF phase.
  static int min_rpc = -1;
  if (min_rpc < 0) { const char* e = getenv("EVRP_MIN_RPC"); min_rpc = e ? atoi(e) : 1; }
  if (rows_per_cta < min_rpc) rows_per_cta = min_rpc < RPC ? min_rpc : RPC;
  grid = (rows + rows_per_cta - 1) / rows_per_cta;
  static int pf_lines = -1;              // tuning knob: bit 0 K, bit 2 V, bit 3 logit-K just-in-time prefetch,
                                         // bit 16 record loads with L2::evict_first
  if (pf_lines < 0) { const char* l = getenv("EVRP_PF_LINES"); pf_lines = l ? atoi(l) : (13 | 0x10000); }
  static int walkers = -1;
  if (walkers < 0) { const char* e = getenv("EVRP_WALKERS"); walkers = e ? atoi(e) : NWALK; if (walkers > NWALK) walkers = NWALK; }
  // two warps per row when every row of the CTA can have its own team (small batches: latency-bound, most warps idle)
  static int team_knob = -1;             // tuning knob: 1 forces one warp per row
  if (team_knob < 0) { const char* e = getenv("EVRP_TEAM"); team_knob = e ? atoi(e) : 2; }
  const bool team2 = team_knob >= 2 && 2 * rows_per_cta <= (walkers & ~1) && walkers >= 2;
  auto pick = [&](auto coords_c, auto team_c) {
    constexpr bool C = decltype(coords_c)::value;
    constexpr int T = decltype(team_c)::value;
    return mode == 2 ? rollout_tc_kernel<2, C, T> : mode == 1 ? rollout_tc_kernel<1, C, T> : rollout_tc_kernel<0, C, T>;
  };
  auto kern = coords ? (team2 ? pick(std::true_type{}, std::integral_constant<int, 2>{}) : pick(std::true_type{}, std::integral_constant<int, 1>{}))
                     : (team2 ? pick(std::false_type{}, std::integral_constant<int, 2>{}) : pick(std::false_type{}, std::integral_constant<int, 1>{}));
  EVRP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  const int walkers_l = team2 ? (walkers & ~1) : walkers;
  kern<<<grid, NT, SMEM_BYTES, st>>>(maps, *phys, *w, *cfg, emb, kvl, fixed_ctx, dynamic0, distances, slope, static_xy, elevations, station_ws,
                                     dem_ws, m_ws, B, S, F, rows, rows_per_cta, pf_lines, walkers_l, forced_actions, uniforms, pi, ll, reward,
                                     energy, saved_mask, saved_veh, steps_used, t_max, status);
  EVRP_LAUNCH_OK();
  return 0;
}

}  // namespace evrp
