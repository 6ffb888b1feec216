// Persistent rollout kernel: ONE launch decodes whole tours.
//
//   AttentionModel._inner            nets/DRLModel.py:240-280   (the step loop lives inside the kernel)
//   route_feature                    nets/DRLModel.py:203-237   (running max + FF_tour)
//   _get_log_p / _one_to_many_logits nets/DRLModel.py:379-452   (8-head masked glimpse, pointer logits)
//   _select_node                     nets/DRLModel.py:316-342   (greedy first-index argmax / sampling)
//   _calc_log_likelihood             nets/DRLModel.py:182-190
//   update_dynamic / update_mask     problems/EVRP.py:105-280   (fused, bit-exact recipe of evrp_common.cuh)
//
// Layout: a CTA (256 threads, 2 CTAs per SM) owns RPC=32 rollout rows for their whole life; each warp owns 4 rows.
// Per step
//   1. FF_tour as two CTA-wide register-tiled fp32 GEMMs over the 32 rows ([32x128]x[128x512], [32x512]x[512x128],
//      8x8 / 8x4 register tiles; weights stream coalesced from L2 once per CTA-step and are shared by 32 rows; the
//      running max, the hidden layer and the context stay in shared memory),
//   2. per row (warp): glimpse attention / logits over the FEASIBLE nodes only (compact index list from the 128-bit
//      mask; K|V|logitK rows are 512-B coalesced float4 loads issued 4 at a time; the rows are L2-prefetched as soon
//      as the mask is known),
//   3. warp-level masked log-softmax, first-index argmax / inverse-CDF sample,
//   4. state transition + look-ahead feasibility mask (bit-exact fp32 recipe), state in shared memory.
// Rows retire individually; the host reads back only t_max and the status word.
#include <math.h>
#include <type_traits>
#include "evrp_common.cuh"

namespace evrp {

constexpr int RPC = 32;          // rollout rows per CTA
constexpr int RPW = 4;           // rows per warp
constexpr int NT = 256;          // 8 warps
constexpr int SP = EVRP_MAX_S;   // padded node count (128)
constexpr int PSTR = SP + 4;     // head stride of the compat buffer (bank-conflict-free for the 4-lane groups)

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

struct AttScratch {
  float ctx[RPC][D];             // FF_tour output (written after the second GEMM has consumed h1)
  WarpScratch ws[NT / 32];
};

union FFUnion {
  float h1[RPC][FF];             // FF_tour hidden (FF phase)
  AttScratch att;                // per-row phase
};

struct __align__(16) DecodeSmem {
  float m[RPC][D];               // running max of visited embeddings          (route_feature :214)
  float dem[RPC][SP];            // demand rows
  float veh[RPC][4];             // load, soc/Start_SOC, time/t_limit          (:215-222)
  RowState rs[RPC];
  int next_row;                  // per-step work counter: warps pull rows dynamically (balances ragged feasible sets)
  int pad[3];
  FFUnion u;
};

#ifdef EVRP_PHASE_TIMING
// debug build only (tools/phase_timing.py): cycles spent by warp 0 of every CTA in [FF1, FF2, row phase, barrier wait]
__device__ unsigned long long g_phase_cycles[4];
#define PT_MARK(i) do { if (tid == 0) { const long long _n = clock64(); pt_acc[i] += _n - pt_t; pt_t = _n; } } while (0)
#else
#define PT_MARK(i) do { } while (0)
#endif

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// K / V / logit-K records are consumed once per decode step: L2::evict_first (see evrp_decode_tc.cu)
__device__ __forceinline__ float4 ldg4_pol(const float* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}

// register-array select without dynamic indexing (keeps the arrays out of local memory)
template <typename T>
__device__ __forceinline__ T sel4(const T (&a)[4], int i) { return i == 0 ? a[0] : i == 1 ? a[1] : i == 2 ? a[2] : a[3]; }

// MODE: 0 greedy, 1 sample, 2 teacher-forced (separate instantiations keep each kernel's code small: the row phase
// is instruction-cache sensitive)
// POLICY: 0 = the paper's model (route context through FF_tour, nets/DRLModel.py:203-237,394); 1 = nets/AM.py
// (:336-337,352-377): query = fixed_ctx + project_step_context([emb[now], load]) -- no FF phase at all; `emb` is the
// encoder's step table (emb @ project_step_context.weight[:, :128]^T), so the step context is one row read + one FMA
template <int MODE, int POLICY, bool COORDS>
__global__ void __launch_bounds__(NT, 2)
rollout_kernel(evrp_phys phys, evrp_decoder_weights w, evrp_rollout_cfg cfg,
               const float* __restrict__ emb, const float* __restrict__ kvl, const float* __restrict__ fixed_ctx,
               const float* __restrict__ dynamic0, const float* __restrict__ Dm, const float* __restrict__ Gm,
                  const float* __restrict__ xyp, const float* __restrict__ elevp,
               float* station_ws, int B, int S, int F, int rows_total, int rows_per_cta,
               const int64_t* __restrict__ forced, const float* __restrict__ uniforms,
               int64_t* __restrict__ pi, float* __restrict__ ll, float* __restrict__ reward,
               float* __restrict__ energy, uint32_t* __restrict__ saved_mask, float* __restrict__ saved_veh,
               int32_t* __restrict__ steps_used, int32_t* __restrict__ t_max, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DecodeSmem& sm = *reinterpret_cast<DecodeSmem*>(smem_raw);
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int Tcap = cfg.max_steps;
  const int nblk = (S + 31) >> 5;
  const float sqrt_d = sqrtf((float)D);

  // ---------------- prologue: per-row state ----------------
  for (int r = 0; r < RPW; ++r) {
    const int lr = r * (NT / 32) + wid;                       // warp w owns local rows w, 8+w, 16+w, 24+w
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
      sm.dem[lr][j] = dmv;
      mk[jb] = __ballot_sync(FULL, j < S);                     // mask = ones (:256)
    }
    d4 = __shfl_sync(FULL, d4, 0);
    if (valid && __any_sync(FULL, nonuni) && lane == 0) atomicOr(status, EVRP_ST_NONUNIFORM_DYNAMIC);
#pragma unroll
    for (int i = 0; i < 4; ++i) sm.m[lr][lane * 4 + i] = 0.f;  // zeros before the first pick (:224)
    if (lane == 0) {
      RowState& rs = sm.rs[lr];
      rs.load = load; rs.soc = soc; rs.tim = tim; rs.d4 = d4;
      rs.now = 0; rs.done = valid ? 0 : 1; rs.inst = inst; rs.valid = valid ? 1 : 0;
      rs.mk[0] = mk[0]; rs.mk[1] = mk[1]; rs.mk[2] = mk[2]; rs.mk[3] = mk[3];
      sm.veh[lr][0] = 0.f; sm.veh[lr][1] = 0.f; sm.veh[lr][2] = 0.f; sm.veh[lr][3] = 0.f;
    }
  }
  __syncthreads();
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  int t = 0;
#ifdef EVRP_PHASE_TIMING
  long long pt_acc[4] = {0, 0, 0, 0};
  long long pt_t = clock64();
#endif

  // ---------------- step loop ----------------
  while (true) {
    int my_done = 1;
#pragma unroll
    for (int r = 0; r < RPW; ++r) my_done &= sm.rs[r * (NT / 32) + wid].done;
    if (__syncthreads_and(my_done)) break;
    if (tid == 0) sm.next_row = 0;
    // vehicle scalars + trace
    if (lane < RPW) {
      const int lr = lane * (NT / 32) + wid;
      const RowState& rs = sm.rs[lr];
      if (!rs.done) {
        const float v1 = __fdiv_rn(rs.soc, phys.start_soc), v2 = __fdiv_rn(rs.tim, phys.t_limit);
        sm.veh[lr][0] = rs.load; sm.veh[lr][1] = v1; sm.veh[lr][2] = v2;
        const size_t row = (size_t)blockIdx.x * rows_per_cta + lr;
        if (saved_veh) { float* sv = saved_veh + (row * Tcap + t) * 3; sv[0] = rs.load; sv[1] = v1; sv[2] = v2; }
        if (saved_mask) {
          uint32_t* smk = saved_mask + (row * Tcap + t) * 4;
          smk[0] = rs.mk[0]; smk[1] = rs.mk[1]; smk[2] = rs.mk[2]; smk[3] = rs.mk[3];
        }
      }
    }
    __syncthreads();
    PT_MARK(3);

    if constexpr (POLICY == 0) {
    // ---- FF_tour layer 1: h1[r][n] = relu(b1[n] + W1veh^T veh_r + W1max^T m_r); 8 rows x 8 cols per thread,
    //      accumulators packed as float2 pairs and updated with FFMA2 (2 fp32 FMAs per issue slot) ----
    {
      const int cg = tid & 63, rg = tid >> 6;
      const int c0 = cg * 8, r0 = rg * 8;
      float2 acc[8][4];
      {
        float2 bb[4], v0[4], v1[4], v2[4];
        *reinterpret_cast<float4*>(&bb[0]) = ldg4(w.b1 + c0); *reinterpret_cast<float4*>(&bb[2]) = ldg4(w.b1 + c0 + 4);
        *reinterpret_cast<float4*>(&v0[0]) = ldg4(w.w1_veh + c0); *reinterpret_cast<float4*>(&v0[2]) = ldg4(w.w1_veh + c0 + 4);
        *reinterpret_cast<float4*>(&v1[0]) = ldg4(w.w1_veh + FF + c0); *reinterpret_cast<float4*>(&v1[2]) = ldg4(w.w1_veh + FF + c0 + 4);
        *reinterpret_cast<float4*>(&v2[0]) = ldg4(w.w1_veh + 2 * FF + c0); *reinterpret_cast<float4*>(&v2[2]) = ldg4(w.w1_veh + 2 * FF + c0 + 4);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float4 vh = *reinterpret_cast<const float4*>(&sm.veh[r0 + r][0]);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[r][c] = bb[c];
            ffma2s(acc[r][c], v0[c], vh.x); ffma2s(acc[r][c], v1[c], vh.y); ffma2s(acc[r][c], v2[c], vh.z);
          }
        }
      }
      const float* wp = w.w1_max + c0;
#pragma unroll 2
      for (int k = 0; k < D; k += 2) {
        float2 wa[4], wb[4];
        *reinterpret_cast<float4*>(&wa[0]) = ldg4(wp + (size_t)k * FF); *reinterpret_cast<float4*>(&wa[2]) = ldg4(wp + (size_t)k * FF + 4);
        *reinterpret_cast<float4*>(&wb[0]) = ldg4(wp + (size_t)(k + 1) * FF); *reinterpret_cast<float4*>(&wb[2]) = ldg4(wp + (size_t)(k + 1) * FF + 4);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float2 a = *reinterpret_cast<const float2*>(&sm.m[r0 + r][k]);
#pragma unroll
          for (int c = 0; c < 4; ++c) ffma2s(acc[r][c], wa[c], a.x);
#pragma unroll
          for (int c = 0; c < 4; ++c) ffma2s(acc[r][c], wb[c], a.y);
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        *reinterpret_cast<float4*>(&sm.u.h1[r0 + r][c0]) =
            make_float4(fmaxf(acc[r][0].x, 0.f), fmaxf(acc[r][0].y, 0.f), fmaxf(acc[r][1].x, 0.f), fmaxf(acc[r][1].y, 0.f));
        *reinterpret_cast<float4*>(&sm.u.h1[r0 + r][c0 + 4]) =
            make_float4(fmaxf(acc[r][2].x, 0.f), fmaxf(acc[r][2].y, 0.f), fmaxf(acc[r][3].x, 0.f), fmaxf(acc[r][3].y, 0.f));
      }
    }
    __syncthreads();
    PT_MARK(0);
    // ---- FF_tour layer 2: ctx[r][d] = W2^T h1_r (b2 added in the query); 8 rows x 4 cols x K-half per thread ----
    {
      const int kh = tid & 1, cg = (tid >> 1) & 31, rg = tid >> 6;
      const int c0 = cg * 4, r0 = rg * 8;
      float2 acc[8][2];
#pragma unroll
      for (int r = 0; r < 8; ++r) { acc[r][0] = make_float2(0.f, 0.f); acc[r][1] = make_float2(0.f, 0.f); }
      const float* wp = w.w2 + (size_t)kh * 256 * D + c0;
      const int kbase = kh * 256;
#pragma unroll 1
      for (int k = 0; k < 256; k += 4) {
        float2 ww[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(&ww[i][0]) = ldg4(wp + (size_t)(k + i) * D);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float4 hh = *reinterpret_cast<const float4*>(&sm.u.h1[r0 + r][kbase + k]);
          ffma2s(acc[r][0], ww[0][0], hh.x); ffma2s(acc[r][1], ww[0][1], hh.x);
          ffma2s(acc[r][0], ww[1][0], hh.y); ffma2s(acc[r][1], ww[1][1], hh.y);
          ffma2s(acc[r][0], ww[2][0], hh.z); ffma2s(acc[r][1], ww[2][1], hh.z);
          ffma2s(acc[r][0], ww[3][0], hh.w); ffma2s(acc[r][1], ww[3][1], hh.w);
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {                   // the two K halves sit in lane pairs
          acc[r][c].x += __shfl_xor_sync(FULL, acc[r][c].x, 1);
          acc[r][c].y += __shfl_xor_sync(FULL, acc[r][c].y, 1);
        }
      }
      __syncthreads();                                   // every read of h1 is done: its storage becomes ctx + scratch
      if (kh == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
          *reinterpret_cast<float4*>(&sm.u.att.ctx[r0 + r][c0]) = make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y);
      }
    }
    }
    __syncthreads();
    PT_MARK(1);

    // ---- per-row phase: each warp walks its rows ----
    WarpScratch& ws = sm.u.att.ws[wid];
    const int hd = lane >> 2, sub = lane & 3;
#pragma unroll 1
    while (true) {
      int lr = 0;
      if (lane == 0) lr = atomicAdd(&sm.next_row, 1);
      lr = __shfl_sync(FULL, lr, 0);
      if (lr >= rows_per_cta) break;
      RowState& rs = sm.rs[lr];
      if (rs.done) continue;                             // warp-uniform
      const int row = blockIdx.x * rows_per_cta + lr;
      const int inst = rs.inst;
      const int now = rs.now;
      const Geo<COORDS> geo = make_geo<COORDS>(Dm, Gm, xyp, elevp, inst, S);
      const float* embb = emb + (size_t)inst * S * D;
      const float* kvlb = kvl + (size_t)inst * S * 384;
      uint32_t mk[4] = {rs.mk[0], rs.mk[1], rs.mk[2], rs.mk[3]};
      // ---- feasible list (ascending node index) ----
      int nf = 0;
#pragma unroll
      for (int jb = 0; jb < 4; ++jb) {
        const uint32_t bits = mk[jb];
        if ((bits >> lane) & 1u) ws.list[nf + __popc(bits & ((1u << lane) - 1u))] = (unsigned char)(jb * 32 + lane);
        nf += __popc(bits);
      }
      __syncwarp();
      // V and logit-K rows of this step: warm L2 while the compat pass runs
      for (int i = lane; i < nf; i += 32) {
        const float* p = kvlb + (size_t)ws.list[i] * 384 + 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) prefetch_l2(p + c * 32);
      }
      // ---- query (:394): (fixed_ctx + ctx) + emb[now] ; lane holds 4 channels of head lane/4 ----
      float q[4];
      {
        const float4 fc = ldg4(fixed_ctx + (size_t)inst * D + lane * 4);
        const float4 en = ldg4(embb + (size_t)now * D + lane * 4);
        if constexpr (POLICY == 0) {
          const float4 cx = *reinterpret_cast<const float4*>(&sm.u.att.ctx[lr][lane * 4]);
          const float4 b2 = ldg4(w.b2 + lane * 4);
          q[0] = (fc.x + (cx.x + b2.x)) + en.x;
          q[1] = (fc.y + (cx.y + b2.y)) + en.y;
          q[2] = (fc.z + (cx.z + b2.z)) + en.z;
          q[3] = (fc.w + (cx.w + b2.w)) + en.w;
        } else {                                       // nets/AM.py:336-337: en = step_tab[now], + load * w[:, 128]
          const float4 wl = ldg4(w.am_w_load + lane * 4);
          const float ld = rs.load;
          q[0] = fc.x + fmaf(ld, wl.x, en.x);
          q[1] = fc.y + fmaf(ld, wl.y, en.y);
          q[2] = fc.z + fmaf(ld, wl.z, en.z);
          q[3] = fc.w + fmaf(ld, wl.w, en.w);
        }
      }
      // ---- compat[h][i] = q_h . K[j_i, h] / 4 ; CH rows (CH x 512 B) in flight per warp ----
      constexpr int CH = 8;
      for (int i0 = 0; i0 < nf; i0 += CH) {
        float4 kv[CH];
#pragma unroll
        for (int a = 0; a < CH; ++a) kv[a] = ldg4_pol(kvlb + (size_t)ws.list[min(i0 + a, nf - 1)] * 384 + lane * 4, pol);
#pragma unroll
        for (int a = 0; a < CH; ++a) {
          float p = fmaf(q[3], kv[a].w, fmaf(q[2], kv[a].z, fmaf(q[1], kv[a].y, q[0] * kv[a].x)));
          p += __shfl_xor_sync(FULL, p, 1);
          p += __shfl_xor_sync(FULL, p, 2);
          if (sub == 0 && i0 + a < nf) ws.pc[hd][i0 + a] = p * 0.25f;
        }
      }
      __syncwarp();
      // ---- per-head softmax over the list; lane (hd, sub) strides i = sub, sub+4, ... ----
      {
        float mx = -INFINITY;
        for (int i = sub; i < nf; i += 4) mx = fmaxf(mx, ws.pc[hd][i]);
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 2));
        float sum = 0.f;
        for (int i = sub; i < nf; i += 4) { const float e = expf(ws.pc[hd][i] - mx); ws.pc[hd][i] = e; sum += e; }
        sum += __shfl_xor_sync(FULL, sum, 1);
        sum += __shfl_xor_sync(FULL, sum, 2);
        for (int i = sub; i < nf; i += 4) ws.pc[hd][i] = __fdiv_rn(ws.pc[hd][i], sum);
      }
      __syncwarp();
      // ---- heads = attn @ V ; lane owns 4 channels ----
      {
        float hv[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i0 = 0; i0 < nf; i0 += CH) {
          float4 vv[CH]; float pr[CH];
#pragma unroll
          for (int a = 0; a < CH; ++a) {
            const int ii = min(i0 + a, nf - 1);
            vv[a] = ldg4_pol(kvlb + (size_t)ws.list[ii] * 384 + 128 + lane * 4, pol);
            pr[a] = (i0 + a < nf) ? ws.pc[hd][ii] : 0.f;
          }
#pragma unroll
          for (int a = 0; a < CH; ++a) {
            hv[0] = fmaf(pr[a], vv[a].x, hv[0]); hv[1] = fmaf(pr[a], vv[a].y, hv[1]);
            hv[2] = fmaf(pr[a], vv[a].z, hv[2]); hv[3] = fmaf(pr[a], vv[a].w, hv[3]);
          }
        }
        *reinterpret_cast<float4*>(&ws.heads[lane * 4]) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      }
      __syncwarp();
      // ---- logits[i] = heads . lK[j_i] / sqrt(128) -> tanh clip -> / temp  (project_out folded into lK).
      //      8 lanes per node: lane (grp, s8) covers channels 4*s8 + 32*c, c = 0..3; 4 nodes per pass ----
      {
        const int grp = lane >> 3, s8 = lane & 7;
        float4 hq[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) hq[c] = *reinterpret_cast<const float4*>(&ws.heads[s8 * 4 + c * 32]);
        for (int i0 = 0; i0 < nf; i0 += 16) {            // 4 nodes per sub-pass, 4 sub-passes in flight
          float4 lv[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ii = min(i0 + u * 4 + grp, nf - 1);
            const float* p = kvlb + (size_t)ws.list[ii] * 384 + 256 + s8 * 4;
#pragma unroll
            for (int c = 0; c < 4; ++c) lv[u][c] = ldg4_pol(p + c * 32, pol);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              acc = fmaf(hq[c].w, lv[u][c].w, fmaf(hq[c].z, lv[u][c].z, fmaf(hq[c].y, lv[u][c].y, fmaf(hq[c].x, lv[u][c].x, acc))));
            acc += __shfl_xor_sync(FULL, acc, 1);
            acc += __shfl_xor_sync(FULL, acc, 2);
            acc += __shfl_xor_sync(FULL, acc, 4);
            const int i = i0 + u * 4 + grp;
            if (s8 == 0 && i < nf) {
              float z = __fdiv_rn(acc, sqrt_d);
              if (cfg.tanh_clipping > 0.f) z = tanhf(z) * cfg.tanh_clipping;
              ws.logit[i] = __fdiv_rn(z, cfg.temp);
            }
          }
        }
      }
      __syncwarp();
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

      // ---- loads for the transition and the new mask, issued together ----
      float dist, slp;
      geo.at(now, c_node, dist, slp);
      float xc = 0.f, yc = 0.f, ec_elev = 0.f;                  // coordinate mode: the chosen node, row D[c,:] / G[c,:] on the fly
      if (geo.coords()) { xc = __ldg(geo.x + c_node); yc = __ldg(geo.y + c_node); ec_elev = __ldg(geo.e + c_node); }
      float4 ec = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (POLICY == 0) ec = ldg4(embb + (size_t)c_node * D + lane * 4);
      float dcj[4], gcj[4], d0j[4], g0j[4], d3j[4], s3j[4], dmj[4];
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
        const float dmc = sm.dem[lr][c_node];
        const float nl = fmaxf(__fsub_rn(load, dmc), 0.f);
        const float nd = fmaxf(__fsub_rn(dmc, load), 0.f);
        __syncwarp();
        if (lane == 0) { sm.dem[lr][c_node] = nd; sm.dem[lr][0] = __fadd_rn(-1.f, nl); }
        load = nl;
      }
      if (is_depot) { load = 1.f; if (lane == 0) sm.dem[lr][0] = 0.f; }
      __syncwarp();
      if (lane == 0) {
        const size_t o = (size_t)row * Tcap + t;
        pi[o] = c_node; ll[o] = lp_sel;
        reward[o] = phys.objective ? dist : e;
        if (energy) energy[o] = e;
      }
      // ---- running max of visited embeddings (route_feature :207-214) ----
      if constexpr (POLICY == 0) {
        float4 mm = *reinterpret_cast<float4*>(&sm.m[lr][lane * 4]);
        if (t == 0) mm = ec;
        else { mm.x = fmaxf(mm.x, ec.x); mm.y = fmaxf(mm.y, ec.y); mm.z = fmaxf(mm.z, ec.z); mm.w = fmaxf(mm.w, ec.w); }
        *reinterpret_cast<float4*>(&sm.m[lr][lane * 4]) = mm;
      }
      // ---- new mask (problems/EVRP.py:105-223) ----
      bool any_dem = false, cust_dem = false;
#pragma unroll
      for (int jb = 0; jb < 4; ++jb) {
        const int j = jb * 32 + lane;
        dmj[jb] = sm.dem[lr][j];
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
            if (mbit && !reachable(phys, j, F, load, load, dm, soc, tim, sel4(dcj, jb), sel4(gcj, jb), sel4(d0j, jb),
                                   sel4(g0j, jb), sel4(d3j, jb), sel4(s3j, jb), d4)) mbit = false;
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
      }
      __syncwarp();
    }
    PT_MARK(2);
    ++t;
  }
#ifdef EVRP_PHASE_TIMING
  if (tid == 0) for (int i = 0; i < 4; ++i) atomicAdd(&g_phase_cycles[i], (unsigned long long)pt_acc[i]);
#endif
}

// sample_many reduction (nets/DRLModel.py:301-311): min / argmin cost over the R replicas of each instance
// (first replica on ties, torch.min semantics) and its tour.  One warp per instance.
__global__ void __launch_bounds__(256)
sample_min_kernel(const float* __restrict__ reward, const int64_t* __restrict__ pi, int B, int R, int Tcap,
                  float* __restrict__ mincost, int32_t* __restrict__ argmin, int64_t* __restrict__ minpi) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  float best = INFINITY; int bi = 0x7fffffff;
  for (int r = lane; r < R; r += 32) {
    const float* rw = reward + ((size_t)r * B + b) * Tcap;
    float s = 0.f;
    for (int t = 0; t < Tcap; ++t) s += rw[t];          // cost = R.sum(1) (:301)
    if (s < best || (s == best && r < bi)) { best = s; bi = r; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(FULL, best, o);
    const int oi = __shfl_xor_sync(FULL, bi, o);
    if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) { mincost[b] = best; argmin[b] = bi; }
  const int64_t* src = pi + ((size_t)bi * B + b) * Tcap;
  for (int t = lane; t < Tcap; t += 32) minpi[(size_t)b * Tcap + t] = src[t];
}

// evrp_decode_tc.cu
#ifdef EVRP_PHASE_TIMING
int debug_phase_cycles_tc(unsigned long long* out4, int reset);
int debug_row_cycles_tc(unsigned long long* out8, int reset);
int debug_ff_cycles_tc(unsigned long long* out12, int reset);
#endif
int launch_rollout_tc(const evrp_phys* phys, const evrp_decoder_weights* w, const evrp_rollout_cfg* cfg, const float* emb,
                      const float* kvl, const float* fixed_ctx, const float* dynamic0, const float* distances,
                      const float* slope, const float* static_xy, const float* elevations, int B, int S, int F,
                      const int64_t* forced_actions, const float* uniforms,
                      int64_t* pi, float* ll, float* reward, float* energy, uint32_t* saved_mask, float* saved_veh,
                      int32_t* steps_used, int32_t* t_max, int32_t* status, float* station_ws, float* dem_ws, float* m_ws,
                      cudaStream_t st);

}  // namespace evrp

extern "C" {

size_t evrp_rollout_workspace_bytes(int B, int S, int n_replicas) {
  (void)S;
  const size_t rows = (size_t)B * (n_replicas > 0 ? n_replicas : 1);
  // instance constants [B][4][128] (rule-7 d3, s3; D[0,:], G[0,:]) | demand rows [rows][128] | exact running-max rows [rows][128]
  return ((size_t)B * 4 * EVRP_MAX_S + rows * EVRP_MAX_S + rows * EVRP_D) * sizeof(float);
}

int evrp_rollout(const evrp_phys* phys, const evrp_decoder_weights* w, const evrp_rollout_cfg* cfg,
                 const float* emb, const float* kvl, const float* fixed_ctx, const float* dynamic0,
                 const float* distances, const float* slope, int B, int S, int F, const int64_t* forced_actions,
                 const float* uniforms, int64_t* pi, float* ll, float* reward, float* energy, uint32_t* saved_mask,
                 float* saved_veh, int32_t* steps_used, int32_t* t_max, int32_t* status, void* workspace,
                 size_t workspace_bytes, void* stream) {
  using namespace evrp;
  // geometry: the [B,S,S] matrices, or (distances == slope == NULL) the coordinate / elevation rows in the configuration
  const float* static_xy = cfg ? cfg->static_xy : nullptr;
  const float* elevations = cfg ? cfg->elevations : nullptr;
  const bool geo_ok = (distances && slope) || (!distances && !slope && static_xy && elevations);
  if (!phys || !w || !cfg || !emb || !kvl || !fixed_ctx || !dynamic0 || !geo_ok || !pi || !ll ||
      !reward || !steps_used || !t_max || !status || !workspace || B < 0 || S < 2 || F < 1 || F >= S ||
      cfg->max_steps < 1 || cfg->n_replicas < 1 || !(cfg->temp > 0.f))
    return -EVRP_E_BADARG;
  if (S > EVRP_MAX_S) return -EVRP_E_TOO_MANY_NODES;
  if (workspace_bytes < evrp_rollout_workspace_bytes(B, S, cfg->n_replicas)) return -EVRP_E_WORKSPACE;
  if (B == 0) return 0;
  const int rows = B * cfg->n_replicas;
  cudaStream_t st = (cudaStream_t)stream;
  if (w->policy == 1 && !w->am_w_load) return -EVRP_E_BADARG;
  if (w->policy != 1 && (!w->w1_veh || !w->w1_max || !w->b1 || !w->w2 || !w->b2)) return -EVRP_E_BADARG;
  if (w->ff_mode == 0 && w->policy != 1) {
    float* station_ws = (float*)workspace;
    float* dem_ws = station_ws + (size_t)B * 4 * EVRP_MAX_S;
    float* m_ws = dem_ws + (size_t)rows * EVRP_MAX_S;
    return launch_rollout_tc(phys, w, cfg, emb, kvl, fixed_ctx, dynamic0, distances, slope, static_xy, elevations, B, S, F, forced_actions,
                             uniforms, pi, ll, reward, energy, saved_mask, saved_veh, steps_used, t_max, status,
                             station_ws, dem_ws, m_ws, st);
  }
  const size_t smem = sizeof(DecodeSmem);
  const int mode = forced_actions ? 2 : (cfg->decode_type == EVRP_DECODE_SAMPLE ? 1 : 0);
  auto pick = [&](auto coords_tag) {
    constexpr bool C = decltype(coords_tag)::value;
    return w->policy == 1 ? (mode == 2 ? rollout_kernel<2, 1, C> : mode == 1 ? rollout_kernel<1, 1, C> : rollout_kernel<0, 1, C>)
                          : (mode == 2 ? rollout_kernel<2, 0, C> : mode == 1 ? rollout_kernel<1, 0, C> : rollout_kernel<0, 0, C>);
  };
  auto kern = distances ? pick(std::false_type{}) : pick(std::true_type{});
  EVRP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // balance: fill whole waves of (2 CTAs x SMs) slots and spread the rows evenly over them (<= RPC rows per CTA)
  int dev = 0, sms = 148;
  EVRP_CUDA_OK(cudaGetDevice(&dev));
  EVRP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int slots = 2 * sms;
  const int min_ctas = (rows + RPC - 1) / RPC;
  const int waves = (min_ctas + slots - 1) / slots;
  int grid = waves * slots;
  if (grid > rows) grid = rows;
  const int rows_per_cta = (rows + grid - 1) / grid;
  grid = (rows + rows_per_cta - 1) / rows_per_cta;
  kern<<<grid, NT, smem, st>>>(*phys, *w, *cfg, emb, kvl, fixed_ctx, dynamic0, distances, slope, static_xy, elevations,
                                         (float*)workspace, B, S, F, rows, rows_per_cta, forced_actions, uniforms, pi, ll, reward,
                                         energy, saved_mask, saved_veh, steps_used, t_max, status);
  EVRP_LAUNCH_OK();
  return 0;
}

#ifdef EVRP_PHASE_TIMING
int evrp_debug_phase_cycles_tc(unsigned long long* out4, int reset) { return evrp::debug_phase_cycles_tc(out4, reset); }
int evrp_debug_row_cycles_tc(unsigned long long* out8, int reset) { return evrp::debug_row_cycles_tc(out8, reset); }
int evrp_debug_ff_cycles_tc(unsigned long long* out12, int reset) { return evrp::debug_ff_cycles_tc(out12, reset); }
int evrp_debug_phase_cycles(unsigned long long* out4, int reset) {
  if (cudaMemcpyFromSymbol(out4, evrp::g_phase_cycles, sizeof(unsigned long long) * 4) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(evrp::g_phase_cycles, z, sizeof(z)); }
  return 0;
}
#endif

int evrp_sample_min(const float* reward, const int64_t* pi, int B, int R, int Tcap, float* mincost, int32_t* argmin,
                    int64_t* minpi, void* stream) {
  if (!reward || !pi || !mincost || !argmin || !minpi || B < 0 || R < 1 || Tcap < 1) return -EVRP_E_BADARG;
  if (B == 0) return 0;
  evrp::sample_min_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(reward, pi, B, R, Tcap, mincost, argmin, minpi);
  EVRP_LAUNCH_OK();
  return 0;
}

}  // extern "C"
