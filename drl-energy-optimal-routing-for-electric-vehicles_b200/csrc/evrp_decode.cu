// Persistent rollout kernel: ONE launch decodes whole tours.
//
//   AttentionModel._inner            nets/DRLModel.py:240-280   (the step loop lives inside the kernel)
//   route_feature                    nets/DRLModel.py:203-237   (running max + FF_tour)
//   _get_log_p / _one_to_many_logits nets/DRLModel.py:379-452   (8-head masked glimpse, pointer logits)
//   _select_node                     nets/DRLModel.py:316-342   (greedy first-index argmax / sampling)
//   _calc_log_likelihood             nets/DRLModel.py:182-190
//   update_dynamic / update_mask     problems/EVRP.py:105-280   (fused, bit-exact recipe of evrp_common.cuh)
//
// Layout: a CTA owns G=8 rollout rows (one warp each) for their whole life.  Per step
//   1. CTA-wide batched FF_tour GEMMs over the 8 rows (weights streamed once per CTA-step from L2, coalesced,
//      shared by the 8 rows),
//   2. per-warp glimpse attention / logits over the FEASIBLE nodes only (compact index list from the 128-bit
//      mask; K|V|logitK rows are 512-B coalesced float4 loads, issued in batches of 4 for memory-level
//      parallelism; next step's K rows are L2-prefetched as soon as the new mask is known),
//   3. warp-level masked log-softmax, argmax / inverse-CDF sample,
//   4. state transition + look-ahead feasibility mask with the dynamic state in registers/shared memory.
// Rows retire individually; only t_max and the status word are read back by the host.
#include <math.h>
#include "evrp_common.cuh"

namespace evrp {

constexpr int G = 8;             // rollout rows per CTA (one warp each)
constexpr int NT = G * 32;       // 256 threads
constexpr int SP = EVRP_MAX_S;   // padded node count (128)
constexpr int PSTR = SP + 4;     // head stride of the compat buffer (bank-conflict-free for the 4-lane groups)

struct __align__(16) DecodeSmem {
  float m[G][D];            // running max of visited embeddings          (route_feature :214)
  float veh[G][4];          // load, soc/Start_SOC, time/t_limit          (:215-222)
  float h1[G][FF];          // FF_tour hidden
  float ctxp[2][G][D];      // FF_tour output, two K-halves
  float pc[G][H][PSTR];     // compat -> attention probabilities, by list position
  float logit[G][SP];       // clipped logits -> log_p, by list position
  float dem[G][SP];         // demand row
  float D0[G][SP], G0[G][SP], d3[G][SP], s3[G][SP];   // depot row + rule-7 constants
  float dcur[G][SP], gcur[G][SP];                     // D[now,:], G[now,:]
  unsigned char list[G][SP];                          // feasible node indices, ascending
};

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// register-array select without dynamic indexing (keeps the arrays out of local memory)
template <typename T>
__device__ __forceinline__ T sel4(const T (&a)[4], int i) { return i == 0 ? a[0] : i == 1 ? a[1] : i == 2 ? a[2] : a[3]; }

__global__ void __launch_bounds__(NT, 2)
rollout_kernel(evrp_phys phys, evrp_decoder_weights w, evrp_rollout_cfg cfg,
               const float* __restrict__ emb, const float* __restrict__ kvl, const float* __restrict__ fixed_ctx,
               const float* __restrict__ dynamic0, const float* __restrict__ Dm, const float* __restrict__ Gm,
               int B, int S, int F, int rows_total,
               const int64_t* __restrict__ forced, const float* __restrict__ uniforms,
               int64_t* __restrict__ pi, float* __restrict__ ll, float* __restrict__ reward,
               float* __restrict__ energy, uint32_t* __restrict__ saved_mask, float* __restrict__ saved_veh,
               int32_t* __restrict__ steps_used, int32_t* __restrict__ t_max, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DecodeSmem& sm = *reinterpret_cast<DecodeSmem*>(smem_raw);
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const int row = blockIdx.x * G + wid;
  const bool valid = row < rows_total;
  const int inst = valid ? row % B : 0;
  const int Tcap = cfg.max_steps;
  const float* Db = Dm + (size_t)inst * S * S;
  const float* Gb = Gm + (size_t)inst * S * S;
  const float* embb = emb + (size_t)inst * S * D;
  const float* kvlb = kvl + (size_t)inst * S * 384;
  const int nblk = (S + 31) >> 5;

  // ---------------- prologue: per-row state ----------------
  float load = 1.f, soc = 0.f, tim = 0.f;
  uint32_t mk[4] = {0u, 0u, 0u, 0u};
  int now = 0;
  bool done = !valid;
  float d4 = 0.f;
  {
    const float* dy = dynamic0 + (size_t)inst * 4 * S;
    load = dy[0]; soc = dy[2 * S]; tim = dy[3 * S];
    bool nonuni = false;
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {
      const int j = jb * 32 + lane;
      float dmv = 0.f, D0v = 0.f, G0v = 0.f, d3v = 0.f, s3v = 0.f;
      if (j < S) {
        dmv = dy[S + j];
        nonuni |= (dy[j] != load) || (dy[2 * S + j] != soc) || (dy[3 * S + j] != tim);
        D0v = Db[j]; G0v = Gb[j];
        int k = 0; float best = Db[(size_t)1 * S + j];
        for (int i = 1; i < F; ++i) { const float v = Db[(size_t)(1 + i) * S + j]; if (v < best) { best = v; k = i; } }
        d3v = (j == 0) ? 0.f : best;
        s3v = Gb[(size_t)(1 + k) * S + j];
        if (j == 0) d4 = Db[(size_t)(1 + k) * S + 0];
      }
      sm.dem[wid][j] = dmv; sm.D0[wid][j] = D0v; sm.G0[wid][j] = G0v; sm.d3[wid][j] = d3v; sm.s3[wid][j] = s3v;
      sm.dcur[wid][j] = D0v; sm.gcur[wid][j] = G0v;           // now = 0 at t = 0 (:257)
      mk[jb] = __ballot_sync(FULL, j < S);                     // mask = ones (:256)
    }
    d4 = __shfl_sync(FULL, d4, 0);
    if (valid && __any_sync(FULL, nonuni) && lane == 0) atomicOr(status, EVRP_ST_NONUNIFORM_DYNAMIC);
#pragma unroll
    for (int i = 0; i < 4; ++i) sm.m[wid][lane * 4 + i] = 0.f;   // zeros before the first pick (:224)
  }
  const float sqrt_d = sqrtf((float)D);
  int t = 0;
  bool first = true;

  // ---------------- step loop ----------------
  while (true) {
    if (__syncthreads_and(done ? 1 : 0)) break;
    if (!done && lane == 0) {
      sm.veh[wid][0] = load;
      sm.veh[wid][1] = __fdiv_rn(soc, phys.start_soc);
      sm.veh[wid][2] = __fdiv_rn(tim, phys.t_limit);
      if (saved_veh) {
        float* sv = saved_veh + ((size_t)row * Tcap + t) * 3;
        sv[0] = load; sv[1] = sm.veh[wid][1]; sv[2] = sm.veh[wid][2];
      }
    }
    if (!done && saved_mask && lane < 4) saved_mask[((size_t)row * Tcap + t) * 4 + lane] = sel4(mk, lane);
    __syncthreads();

    // ---- FF_tour layer 1: h1[g][n] = relu(b1 + W1veh^T veh_g + W1max^T m_g), n = tid, tid+256 ----
    {
      float acc0[G], acc1[G];
      const float ba = __ldg(w.b1 + tid), bb = __ldg(w.b1 + 256 + tid);
      const float va0 = __ldg(w.w1_veh + tid), va1 = __ldg(w.w1_veh + FF + tid), va2 = __ldg(w.w1_veh + 2 * FF + tid);
      const float vb0 = __ldg(w.w1_veh + 256 + tid), vb1 = __ldg(w.w1_veh + FF + 256 + tid), vb2 = __ldg(w.w1_veh + 2 * FF + 256 + tid);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float x0 = sm.veh[g][0], x1 = sm.veh[g][1], x2 = sm.veh[g][2];
        acc0[g] = fmaf(va2, x2, fmaf(va1, x1, fmaf(va0, x0, ba)));
        acc1[g] = fmaf(vb2, x2, fmaf(vb1, x1, fmaf(vb0, x0, bb)));
      }
      const float* wp = w.w1_max + tid;
#pragma unroll 2
      for (int k = 0; k < D; k += 4) {
        float wa[4], wb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { wa[i] = __ldg(wp + (size_t)(k + i) * FF); wb[i] = __ldg(wp + (size_t)(k + i) * FF + 256); }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float4 mm = *reinterpret_cast<const float4*>(&sm.m[g][k]);
          acc0[g] = fmaf(wa[0], mm.x, acc0[g]); acc1[g] = fmaf(wb[0], mm.x, acc1[g]);
          acc0[g] = fmaf(wa[1], mm.y, acc0[g]); acc1[g] = fmaf(wb[1], mm.y, acc1[g]);
          acc0[g] = fmaf(wa[2], mm.z, acc0[g]); acc1[g] = fmaf(wb[2], mm.z, acc1[g]);
          acc0[g] = fmaf(wa[3], mm.w, acc0[g]); acc1[g] = fmaf(wb[3], mm.w, acc1[g]);
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) { sm.h1[g][tid] = fmaxf(acc0[g], 0.f); sm.h1[g][256 + tid] = fmaxf(acc1[g], 0.f); }
    }
    __syncthreads();
    // ---- FF_tour layer 2: ctx[g][d] = b2 + W2^T h1_g ; thread = (d, K-half) ----
    {
      const int d = tid & 127, half = tid >> 7;
      float acc[G];
#pragma unroll
      for (int g = 0; g < G; ++g) acc[g] = 0.f;
      const float* wp = w.w2 + (size_t)half * 256 * D + d;
#pragma unroll 2
      for (int k = 0; k < 256; k += 4) {
        float ww[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ww[i] = __ldg(wp + (size_t)(k + i) * D);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float4 hh = *reinterpret_cast<const float4*>(&sm.h1[g][half * 256 + k]);
          acc[g] = fmaf(ww[0], hh.x, acc[g]); acc[g] = fmaf(ww[1], hh.y, acc[g]);
          acc[g] = fmaf(ww[2], hh.z, acc[g]); acc[g] = fmaf(ww[3], hh.w, acc[g]);
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) sm.ctxp[half][g][d] = acc[g];
    }
    __syncthreads();

    if (!done) {
      // ---- feasible list (ascending node index) ----
      int nf = 0;
#pragma unroll
      for (int jb = 0; jb < 4; ++jb) {
        const uint32_t bits = mk[jb];
        if ((bits >> lane) & 1u) sm.list[wid][nf + __popc(bits & ((1u << lane) - 1u))] = (unsigned char)(jb * 32 + lane);
        nf += __popc(bits);
      }
      __syncwarp();
      // V and logit-K rows of this step: warm L2 while the compat pass runs
      for (int i = lane; i < nf; i += 32) {
        const float* r = kvlb + (size_t)sm.list[wid][i] * 384 + 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) prefetch_l2(r + c * 32);
      }
      // ---- query (:394): (fixed_ctx + ctx) + emb[now] ; lane holds 4 channels of head lane/4 ----
      float q[4];
      {
        const float4 fc = ldg4(fixed_ctx + (size_t)inst * D + lane * 4);
        const float4 en = ldg4(embb + (size_t)now * D + lane * 4);
        const float4 c0 = *reinterpret_cast<const float4*>(&sm.ctxp[0][wid][lane * 4]);
        const float4 c1 = *reinterpret_cast<const float4*>(&sm.ctxp[1][wid][lane * 4]);
        const float4 b2 = ldg4(w.b2 + lane * 4);
        q[0] = (fc.x + ((c0.x + c1.x) + b2.x)) + en.x;
        q[1] = (fc.y + ((c0.y + c1.y) + b2.y)) + en.y;
        q[2] = (fc.z + ((c0.z + c1.z) + b2.z)) + en.z;
        q[3] = (fc.w + ((c0.w + c1.w) + b2.w)) + en.w;
      }
      const int hd = lane >> 2, sub = lane & 3;
      // ---- compat[h][i] = q_h . K[j_i, h] / 4 ----
      for (int i0 = 0; i0 < nf; i0 += 4) {
        float4 kv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int j = sm.list[wid][min(i0 + a, nf - 1)];
          kv[a] = ldg4(kvlb + (size_t)j * 384 + lane * 4);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          float p = fmaf(q[3], kv[a].w, fmaf(q[2], kv[a].z, fmaf(q[1], kv[a].y, q[0] * kv[a].x)));
          p += __shfl_xor_sync(FULL, p, 1);
          p += __shfl_xor_sync(FULL, p, 2);
          if (sub == 0 && i0 + a < nf) sm.pc[wid][hd][i0 + a] = p * 0.25f;
        }
      }
      __syncwarp();
      // ---- per-head softmax over the list; lane (hd, sub) strides i = sub, sub+4, ... ----
      {
        float mx = -INFINITY;
        for (int i = sub; i < nf; i += 4) mx = fmaxf(mx, sm.pc[wid][hd][i]);
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 2));
        float sum = 0.f;
        for (int i = sub; i < nf; i += 4) { const float e = expf(sm.pc[wid][hd][i] - mx); sm.pc[wid][hd][i] = e; sum += e; }
        sum += __shfl_xor_sync(FULL, sum, 1);
        sum += __shfl_xor_sync(FULL, sum, 2);
        for (int i = sub; i < nf; i += 4) sm.pc[wid][hd][i] = __fdiv_rn(sm.pc[wid][hd][i], sum);
      }
      __syncwarp();
      // ---- heads = attn @ V ; lane keeps its 4 channels ----
      float hv[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i0 = 0; i0 < nf; i0 += 4) {
        float4 vv[4]; float pr[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int ii = min(i0 + a, nf - 1);
          const int j = sm.list[wid][ii];
          vv[a] = ldg4(kvlb + (size_t)j * 384 + 128 + lane * 4);
          pr[a] = (i0 + a < nf) ? sm.pc[wid][hd][ii] : 0.f;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          hv[0] = fmaf(pr[a], vv[a].x, hv[0]); hv[1] = fmaf(pr[a], vv[a].y, hv[1]);
          hv[2] = fmaf(pr[a], vv[a].z, hv[2]); hv[3] = fmaf(pr[a], vv[a].w, hv[3]);
        }
      }
      // ---- logits[i] = heads . lK[j_i] / sqrt(128) -> tanh clip -> / temp   (project_out folded into lK) ----
      for (int i0 = 0; i0 < nf; i0 += 4) {
        float4 lv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int j = sm.list[wid][min(i0 + a, nf - 1)];
          lv[a] = ldg4(kvlb + (size_t)j * 384 + 256 + lane * 4);
        }
        float p[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) p[a] = fmaf(hv[3], lv[a].w, fmaf(hv[2], lv[a].z, fmaf(hv[1], lv[a].y, hv[0] * lv[a].x)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int a = 0; a < 4; ++a) p[a] += __shfl_xor_sync(FULL, p[a], o);
        }
        if (lane < 4 && i0 + lane < nf) {
          float z = __fdiv_rn(sel4(p, lane), sqrt_d);
          if (cfg.tanh_clipping > 0.f) z = tanhf(z) * cfg.tanh_clipping;
          sm.logit[wid][i0 + lane] = __fdiv_rn(z, cfg.temp);
        }
      }
      __syncwarp();
      // ---- log_softmax over the list (:403) ----
      float zl[4];
      float mx = -INFINITY;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = r * 32 + lane;
        zl[r] = (i < nf) ? sm.logit[wid][i] : -INFINITY;
        mx = fmaxf(mx, zl[r]);
      }
      mx = warp_max(mx);
      float se = 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) if (r * 32 + lane < nf) se += expf(zl[r] - mx);
      se = warp_sum(se);
      const float lgs = logf(se);
      const float lse = lgs + mx;
      float lp[4], pr[4];
      bool bad = false;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        lp[r] = (zl[r] - mx) - lgs;                            // x - max - log(sum(exp(x - max)))
        pr[r] = (r * 32 + lane < nf) ? expf(lp[r]) : -1.f;     // probs = log_p.exp() (:265)
        bad |= (r * 32 + lane < nf) && !(lp[r] == lp[r]);
      }
      if (__any_sync(FULL, bad) && lane == 0) atomicOr(status, EVRP_ST_NAN_LOGIT);
      // ---- pick ----
      int sel_i = 0;
      if (forced) {
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
      } else if (cfg.decode_type == EVRP_DECODE_GREEDY) {
        // first-index argmax of probs (torch.max(1), :327)
        float bv = pr[0]; int bi = lane;
#pragma unroll
        for (int r = 1; r < 4; ++r) if (pr[r] > bv) { bv = pr[r]; bi = r * 32 + lane; }
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
          const float pv = (i < nf) ? expf(sm.logit[wid][i] - lse) : 0.f;
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
        c_node = sm.list[wid][sel_i];
        lp_sel = __shfl_sync(FULL, sel4(lp, sel_i >> 5), sel_i & 31);
      } else { c_node = -1 - sel_i; lp_sel = -INFINITY; }

      // ---- state transition (problems/EVRP.py:226-280) on warp-uniform scalars ----
      const float dist = sm.dcur[wid][c_node];
      const float slp = sm.gcur[wid][c_node];
      const float tc = travel_time(phys, dist);
      const float e = leg_energy(phys, vehicle_mass(phys, load), slp, tc);
      const bool is_depot = (c_node == 0), is_station = (c_node > 0 && c_node <= F), is_cust = (c_node > F);
      tim = __fsub_rn(tim, tc);
      if (is_depot) tim = phys.t_limit;
      if (is_station) tim = __fsub_rn(tim, phys.charging_time);
      if (is_cust) tim = __fsub_rn(tim, phys.serve_time);
      soc = __fsub_rn(soc, e);
      if (c_node <= F) soc = phys.start_soc;
      __syncwarp();
      if (is_cust) {
        const float dmc = sm.dem[wid][c_node];
        const float nl = fmaxf(__fsub_rn(load, dmc), 0.f);
        const float nd = fmaxf(__fsub_rn(dmc, load), 0.f);
        __syncwarp();
        if (lane == 0) { sm.dem[wid][c_node] = nd; sm.dem[wid][0] = __fadd_rn(-1.f, nl); }
        load = nl;
      }
      if (is_depot) { load = 1.f; if (lane == 0) sm.dem[wid][0] = 0.f; }
      __syncwarp();
      if (lane == 0) {
        const size_t o = (size_t)row * Tcap + t;
        pi[o] = c_node; ll[o] = lp_sel;
        reward[o] = phys.objective ? dist : e;
        if (energy) energy[o] = e;
      }
      // ---- running max of visited embeddings (route_feature :207-214) ----
      {
        const float4 ec = ldg4(embb + (size_t)c_node * D + lane * 4);
        float4 mm = *reinterpret_cast<float4*>(&sm.m[wid][lane * 4]);
        if (first) mm = ec;
        else { mm.x = fmaxf(mm.x, ec.x); mm.y = fmaxf(mm.y, ec.y); mm.z = fmaxf(mm.z, ec.z); mm.w = fmaxf(mm.w, ec.w); }
        *reinterpret_cast<float4*>(&sm.m[wid][lane * 4]) = mm;
        first = false;
      }
      // ---- new mask (problems/EVRP.py:105-223) ----
      bool any_dem = false, cust_dem = false;
      float dmj[4];
#pragma unroll
      for (int jb = 0; jb < 4; ++jb) {
        const int j = jb * 32 + lane;
        dmj[jb] = sm.dem[wid][j];
        any_dem |= (j < S) && (dmj[jb] != 0.f);
        cust_dem |= (j < S) && (j >= F) && (dmj[jb] != 0.f);
      }
      any_dem = __any_sync(FULL, any_dem);
      cust_dem = __any_sync(FULL, cust_dem);
      ++t;
      if (!any_dem) {
        done = true;                                   // this row's tour is complete (:127)
      } else {
        const bool combined = (load == 0.f) || !cust_dem;
        bool cust_ok = false;
        uint32_t nm[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int jb = 0; jb < 4; ++jb) {
          if (jb >= nblk) break;                         // warp-uniform
          const int j = jb * 32 + lane;
          bool mbit = false;
          if (j < S) {
            const float dcj = __ldg(Db + (size_t)c_node * S + j);
            const float gcj = __ldg(Gb + (size_t)c_node * S + j);
            sm.dcur[wid][j] = dcj; sm.gcur[wid][j] = gcj;
            const float dm = dmj[jb];
            mbit = (dm != 0.f) && (dm < load);
            if (j == c_node) mbit = false;
            if (c_node <= F && j >= 1 && j <= F) mbit = false;
            if (is_station && j == 0) mbit = true;
            if (is_cust && j <= F) mbit = true;
            if (is_depot && j == 0) mbit = false;
            if (combined) mbit = (j == 0);
            if (!reachable(phys, j, F, load, load, dm, soc, tim, dcj, gcj, sm.D0[wid][j], sm.G0[wid][j],
                           sm.d3[wid][j], sm.s3[wid][j], d4)) mbit = false;
            if (j > F && mbit) cust_ok = true;
          }
          nm[jb] = __ballot_sync(FULL, mbit);
        }
        cust_ok = __any_sync(FULL, cust_ok);
        if (!cust_ok) nm[0] |= 1u;                     // rule 8 (:220-221)
        mk[0] = nm[0]; mk[1] = nm[1]; mk[2] = nm[2]; mk[3] = nm[3];
        now = c_node;
        if ((mk[0] | mk[1] | mk[2] | mk[3]) == 0u) {
          if (lane == 0) atomicOr(status, EVRP_ST_NO_FEASIBLE);
          done = true;
        } else if (t >= Tcap) {
          if (lane == 0) atomicOr(status, EVRP_ST_STEP_CAP);
          done = true;
        } else {
          // next step's glimpse keys: start pulling them into L2 now (the FF_tour GEMMs cover the latency)
#pragma unroll
          for (int jb = 0; jb < 4; ++jb) {
            if ((mk[jb] >> lane) & 1u) {
              const float* r = kvlb + (size_t)(jb * 32 + lane) * 384;
#pragma unroll
              for (int c = 0; c < 4; ++c) prefetch_l2(r + c * 32);
            }
          }
        }
      }
      if (done && lane == 0) { steps_used[row] = t; atomicMax(t_max, t); }
      __syncwarp();
    }
  }
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
    for (int t = 0; t < Tcap; ++t) s += rw[t];          // cost = R.sum(1) (:301), sequential fp32 like ATen's row sum for short rows
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

}  // namespace evrp

extern "C" {

int evrp_rollout(const evrp_phys* phys, const evrp_decoder_weights* w, const evrp_rollout_cfg* cfg,
                 const float* emb, const float* kvl, const float* fixed_ctx, const float* dynamic0,
                 const float* distances, const float* slope, int B, int S, int F, const int64_t* forced_actions,
                 const float* uniforms, int64_t* pi, float* ll, float* reward, float* energy, uint32_t* saved_mask,
                 float* saved_veh, int32_t* steps_used, int32_t* t_max, int32_t* status, void* stream) {
  using namespace evrp;
  if (!phys || !w || !cfg || !emb || !kvl || !fixed_ctx || !dynamic0 || !distances || !slope || !pi || !ll ||
      !reward || !steps_used || !t_max || !status || B < 0 || S < 2 || F < 1 || F >= S || cfg->max_steps < 1 ||
      cfg->n_replicas < 1 || !(cfg->temp > 0.f))
    return -EVRP_E_BADARG;
  if (S > EVRP_MAX_S) return -EVRP_E_TOO_MANY_NODES;
  if (B == 0) return 0;
  const int rows = B * cfg->n_replicas;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(DecodeSmem);
  EVRP_CUDA_OK(cudaFuncSetAttribute(rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (rows + G - 1) / G;
  rollout_kernel<<<grid, NT, smem, st>>>(*phys, *w, *cfg, emb, kvl, fixed_ctx, dynamic0, distances, slope, B, S, F,
                                         rows, forced_actions, uniforms, pi, ll, reward, energy, saved_mask,
                                         saved_veh, steps_used, t_max, status);
  EVRP_LAUNCH_OK();
  return 0;
}

int evrp_sample_min(const float* reward, const int64_t* pi, int B, int R, int Tcap, float* mincost, int32_t* argmin,
                    int64_t* minpi, void* stream) {
  if (!reward || !pi || !mincost || !argmin || !minpi || B < 0 || R < 1 || Tcap < 1) return -EVRP_E_BADARG;
  if (B == 0) return 0;
  evrp::sample_min_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(reward, pi, B, R, Tcap, mincost, argmin, minpi);
  EVRP_LAUNCH_OK();
  return 0;
}

}  // extern "C"
