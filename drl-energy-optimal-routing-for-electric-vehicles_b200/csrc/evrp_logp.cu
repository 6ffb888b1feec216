// Teacher-forced decoder log-probabilities and their backward (the differentiable half of
// nets/DRLModel.py:379-452 for the REINFORCE step of trainer.py:115-127).
//
// Given the recorded rollout of a batch -- actions pi[b,t], feasibility masks (bit words), step counts -- and the
// step queries q[b,t,:] = fixed_ctx + route context + emb[now] (built by the caller, so autograd carries their
// gradient on into FF_tour / the encoder), this computes for every live step
//     compat_h[j] = q_h . K_h[j] / 4 (feasible j only),  a_h = softmax(compat_h),  heads_h = sum_j a_h[j] V_h[j]
//     logits[j]   = heads . lK[j] / sqrt(128)            (project_out is folded into lK, as in the rollout kernels)
//     ll          = log_softmax(clip * tanh(logits) / temp)[pi]
// and, backward, d ll -> (d q[b,t,:], d kvl[b,j,:]) in ONE pass structure without materialising any [B,H,T,S] tensor
// (the PyTorch formulation moves ~20 of them through HBM per training step).
//
// One CTA per instance; its warps stride over the instance's decode steps; a lane owns 4 of the 128 channels (head =
// lane / 4), exactly the layout of the rollout kernels' row phase.  The softmaxes are evaluated online (running max /
// sum), so the forward keeps no per-node scratch.  The backward recomputes the forward quantities of its step and adds
// the step's contribution to the instance's d kvl [S,384] with vectorised fire-and-forget reductions
// (red.global.add.v4.f32, resolved in L2; shared-memory fp32 atomics compile to a CAS spin loop on sm_100 and measured
// 5.1 ms).
#include <math.h>
#include "evrp_common.cuh"

namespace evrp {
namespace lp {

constexpr int NW_FWD = 8;        // warps per CTA, forward
constexpr int NW_BWD = 8;        // warps per CTA, backward
constexpr int CH = 4;            // node records in flight per warp and pass
constexpr int SAVED = 196;       // floats of saved statistics per (instance, step): heads 128 | (m, l) x 32 lanes | lse | pad

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}
__device__ __forceinline__ float head_sum(float v) {          // over the 4 lanes of a head
  v += __shfl_xor_sync(FULL, v, 1);
  v += __shfl_xor_sync(FULL, v, 2);
  return v;
}
__device__ __forceinline__ float warp_sum32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

struct StepFwd {
  float m, l;          // running max / sum of this lane's head
  float4 heads;        // glimpse head outputs (this lane's 4 channels), normalised
  float lse;           // log-sum-exp of the clipped, tempered logits over the feasible nodes
  float z_sel;         // clipped, tempered logit of the chosen node
};

// forward of one (instance, step) by one warp.  list[0..nf) = feasible node indices (shared memory, this warp's)
__device__ __forceinline__ StepFwd step_forward(const float* __restrict__ kv, const unsigned char* list, int nf, int lane,
                                                const float4& q4, int sel, float clip, float inv_temp) {
  StepFwd r;
  float m = -INFINITY, l = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i0 = 0; i0 < nf; i0 += CH) {
    float4 k4[CH], v4[CH];
#pragma unroll
    for (int a = 0; a < CH; ++a) {
      const float* p = kv + (size_t)list[min(i0 + a, nf - 1)] * 384 + lane * 4;
      k4[a] = ldg4(p); v4[a] = ldg4(p + 128);
    }
#pragma unroll
    for (int a = 0; a < CH; ++a) {
      if (i0 + a < nf) {                                       // warp-uniform
        const float c = head_sum(dot4(q4, k4[a])) * 0.25f;
        const float mn = fmaxf(m, c);
        const float sc = expf(m - mn), e = expf(c - mn);       // first node: exp(-inf) = 0
        acc.x = fmaf(acc.x, sc, e * v4[a].x); acc.y = fmaf(acc.y, sc, e * v4[a].y);
        acc.z = fmaf(acc.z, sc, e * v4[a].z); acc.w = fmaf(acc.w, sc, e * v4[a].w);
        l = fmaf(l, sc, e); m = mn;
      }
    }
  }
  const float inv_l = 1.f / l;
  r.m = m; r.l = l;
  r.heads = make_float4(acc.x * inv_l, acc.y * inv_l, acc.z * inv_l, acc.w * inv_l);
  const float rs = rsqrtf((float)D);
  float M = -INFINITY, L = 0.f, z_sel = -INFINITY;
  for (int i0 = 0; i0 < nf; i0 += CH) {
    float4 k4[CH];
#pragma unroll
    for (int a = 0; a < CH; ++a) k4[a] = ldg4(kv + (size_t)list[min(i0 + a, nf - 1)] * 384 + 256 + lane * 4);
#pragma unroll
    for (int a = 0; a < CH; ++a) {
      if (i0 + a < nf) {
        const float lg = warp_sum32(dot4(r.heads, k4[a])) * rs;
        const float z = (clip > 0.f ? clip * tanhf(lg) : lg) * inv_temp;
        const float Mn = fmaxf(M, z);
        L = fmaf(L, expf(M - Mn), expf(z - Mn)); M = Mn;
        if ((int)list[i0 + a] == sel) z_sel = z;
      }
    }
  }
  r.lse = M + logf(L);
  r.z_sel = z_sel;
  return r;
}

__device__ __forceinline__ int build_list(const uint32_t* __restrict__ mk, unsigned char* list, int lane) {
  int nf = 0;
#pragma unroll
  for (int jb = 0; jb < 4; ++jb) {
    const uint32_t bits = mk[jb];
    if ((bits >> lane) & 1u) list[nf + __popc(bits & ((1u << lane) - 1u))] = (unsigned char)(jb * 32 + lane);
    nf += __popc(bits);
  }
  __syncwarp();
  return nf;
}

__global__ void __launch_bounds__(NW_FWD * 32)
pointer_logp_fwd_kernel(const float* __restrict__ q, const float* __restrict__ kvl, const uint32_t* __restrict__ mask_bits,
                        const int64_t* __restrict__ pi, const int32_t* __restrict__ steps, int T, int S, float clip,
                        float inv_temp, float* __restrict__ ll, float* __restrict__ saved /*[B*T,193] or null*/) {
  __shared__ unsigned char lists[NW_FWD][EVRP_MAX_S];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* kv = kvl + (size_t)b * S * 384;
  const int nsteps = min(steps[b], T);
  for (int t = warp; t < T; t += NW_FWD) {
    const size_t bt = (size_t)b * T + t;
    if (t >= nsteps) { if (lane == 0) ll[bt] = 0.f; continue; }      // padded step (finished instance): log-prob 0
    uint32_t mk[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) mk[w] = mask_bits[bt * 4 + w];
    __syncwarp();
    const int nf = build_list(mk, lists[warp], lane);
    if (nf == 0) { if (lane == 0) ll[bt] = 0.f; continue; }
    const float4 q4 = ldg4(q + bt * D + lane * 4);
    const StepFwd r = step_forward(kv, lists[warp], nf, lane, q4, (int)pi[bt], clip, inv_temp);
    if (lane == 0) ll[bt] = r.z_sel - r.lse;
    if (saved) {                                   // per-step statistics for the backward: heads | (m, l) per lane | lse
      float* sv = saved + bt * SAVED;
      *reinterpret_cast<float4*>(sv + lane * 4) = r.heads;
      *reinterpret_cast<float2*>(sv + 128 + lane * 2) = make_float2(r.m, r.l);
      if (lane == 0) sv[192] = r.lse;
    }
  }
}

__device__ __forceinline__ void red4(float* p, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

__global__ void __launch_bounds__(NW_BWD * 32)
pointer_logp_bwd_kernel(const float* __restrict__ q, const float* __restrict__ kvl, const uint32_t* __restrict__ mask_bits,
                        const int64_t* __restrict__ pi, const int32_t* __restrict__ steps, const float* __restrict__ grad_ll,
                        int T, int S, float clip, float inv_temp, float* __restrict__ grad_q, float* __restrict__ grad_kvl,
                        const float* __restrict__ saved /*forward statistics or null (recompute)*/) {
  __shared__ unsigned char lists[NW_BWD][EVRP_MAX_S];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* kv = kvl + (size_t)b * S * 384;
  float* gacc = grad_kvl + (size_t)b * S * 384;                       // zero-filled by the launcher
  const int nsteps = min(steps[b], T);
  const float rs = rsqrtf((float)D);
  for (int t = warp; t < T; t += NW_BWD) {
    const size_t bt = (size_t)b * T + t;
    float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
    const float g = (t < nsteps) ? grad_ll[bt] : 0.f;
    if (t < nsteps && g != 0.f) {
      uint32_t mk[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) mk[w] = mask_bits[bt * 4 + w];
      __syncwarp();
      unsigned char* list = lists[warp];
      const int nf = build_list(mk, list, lane);
      if (nf > 0) {
        const float4 q4 = ldg4(q + bt * D + lane * 4);
        const int sel = (int)pi[bt];
        StepFwd r;
        if (saved) {                               // the forward kept heads / m / l / lse of this step: no recompute passes
          const float* sv = saved + bt * SAVED;
          r.heads = ldg4(sv + lane * 4);
          const float2 ml = __ldg(reinterpret_cast<const float2*>(sv + 128 + lane * 2));
          r.m = ml.x; r.l = ml.y; r.lse = __ldg(sv + 192); r.z_sel = 0.f;
        } else {
          r = step_forward(kv, list, nf, lane, q4, sel, clip, inv_temp);
        }
        // ---- logits backward: d z_j = g (1[j = sel] - p_j);  through temp, tanh clip, 1/sqrt(d) ----
        float4 dh = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i0 = 0; i0 < nf; i0 += CH) {
          float4 k4[CH];
#pragma unroll
          for (int a = 0; a < CH; ++a) k4[a] = ldg4(kv + (size_t)list[min(i0 + a, nf - 1)] * 384 + 256 + lane * 4);
#pragma unroll
          for (int a = 0; a < CH; ++a) {
            if (i0 + a < nf) {
              const int j = list[i0 + a];
              const float lg = warp_sum32(dot4(r.heads, k4[a])) * rs;
              const float th = tanhf(lg);
              const float z = (clip > 0.f ? clip * th : lg) * inv_temp;
              const float p = expf(z - r.lse);
              const float dz = g * ((j == sel ? 1.f : 0.f) - p);
              const float dlg = dz * inv_temp * (clip > 0.f ? clip * (1.f - th * th) : 1.f) * rs;
              dh.x = fmaf(dlg, k4[a].x, dh.x); dh.y = fmaf(dlg, k4[a].y, dh.y);
              dh.z = fmaf(dlg, k4[a].z, dh.z); dh.w = fmaf(dlg, k4[a].w, dh.w);
              red4(gacc + (size_t)j * 384 + 256 + lane * 4, dlg * r.heads.x, dlg * r.heads.y, dlg * r.heads.z, dlg * r.heads.w);
            }
          }
        }
        // ---- glimpse backward: softmax Jacobian with D_h = sum_j a_j (dh . V_j) = dh . heads_h ----
        const float Dh = head_sum(dot4(dh, r.heads));
        const float inv_l = 1.f / r.l;
        for (int i0 = 0; i0 < nf; i0 += CH) {
          float4 k4[CH], v4[CH];
#pragma unroll
          for (int a = 0; a < CH; ++a) {
            const float* p = kv + (size_t)list[min(i0 + a, nf - 1)] * 384 + lane * 4;
            k4[a] = ldg4(p); v4[a] = ldg4(p + 128);
          }
#pragma unroll
          for (int a = 0; a < CH; ++a) {
            if (i0 + a < nf) {
              const int j = list[i0 + a];
              const float c = head_sum(dot4(q4, k4[a])) * 0.25f;
              const float at = expf(c - r.m) * inv_l;                    // attention weight of this lane's head
              const float da = head_sum(dot4(dh, v4[a]));
              const float dc = at * (da - Dh) * 0.25f;
              dq.x = fmaf(dc, k4[a].x, dq.x); dq.y = fmaf(dc, k4[a].y, dq.y);
              dq.z = fmaf(dc, k4[a].z, dq.z); dq.w = fmaf(dc, k4[a].w, dq.w);
              float* gj = gacc + (size_t)j * 384 + lane * 4;
              red4(gj, dc * q4.x, dc * q4.y, dc * q4.z, dc * q4.w);
              red4(gj + 128, at * dh.x, at * dh.y, at * dh.z, at * dh.w);
            }
          }
        }
      }
    }
    *reinterpret_cast<float4*>(grad_q + bt * D + lane * 4) = dq;
  }
}

}  // namespace lp
}  // namespace evrp

extern "C" {

int evrp_pointer_logp_saved_floats(void) { return evrp::lp::SAVED; }

int evrp_pointer_logp_forward(const float* q, const float* kvl, const uint32_t* mask_bits, const int64_t* pi,
                              const int32_t* steps, int B, int T, int S, float tanh_clipping, float temp, float* ll,
                              float* saved_stats, void* stream) {
  if (!q || !kvl || !mask_bits || !pi || !steps || !ll || B < 0 || T < 1 || S < 2 || !(temp > 0.f)) return -EVRP_E_BADARG;
  if (S > EVRP_MAX_S) return -EVRP_E_TOO_MANY_NODES;
  if (B == 0) return 0;
  evrp::lp::pointer_logp_fwd_kernel<<<B, evrp::lp::NW_FWD * 32, 0, (cudaStream_t)stream>>>(
      q, kvl, mask_bits, pi, steps, T, S, tanh_clipping, 1.f / temp, ll, saved_stats);
  EVRP_LAUNCH_OK();
  return 0;
}

int evrp_pointer_logp_backward(const float* q, const float* kvl, const uint32_t* mask_bits, const int64_t* pi,
                               const int32_t* steps, const float* grad_ll, int B, int T, int S, float tanh_clipping,
                               float temp, float* grad_q, float* grad_kvl, const float* saved_stats, void* stream) {
  if (!q || !kvl || !mask_bits || !pi || !steps || !grad_ll || !grad_q || !grad_kvl || B < 0 || T < 1 || S < 2 ||
      !(temp > 0.f))
    return -EVRP_E_BADARG;
  if (S > EVRP_MAX_S) return -EVRP_E_TOO_MANY_NODES;
  if (B == 0) return 0;
  EVRP_CUDA_OK(cudaMemsetAsync(grad_kvl, 0, (size_t)B * S * 384 * sizeof(float), (cudaStream_t)stream));
  evrp::lp::pointer_logp_bwd_kernel<<<B, evrp::lp::NW_BWD * 32, 0, (cudaStream_t)stream>>>(
      q, kvl, mask_bits, pi, steps, grad_ll, T, S, tanh_clipping, 1.f / temp, grad_q, grad_kvl, saved_stats);
  EVRP_LAUNCH_OK();
  return 0;
}

}  // extern "C"
