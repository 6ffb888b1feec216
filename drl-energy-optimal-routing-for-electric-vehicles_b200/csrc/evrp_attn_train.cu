// Encoder self-attention for the TRAINING forward / backward (SURVEY.md §8 a2 + a10; nets/GraphEncoder.py:81-102 and
// its autograd backward inside trainer.py:124-125 `loss.backward()`):
//     heads[b, i, h*16 : h*16+16] = softmax_j( Q_h[i] . K_h[j] / 4 ) @ V_h          qkv [B,S,384] = Q | K | V, head-major
// fp32 SIMT, flash-style: nothing of size [B,H,S,S] touches memory (the PyTorch formulation runs 6 batched GEMMs with
// K = 16 plus softmax / permute / copy kernels per layer and direction).  The forward also returns the row
// log-sum-exp L[b,h,i], from which the backward recomputes the probabilities:
//     P_ij = exp(s_ij - L_i),  D_i = dO_i . O_i,  dS_ij = P_ij (dO_i . V_j - D_i) / 4
//     dQ_i = sum_j dS_ij K_j     (lanes own query rows; K_j, V_j broadcast from shared memory)
//     dK_j = sum_i dS_ij Q_i,  dV_j = sum_i P_ij dO_i     (lanes own key rows; Q_i, dO_i broadcast from shared memory)
// so no cross-lane reduction and no atomics are needed.  4 heads per CTA, WPH warps per head sharing the head's tiles
// (each takes every WPH-th block of 32 rows).
#include <math.h>
#include "evrp_common.cuh"

namespace evrp {
namespace atr {

constexpr int HW = 4;                         // heads per CTA
constexpr int WPH = 2;                        // warps per head
constexpr int NTH = HW * WPH * 32;

// Every FMA below is a packed FFMA2 (PTX fma.rn.f32x2, evrp_common.cuh::ffma2): the (d, d+1) pairs of a 16-float head row
// sit in adjacent registers exactly as the 16-byte loads deliver them; each half is an ordinary IEEE fp32 FMA.
__device__ __forceinline__ void load_row16(const float* __restrict__ p, float2 (&r)[8]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p) + c);
    r[2 * c] = make_float2(v.x, v.y); r[2 * c + 1] = make_float2(v.z, v.w);
  }
}
__device__ __forceinline__ void lds_row16(const float* p, float2 (&r)[8]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 v = reinterpret_cast<const float4*>(p)[c];
    r[2 * c] = make_float2(v.x, v.y); r[2 * c + 1] = make_float2(v.z, v.w);
  }
}
__device__ __forceinline__ void store_row16(float* p, const float2 (&r)[8]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) reinterpret_cast<float4*>(p)[c] = make_float4(r[2 * c].x, r[2 * c].y, r[2 * c + 1].x, r[2 * c + 1].y);
}
__device__ __forceinline__ float dot16(const float2 (&a)[8], const float2 (&b)[8]) {
  float2 s0 = make_float2(0.f, 0.f), s1 = s0;                  // 2 packed chains of 4
#pragma unroll
  for (int c = 0; c < 8; c += 2) { ffma2(s0, a[c], b[c]); ffma2(s1, a[c + 1], b[c + 1]); }
  return (s0.x + s1.x) + (s0.y + s1.y);
}
__device__ __forceinline__ void scale16(float2 (&a)[8], float s) {
#pragma unroll
  for (int c = 0; c < 8; ++c) { a[c].x *= s; a[c].y *= s; }
}

// grid (B, 8 / HW), block HW * WPH * 32, dynamic shared memory HW * 2 * S * 16 floats.  Two query rows per lane (rows
// part * 32 + lane and + 64) share every K / V row read from shared memory.
__global__ void __launch_bounds__(NTH)
attn_train_fwd_kernel(const float* __restrict__ qkv, int S, float* __restrict__ heads, float* __restrict__ lse) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, hw = (threadIdx.x >> 5) / WPH, part = (threadIdx.x >> 5) % WPH, lane = threadIdx.x & 31;
  const int h = blockIdx.y * HW + hw, row0 = part * 32 + lane, rstep = WPH * 32;
  float* Ks = sm + (size_t)hw * 2 * S * 16;
  float* Vs = Ks + (size_t)S * 16;
  const float* base = qkv + (size_t)b * S * 384 + h * 16;
  for (int j = row0; j < S; j += rstep) {
    float2 r[8];
    load_row16(base + (size_t)j * 384 + 128, r); store_row16(Ks + j * 16, r);
    load_row16(base + (size_t)j * 384 + 256, r); store_row16(Vs + j * 16, r);
  }
  __syncthreads();
  for (int i0 = row0; i0 < S; i0 += 2 * rstep) {             // (one pass for S <= 128)
    const int ia = i0, ib = i0 + rstep;
    const bool vb = ib < S;
    float2 qa[8], qb[8], oa[8], ob[8];
    load_row16(base + (size_t)ia * 384, qa);
    load_row16(base + (size_t)(vb ? ib : ia) * 384, qb);
    scale16(qa, 0.25f); scale16(qb, 0.25f);                    // norm_factor 1/sqrt(16), exact scaling
#pragma unroll
    for (int c = 0; c < 8; ++c) { oa[c] = make_float2(0.f, 0.f); ob[c] = oa[c]; }
    float ma = -INFINITY, mb = -INFINITY, la = 0.f, lb = 0.f;
    for (int j = 0; j < S; ++j) {
      float2 k[8], v[8];
      lds_row16(Ks + j * 16, k); lds_row16(Vs + j * 16, v);
      const float sa = dot16(qa, k), sb = dot16(qb, k);
      if (sa > ma) {                                           // online softmax: rescale on a new maximum
        const float r = expf(ma - sa);                         // exp(-inf) = 0 on the first hit
        la *= r; scale16(oa, r); ma = sa;
      }
      if (sb > mb) { const float r = expf(mb - sb); lb *= r; scale16(ob, r); mb = sb; }
      const float pa = expf(sa - ma), pb = expf(sb - mb);
      la += pa; lb += pb;
      const float2 pa2 = make_float2(pa, pa), pb2 = make_float2(pb, pb);
#pragma unroll
      for (int c = 0; c < 8; ++c) { ffma2(oa[c], pa2, v[c]); ffma2(ob[c], pb2, v[c]); }
    }
    scale16(oa, 1.f / la);
    store_row16(heads + ((size_t)b * S + ia) * 128 + h * 16, oa);
    lse[((size_t)b * 8 + h) * S + ia] = ma + logf(la);
    if (vb) {
      scale16(ob, 1.f / lb);
      store_row16(heads + ((size_t)b * S + ib) * 128 + h * 16, ob);
      lse[((size_t)b * 8 + h) * S + ib] = mb + logf(lb);
    }
  }
}

// grid (B, 8 / HW), block HW * WPH * 32, dynamic shared memory HW * (4 * S * 16 + 2 * roundup4(S)) floats
__global__ void __launch_bounds__(NTH)
attn_train_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ heads, const float* __restrict__ lse,
                      const float* __restrict__ grad_heads, int S, float* __restrict__ grad_qkv) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, hw = (threadIdx.x >> 5) / WPH, part = (threadIdx.x >> 5) % WPH, lane = threadIdx.x & 31;
  const int h = blockIdx.y * HW + hw, row0 = part * 32 + lane, rstep = WPH * 32;
  float* Qs = sm + (size_t)hw * (4 * S * 16 + 2 * ((S + 3) & ~3));     // 16-byte aligned per-warp slices
  float* Ks = Qs + (size_t)S * 16;
  float* Vs = Ks + (size_t)S * 16;
  float* Gs = Vs + (size_t)S * 16;                         // dO rows
  float* Ls = Gs + (size_t)S * 16;                         // row log-sum-exp
  float* Ds = Ls + S;                                      // D_i = dO_i . O_i
  const float* base = qkv + (size_t)b * S * 384 + h * 16;
  float* gbase = grad_qkv + (size_t)b * S * 384 + h * 16;
  for (int i = row0; i < S; i += rstep) {
    float2 r[8], g[8], o[8];
    load_row16(base + (size_t)i * 384, r);
    scale16(r, 0.25f);                                     // scaled queries: s_ij = Qs_i . K_j
    store_row16(Qs + i * 16, r);
    load_row16(base + (size_t)i * 384 + 128, r); store_row16(Ks + i * 16, r);
    load_row16(base + (size_t)i * 384 + 256, r); store_row16(Vs + i * 16, r);
    load_row16(grad_heads + ((size_t)b * S + i) * 128 + h * 16, g); store_row16(Gs + i * 16, g);
    load_row16(heads + ((size_t)b * S + i) * 128 + h * 16, o);
    Ls[i] = lse[((size_t)b * 8 + h) * S + i];
    Ds[i] = dot16(g, o);
  }
  __syncthreads();
  // ---- dQ: lanes own query rows ----
  for (int i = row0; i < S; i += rstep) {
    float2 q[8], g[8], dq[8];
    lds_row16(Qs + i * 16, q); lds_row16(Gs + i * 16, g);
#pragma unroll
    for (int c = 0; c < 8; ++c) dq[c] = make_float2(0.f, 0.f);
    const float Li = Ls[i], Di = Ds[i];
#pragma unroll 2
    for (int j = 0; j < S; ++j) {
      float2 k[8], v[8];
      lds_row16(Ks + j * 16, k); lds_row16(Vs + j * 16, v);
      const float p = expf(dot16(q, k) - Li);
      const float ds = p * (dot16(g, v) - Di);
      const float2 ds2 = make_float2(ds, ds);
#pragma unroll
      for (int c = 0; c < 8; ++c) ffma2(dq[c], ds2, k[c]);
    }
    scale16(dq, 0.25f);                                    // d s / d Q = K / 4
    store_row16(gbase + (size_t)i * 384, dq);
  }
  // ---- dK, dV: lanes own key rows ----
  for (int j = row0; j < S; j += rstep) {
    float2 k[8], v[8], dk[8], dv[8];
    lds_row16(Ks + j * 16, k); lds_row16(Vs + j * 16, v);
#pragma unroll
    for (int c = 0; c < 8; ++c) { dk[c] = make_float2(0.f, 0.f); dv[c] = dk[c]; }
#pragma unroll 2
    for (int i = 0; i < S; ++i) {
      float2 q[8], g[8];
      lds_row16(Qs + i * 16, q); lds_row16(Gs + i * 16, g);
      const float p = expf(dot16(q, k) - Ls[i]);
      const float ds = p * (dot16(g, v) - Ds[i]);
      const float2 p2 = make_float2(p, p), ds2 = make_float2(ds, ds);
#pragma unroll
      for (int c = 0; c < 8; ++c) { ffma2(dv[c], p2, g[c]); ffma2(dk[c], ds2, q[c]); }   // q already carries the 1/4
    }
    store_row16(gbase + (size_t)j * 384 + 128, dk);
    store_row16(gbase + (size_t)j * 384 + 256, dv);
  }
}

}  // namespace atr
}  // namespace evrp

extern "C" {

int evrp_self_attention_forward(const float* qkv, int B, int S, float* heads, float* lse, void* stream) {
  using namespace evrp::atr;
  if (!qkv || !heads || !lse || B < 0 || S < 1) return -EVRP_E_BADARG;
  if (S > EVRP_MAX_S) return -EVRP_E_TOO_MANY_NODES;
  if (B == 0) return 0;
  const size_t smem = (size_t)HW * 2 * S * 16 * sizeof(float);
  EVRP_CUDA_OK(cudaFuncSetAttribute(attn_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_train_fwd_kernel<<<dim3(B, 8 / HW), NTH, smem, (cudaStream_t)stream>>>(qkv, S, heads, lse);
  EVRP_LAUNCH_OK();
  return 0;
}

int evrp_self_attention_backward(const float* qkv, const float* heads, const float* lse, const float* grad_heads, int B,
                                 int S, float* grad_qkv, void* stream) {
  using namespace evrp::atr;
  if (!qkv || !heads || !lse || !grad_heads || !grad_qkv || B < 0 || S < 1) return -EVRP_E_BADARG;
  if (S > EVRP_MAX_S) return -EVRP_E_TOO_MANY_NODES;
  if (B == 0) return 0;
  const size_t smem = (size_t)HW * (4 * S * 16 + 2 * ((S + 3) & ~3)) * sizeof(float);
  EVRP_CUDA_OK(cudaFuncSetAttribute(attn_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_train_bwd_kernel<<<dim3(B, 8 / HW), NTH, smem, (cudaStream_t)stream>>>(qkv, heads, lse, grad_heads, S, grad_qkv);
  EVRP_LAUNCH_OK();
  return 0;
}

}  // extern "C"
