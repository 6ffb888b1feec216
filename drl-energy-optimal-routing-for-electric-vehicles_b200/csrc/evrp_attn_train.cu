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

__device__ __forceinline__ void load_row16(const float* __restrict__ p, float (&r)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p) + c);
    r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
  }
}
__device__ __forceinline__ void lds_row16(const float* p, float (&r)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 v = reinterpret_cast<const float4*>(p)[c];
    r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
  }
}
__device__ __forceinline__ void store_row16(float* p, const float (&r)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) reinterpret_cast<float4*>(p)[c] = make_float4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
}
__device__ __forceinline__ float dot16(const float (&a)[16], const float (&b)[16]) {
  float s0 = a[0] * b[0], s1 = a[1] * b[1], s2 = a[2] * b[2], s3 = a[3] * b[3];      // 4 chains: latency / 4
#pragma unroll
  for (int d = 4; d < 16; d += 4) {
    s0 = fmaf(a[d], b[d], s0); s1 = fmaf(a[d + 1], b[d + 1], s1); s2 = fmaf(a[d + 2], b[d + 2], s2); s3 = fmaf(a[d + 3], b[d + 3], s3);
  }
  return (s0 + s1) + (s2 + s3);
}

// grid (B, 8 / HW), block HW * 32, dynamic shared memory HW * 2 * S * 16 floats
__global__ void __launch_bounds__(NTH)
attn_train_fwd_kernel(const float* __restrict__ qkv, int S, float* __restrict__ heads, float* __restrict__ lse) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, hw = (threadIdx.x >> 5) / WPH, part = (threadIdx.x >> 5) % WPH, lane = threadIdx.x & 31;
  const int h = blockIdx.y * HW + hw, row0 = part * 32 + lane, rstep = WPH * 32;
  float* Ks = sm + (size_t)hw * 2 * S * 16;
  float* Vs = Ks + (size_t)S * 16;
  const float* base = qkv + (size_t)b * S * 384 + h * 16;
  for (int j = row0; j < S; j += rstep) {
    float r[16];
    load_row16(base + (size_t)j * 384 + 128, r); store_row16(Ks + j * 16, r);
    load_row16(base + (size_t)j * 384 + 256, r); store_row16(Vs + j * 16, r);
  }
  __syncthreads();
  for (int i = row0; i < S; i += rstep) {
    float q[16], acc[16];
    load_row16(base + (size_t)i * 384, q);
#pragma unroll
    for (int d = 0; d < 16; ++d) { q[d] *= 0.25f; acc[d] = 0.f; }          // norm_factor 1/sqrt(16), exact scaling
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < S; ++j) {
      float k[16], v[16];
      lds_row16(Ks + j * 16, k); lds_row16(Vs + j * 16, v);
      const float s = dot16(q, k);
      const float mn = fmaxf(m, s);
      const float sc = expf(m - mn), e = expf(s - mn);
#pragma unroll
      for (int d = 0; d < 16; ++d) acc[d] = fmaf(acc[d], sc, e * v[d]);
      l = fmaf(l, sc, e); m = mn;
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < 16; ++d) acc[d] *= inv;
    store_row16(heads + ((size_t)b * S + i) * 128 + h * 16, acc);
    lse[((size_t)b * 8 + h) * S + i] = m + logf(l);
  }
}

// grid (B, 8 / HW), block HW * 32, dynamic shared memory HW * (4 * S * 16 + 2 * roundup4(S)) floats
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
    float r[16], g[16], o[16];
    load_row16(base + (size_t)i * 384, r);
#pragma unroll
    for (int d = 0; d < 16; ++d) r[d] *= 0.25f;            // scaled queries: s_ij = Qs_i . K_j
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
    float q[16], g[16], dq[16];
    lds_row16(Qs + i * 16, q); lds_row16(Gs + i * 16, g);
#pragma unroll
    for (int d = 0; d < 16; ++d) dq[d] = 0.f;
    const float Li = Ls[i], Di = Ds[i];
    for (int j = 0; j < S; ++j) {
      float k[16], v[16];
      lds_row16(Ks + j * 16, k); lds_row16(Vs + j * 16, v);
      const float p = expf(dot16(q, k) - Li);
      const float ds = p * (dot16(g, v) - Di);
#pragma unroll
      for (int d = 0; d < 16; ++d) dq[d] = fmaf(ds, k[d], dq[d]);
    }
#pragma unroll
    for (int d = 0; d < 16; ++d) dq[d] *= 0.25f;           // d s / d Q = K / 4
    store_row16(gbase + (size_t)i * 384, dq);
  }
  // ---- dK, dV: lanes own key rows ----
  for (int j = row0; j < S; j += rstep) {
    float k[16], v[16], dk[16], dv[16];
    lds_row16(Ks + j * 16, k); lds_row16(Vs + j * 16, v);
#pragma unroll
    for (int d = 0; d < 16; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    for (int i = 0; i < S; ++i) {
      float q[16], g[16];
      lds_row16(Qs + i * 16, q); lds_row16(Gs + i * 16, g);
      const float p = expf(dot16(q, k) - Ls[i]);
      const float ds = p * (dot16(g, v) - Ds[i]);
#pragma unroll
      for (int d = 0; d < 16; ++d) { dv[d] = fmaf(p, g[d], dv[d]); dk[d] = fmaf(ds, q[d], dk[d]); }   // q already carries the 1/4
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
