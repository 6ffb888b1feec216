// Encoder + decoder precompute:
//   AttentionModel._init_embed     nets/DRLModel.py:192-200
//   GraphAttentionEncoder.forward  nets/GraphEncoder.py:202-214  (3 x MHA + BN + FF + BN)
//   AttentionModel._precompute     nets/DRLModel.py:345-362
//
// fp32-faithful path: greedy tours of this policy flip under TF32-level noise, so the
// dense contractions run as fp32 FFMA GEMMs (128x128x8 tiles, 8x8 register micro-tiles) with the bias / ReLU /
// residual / folded-BatchNorm epilogues fused.  Rows are the flattened B*S node dimension.
#include <math.h>
#include "evrp_common.cuh"

namespace evrp {

// ------------------------------------------------------------------------------------------------
// init embedding: one warp per node row, one lane per 4 channels (512-byte coalesced store per row)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
init_embed_kernel(const float* __restrict__ xy, const float* __restrict__ dyn, int B, int S, int F,
                  const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
                  const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                  float* __restrict__ h) {
  const int lane = threadIdx.x & 31;
  const size_t rows = (size_t)B * S;
  const size_t warp0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * (size_t)blockDim.x) >> 5;
  // this lane's 4 channels of the three embedding layers (nets/DRLModel.py:78-80), kept in registers
  float wd[4][2], ws[4][2], wc[4][3], bd[4], bs[4], bc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane * 4 + i;
    wd[i][0] = w0[d * 2]; wd[i][1] = w0[d * 2 + 1]; bd[i] = b0[d];
    ws[i][0] = w1[d * 2]; ws[i][1] = w1[d * 2 + 1]; bs[i] = b1[d];
    wc[i][0] = w2[d * 3]; wc[i][1] = w2[d * 3 + 1]; wc[i][2] = w2[d * 3 + 2]; bc[i] = b2[d];
  }
  for (size_t row = warp0; row < rows; row += nwarps) {
    const int j = (int)(row % S);
    const size_t b = row / S;
    const float x = __fdiv_rn(xy[(b * 2 + 0) * S + j], 100.f);
    const float y = __fdiv_rn(xy[(b * 2 + 1) * S + j], 100.f);
    float v[4];
    if (j == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaf(wd[i][1], y, fmaf(wd[i][0], x, 0.f)) + bd[i];
    } else if (j <= F) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaf(ws[i][1], y, fmaf(ws[i][0], x, 0.f)) + bs[i];
    } else {
      const float dem = dyn[(b * 4 + 1) * S + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaf(wc[i][2], dem, fmaf(wc[i][1], y, fmaf(wc[i][0], x, 0.f))) + bc[i];
    }
    *reinterpret_cast<float4*>(h + row * D + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// init embedding on its own (training path) and its backward: the three tiny nn.Linear layers of
// nets/DRLModel.py:192-200 have K = 2 / 2 / 3 inputs, so their weight gradients are 10 column reductions of
// g [B*S,128] weighted by (x, y, 1 | x, y, 1 | x, y, demand, 1) per node class.  One CTA = 128 channels x a stride of
// rows, partial sums per CTA, second kernel sums the CTAs in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
init_embed_bwd_kernel(const float* __restrict__ xy, const float* __restrict__ dyn, const float* __restrict__ g, int B, int S,
                      int F, float* __restrict__ partial /*[grid][10][128]*/) {
  __shared__ float red[3][10][D];
  const int c = threadIdx.x & 127, rl = threadIdx.x >> 7;              // channel, row lane (4 rows in flight per CTA)
  float a[10] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const size_t rows = (size_t)B * S;
#pragma unroll 4
  for (size_t row = (size_t)blockIdx.x * 4 + rl; row < rows; row += (size_t)gridDim.x * 4) {
    const int j = (int)(row % S);
    const size_t b = row / S;
    const float gv = __ldg(g + row * D + c);
    const float x = __fdiv_rn(__ldg(xy + (b * 2 + 0) * S + j), 100.f), y = __fdiv_rn(__ldg(xy + (b * 2 + 1) * S + j), 100.f);
    if (j == 0) { a[0] = fmaf(gv, x, a[0]); a[1] = fmaf(gv, y, a[1]); a[2] += gv; }
    else if (j <= F) { a[3] = fmaf(gv, x, a[3]); a[4] = fmaf(gv, y, a[4]); a[5] += gv; }
    else {
      const float dem = __ldg(dyn + (b * 4 + 1) * S + j);
      a[6] = fmaf(gv, x, a[6]); a[7] = fmaf(gv, y, a[7]); a[8] = fmaf(gv, dem, a[8]); a[9] += gv;
    }
  }
  if (rl > 0) {
#pragma unroll
    for (int k = 0; k < 10; ++k) red[rl - 1][k][c] = a[k];
  }
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int k = 0; k < 10; ++k) partial[((size_t)blockIdx.x * 10 + k) * D + c] = ((a[k] + red[0][k][c]) + red[1][k][c]) + red[2][k][c];
  }
}

__global__ void init_embed_bwd_reduce_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ gw0, float* __restrict__ gb0,
                                             float* __restrict__ gw1, float* __restrict__ gb1, float* __restrict__ gw2,
                                             float* __restrict__ gb2) {
  const int k = blockIdx.x, c = threadIdx.x;           // 10 blocks x 128 threads
  float acc = 0.f;
  for (int i = 0; i < nblk; ++i) acc += partial[((size_t)i * 10 + k) * D + c];
  switch (k) {                                         // nn.Linear layouts: weight [128][in], bias [128]
    case 0: case 1: gw0[c * 2 + k] = acc; break;
    case 2: gb0[c] = acc; break;
    case 3: case 4: gw1[c * 2 + (k - 3)] = acc; break;
    case 5: gb1[c] = acc; break;
    case 6: case 7: case 8: gw2[c * 3 + (k - 6)] = acc; break;
    default: gb2[c] = acc; break;
  }
}

// ------------------------------------------------------------------------------------------------
// SGEMM  C[M,N] = epi(A[M,K] @ Bm[K,N]);  A row-major (lda), Bm k-major (ldb), N % 128 == 0, K % 8 == 0
// epilogue: v = acc (+bias[n]) ; relu ; (+resid[m,n]) ; (*scale[n] + shift[n])
// ------------------------------------------------------------------------------------------------
constexpr int EPI_BIAS = 1, EPI_RELU = 2, EPI_RESID = 4, EPI_AFFINE = 8;
constexpr int BM = 128, BN = 128, BK = 8;

template <int EPI>
__global__ void __launch_bounds__(256, 2)
sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, float* __restrict__ C,
             int ldc, int M, int N, int K, const float* __restrict__ bias, const float* __restrict__ resid,
             int ldr, const float* __restrict__ scale, const float* __restrict__ shift) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int a_row = tid >> 1, a_k = (tid & 1) * 4;       // A tile: 128 rows x 8 k  (one float4 per thread)
  const int b_k = tid >> 5, b_n = (tid & 31) * 4;        // B tile: 8 k x 128 n
  const int ty = tid >> 4, tx = tid & 15;                // 16x16 threads, each 8x8 outputs (4+4 split)
  const bool a_ok = (m0 + a_row) < M;
  const float* Ap = A + (size_t)(m0 + (a_ok ? a_row : 0)) * lda + a_k;
  const float* Bp = Bm + (size_t)b_k * ldb + n0 + b_n;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 av = a_ok ? *reinterpret_cast<const float4*>(Ap) : make_float4(0, 0, 0, 0);
  float4 bv = *reinterpret_cast<const float4*>(Bp);
  As[0][a_k + 0][a_row] = av.x; As[0][a_k + 1][a_row] = av.y; As[0][a_k + 2][a_row] = av.z; As[0][a_k + 3][a_row] = av.w;
  *reinterpret_cast<float4*>(&Bs[0][b_k][b_n]) = bv;
  __syncthreads();
  const int nk = K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      av = a_ok ? *reinterpret_cast<const float4*>(Ap + (size_t)(kt + 1) * BK) : make_float4(0, 0, 0, 0);
      bv = *reinterpret_cast<const float4*>(Bp + (size_t)(kt + 1) * BK * ldb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      const int nxt = cur ^ 1;
      As[nxt][a_k + 0][a_row] = av.x; As[nxt][a_k + 1][a_row] = av.y; As[nxt][a_k + 2][a_row] = av.z; As[nxt][a_k + 3][a_row] = av.w;
      *reinterpret_cast<float4*>(&Bs[nxt][b_k][b_n]) = bv;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      if (EPI & EPI_BIAS) {
        const float4 bb = *reinterpret_cast<const float4*>(bias + n);
        v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
      }
      if (EPI & EPI_RELU) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = fmaxf(v[q], 0.f);
      }
      if (EPI & EPI_RESID) {
        const float4 rr = *reinterpret_cast<const float4*>(resid + (size_t)m * ldr + n);
        v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
      }
      if (EPI & EPI_AFFINE) {
        const float4 sc = *reinterpret_cast<const float4*>(scale + n);
        const float4 sh = *reinterpret_cast<const float4*>(shift + n);
        v[0] = fmaf(v[0], sc.x, sh.x); v[1] = fmaf(v[1], sc.y, sh.y);
        v[2] = fmaf(v[2], sc.z, sh.z); v[3] = fmaf(v[3], sc.w, sh.w);
      }
      *reinterpret_cast<float4*>(C + (size_t)m * ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

template <int EPI>
static int launch_sgemm(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M, int N, int K,
                        const float* bias, const float* resid, int ldr, const float* scale, const float* shift,
                        cudaStream_t st) {
  dim3 grid(N / BN, (M + BM - 1) / BM);
  sgemm_kernel<EPI><<<grid, 256, 0, st>>>(A, lda, Bm, ldb, C, ldc, M, N, K, bias, resid, ldr, scale, shift);
  EVRP_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// self-attention over the S nodes of one instance: one CTA per instance, one warp per head.
// qkv [rows,384] (Q | K | V, head h at columns h*16..h*16+15 of each block) -> heads [rows,128]
// (nets/GraphEncoder.py:87-102: softmax(norm_factor * Q K^T) V, no mask)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
self_attention_kernel(const float* __restrict__ qkv, int S, float* __restrict__ heads) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Ks = smem + (size_t)h * 2 * S * DK;
  float* Vs = Ks + (size_t)S * DK;
  const float* base = qkv + (size_t)b * S * 384;
  for (int idx = lane; idx < S * 4; idx += 32) {
    const int row = idx >> 2, c4 = (idx & 3) * 4;
    *reinterpret_cast<float4*>(Ks + row * DK + c4) =
        *reinterpret_cast<const float4*>(base + (size_t)row * 384 + 128 + h * DK + c4);
    *reinterpret_cast<float4*>(Vs + row * DK + c4) =
        *reinterpret_cast<const float4*>(base + (size_t)row * 384 + 256 + h * DK + c4);
  }
  __syncwarp();
  for (int i0 = 0; i0 < S; i0 += 64) {          // two query rows per lane per pass
    const int ia = i0 + lane, ib = i0 + 32 + lane;
    const bool va = ia < S, vb = ib < S;
    // every FMA is an FFMA2 packed along the head dimension: (d, d+1) pairs of q, K and V sit in adjacent
    // registers exactly as the 16-byte loads deliver them, so no operand has to be duplicated
    float2 qa[DK / 2], qb[DK / 2], oa[DK / 2], ob[DK / 2];
#pragma unroll
    for (int c4 = 0; c4 < DK; c4 += 4) {
      const float4 t = va ? *reinterpret_cast<const float4*>(base + (size_t)ia * 384 + h * DK + c4) : make_float4(0, 0, 0, 0);
      const float4 u = vb ? *reinterpret_cast<const float4*>(base + (size_t)ib * 384 + h * DK + c4) : make_float4(0, 0, 0, 0);
      qa[c4 / 2] = make_float2(t.x, t.y); qa[c4 / 2 + 1] = make_float2(t.z, t.w);
      qb[c4 / 2] = make_float2(u.x, u.y); qb[c4 / 2 + 1] = make_float2(u.z, u.w);
    }
#pragma unroll
    for (int c = 0; c < DK / 2; ++c) { oa[c] = make_float2(0.f, 0.f); ob[c] = make_float2(0.f, 0.f); }
    float ma = -INFINITY, mb = -INFINITY, la = 0.f, lb = 0.f;
    for (int j = 0; j < S; ++j) {
      float2 kk[DK / 2], vv[DK / 2];
#pragma unroll
      for (int c4 = 0; c4 < DK; c4 += 4) {
        const float4 t = *reinterpret_cast<const float4*>(Ks + j * DK + c4);
        const float4 u = *reinterpret_cast<const float4*>(Vs + j * DK + c4);
        kk[c4 / 2] = make_float2(t.x, t.y); kk[c4 / 2 + 1] = make_float2(t.z, t.w);
        vv[c4 / 2] = make_float2(u.x, u.y); vv[c4 / 2 + 1] = make_float2(u.z, u.w);
      }
      // two partial sums per row keep the FMA chains short (4 deep)
      float2 sa0 = make_float2(0.f, 0.f), sa1 = sa0, sb0 = sa0, sb1 = sa0;
#pragma unroll
      for (int c = 0; c < DK / 2; c += 2) {
        ffma2(sa0, qa[c], kk[c]); ffma2(sa1, qa[c + 1], kk[c + 1]);
        ffma2(sb0, qb[c], kk[c]); ffma2(sb1, qb[c + 1], kk[c + 1]);
      }
      const float sa = ((sa0.x + sa1.x) + (sa0.y + sa1.y)) * 0.25f;    // norm_factor = 1/sqrt(16), exact
      const float sb = ((sb0.x + sb1.x) + (sb0.y + sb1.y)) * 0.25f;
      if (sa > ma) {                                          // online softmax: rescale on a new maximum
        const float r = expf(ma - sa);                        // exp(-inf) = 0 on the first hit
        la *= r;
#pragma unroll
        for (int c = 0; c < DK / 2; ++c) { oa[c].x *= r; oa[c].y *= r; }
        ma = sa;
      }
      if (sb > mb) {
        const float r = expf(mb - sb);
        lb *= r;
#pragma unroll
        for (int c = 0; c < DK / 2; ++c) { ob[c].x *= r; ob[c].y *= r; }
        mb = sb;
      }
      const float pa = expf(sa - ma), pb = expf(sb - mb);
      la += pa; lb += pb;
      const float2 pa2 = make_float2(pa, pa), pb2 = make_float2(pb, pb);
#pragma unroll
      for (int c = 0; c < DK / 2; ++c) { ffma2(oa[c], pa2, vv[c]); ffma2(ob[c], pb2, vv[c]); }
    }
    if (va) {
      const float inv = 1.f / la;
      float* o = heads + ((size_t)b * S + ia) * D + h * DK;
#pragma unroll
      for (int c = 0; c < DK / 2; c += 2)
        *reinterpret_cast<float4*>(o + 2 * c) = make_float4(oa[c].x * inv, oa[c].y * inv, oa[c + 1].x * inv, oa[c + 1].y * inv);
    }
    if (vb) {
      const float inv = 1.f / lb;
      float* o = heads + ((size_t)b * S + ib) * D + h * DK;
#pragma unroll
      for (int c = 0; c < DK / 2; c += 2)
        *reinterpret_cast<float4*>(o + 2 * c) = make_float4(ob[c].x * inv, ob[c].y * inv, ob[c + 1].x * inv, ob[c + 1].y * inv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fixed context: graph mean over S then project_fixed_context (nets/DRLModel.py:348-350)
// one CTA of 128 threads per instance
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
fixed_context_kernel(const float* __restrict__ emb, int S, const float* __restrict__ WfcT, float* __restrict__ out) {
  __shared__ float mean[D];
  const int b = blockIdx.x, d = threadIdx.x;
  const float* e = emb + (size_t)b * S * D;
  float s = 0.f;
  for (int j = 0; j < S; ++j) s += e[(size_t)j * D + d];
  mean[d] = __fdiv_rn(s, (float)S);
  __syncthreads();
  float acc = 0.f;
#pragma unroll 8
  for (int k = 0; k < D; ++k) acc = fmaf(mean[k], WfcT[k * D + d], acc);
  out[(size_t)b * D + d] = acc;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// evrp_attn_tc.cu
int launch_attn_tc(const float* qkv, int B, int S, float* heads, int sm_count, cudaStream_t st);

// evrp_encoder_fused.cu
int launch_encoder_fused(const evrp_encoder_weights* w, const float* xy, const float* dyn, int B, int S, int F, float* emb,
                         float* kvl, float* step_tab, int sm_count, cudaStream_t st);

// evrp_tc_gemm.cu
int tc_gemm(int epi, const float* A, int lda, const float* Wh, const float* Wl, float* C, int ldc, int M, int N, int K,
            const float* bias, const float* resid, int ldr, const float* scale, const float* shift, int sm_count,
            cudaStream_t st, const float* aux = nullptr, const float* auxw = nullptr);

}  // namespace evrp

extern "C" {

size_t evrp_encoder_workspace_bytes(int B, int S) {
  const size_t rows = (size_t)B * S;
  // h_a, h_b, heads: [rows,128]; wide: [rows,512] (holds qkv [rows,384] or the FF hidden)
  return evrp::align_up(rows * 128 * 4, 256) * 3 + evrp::align_up(rows * 512 * 4, 256);
}

int evrp_encoder_forward(const evrp_encoder_weights* w, const float* static_xy, const float* dynamic, int B, int S,
                         int F, float* emb, float* kvl, float* fixed_ctx, float* step_tab, void* workspace,
                         size_t workspace_bytes, void* stream) {
  using namespace evrp;
  if (!w || !static_xy || !dynamic || !emb || !kvl || !fixed_ctx || B < 0 || S < 2 || F < 1 || F >= S ||
      w->n_layers < 0 || w->n_layers > EVRP_MAX_LAYERS)
    return -EVRP_E_BADARG;
  if (S > EVRP_MAX_S) return -EVRP_E_TOO_MANY_NODES;
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (w->gemm_mode == 2) {
    // ONE persistent kernel: init embedding, every layer and the node projection; no activation leaves the SM
    int dev = 0, sms = 148, rc;
    EVRP_CUDA_OK(cudaGetDevice(&dev));
    EVRP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if ((rc = launch_encoder_fused(w, static_xy, dynamic, B, S, F, emb, kvl, step_tab, sms, st))) return rc;
    fixed_context_kernel<<<B, 128, 0, st>>>(emb, S, w->fixed_ctx_w, fixed_ctx);
    EVRP_LAUNCH_OK();
    return 0;
  }
  if (!workspace || workspace_bytes < evrp_encoder_workspace_bytes(B, S)) return -EVRP_E_WORKSPACE;
  const int M = B * S;
  const size_t slab = align_up((size_t)M * 128 * 4, 256);
  float* ha = (float*)workspace;
  float* hb = (float*)((char*)workspace + slab);
  float* heads = (float*)((char*)workspace + 2 * slab);
  float* wide = (float*)((char*)workspace + 3 * slab);

  {
    int dev = 0, sms = 148;
    EVRP_CUDA_OK(cudaGetDevice(&dev));
    EVRP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long want = ((long long)M + 7) / 8;                       // 8 warps (rows) per CTA per sweep
    const unsigned grid = (unsigned)(want < (long long)sms * 16 ? want : (long long)sms * 16);
    init_embed_kernel<<<grid, 256, 0, st>>>(static_xy, dynamic, B, S, F, w->depot_w, w->depot_b, w->station_w,
                                            w->station_b, w->cust_w, w->cust_b, ha);
    EVRP_LAUNCH_OK();
  }
  const size_t attn_smem = (size_t)H * 2 * S * DK * sizeof(float);
  EVRP_CUDA_OK(cudaFuncSetAttribute(self_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem));
  float* cur = ha;
  float* other = hb;
  const bool tcm = (w->gemm_mode == 0);
  int sm_count = 148;
  if (tcm) {
    int dev = 0;
    EVRP_CUDA_OK(cudaGetDevice(&dev));
    EVRP_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  int rc;
  for (int l = 0; l < w->n_layers; ++l) {
    const evrp_layer_weights& L = w->layers[l];
    float* dst = (l == w->n_layers - 1) ? emb : cur;
    if (tcm) {
      // Q|K|V = h @ Wqkv
      if ((rc = tc_gemm(0, cur, D, L.Wqkv_hi, L.Wqkv_lo, wide, 384, M, 384, D, nullptr, nullptr, 0, nullptr, nullptr, sm_count, st))) return rc;
      if (w->attn_mode == 0 && S > 64) {                 // (<= 64 nodes: the 128-row tiles are mostly padding)
        if ((rc = launch_attn_tc(wide, B, S, heads, sm_count, st))) return rc;
      } else {
        self_attention_kernel<<<B, 256, attn_smem, st>>>(wide, S, heads);
        EVRP_LAUNCH_OK();
      }
      // h1 = BN1(h + heads @ Wo)
      if ((rc = tc_gemm(EPI_RESID | EPI_AFFINE, heads, D, L.Wo_hi, L.Wo_lo, other, D, M, D, D, nullptr, cur, D, L.bn1_scale, L.bn1_shift, sm_count, st))) return rc;
      // f = relu(h1 @ W1^T + b1)
      if ((rc = tc_gemm(EPI_BIAS | EPI_RELU, other, D, L.ff1_hi, L.ff1_lo, wide, FF, M, FF, D, L.ff1_b, nullptr, 0, nullptr, nullptr, sm_count, st))) return rc;
      // h2 = BN2(h1 + f @ W2^T + b2)
      if ((rc = tc_gemm(EPI_BIAS | EPI_RESID | EPI_AFFINE, wide, FF, L.ff2_hi, L.ff2_lo, dst, D, M, D, FF, L.ff2_b, other, D, L.bn2_scale, L.bn2_shift, sm_count, st))) return rc;
    } else {
      if ((rc = launch_sgemm<0>(cur, D, L.Wqkv, 384, wide, 384, M, 384, D, nullptr, nullptr, 0, nullptr, nullptr, st))) return rc;
      self_attention_kernel<<<B, 256, attn_smem, st>>>(wide, S, heads);
      EVRP_LAUNCH_OK();
      if ((rc = launch_sgemm<EPI_RESID | EPI_AFFINE>(heads, D, L.Wo, D, other, D, M, D, D, nullptr, cur, D, L.bn1_scale, L.bn1_shift, st))) return rc;
      if ((rc = launch_sgemm<EPI_BIAS | EPI_RELU>(other, D, L.ff1_w, FF, wide, FF, M, FF, D, L.ff1_b, nullptr, 0, nullptr, nullptr, st))) return rc;
      if ((rc = launch_sgemm<EPI_BIAS | EPI_RESID | EPI_AFFINE>(wide, FF, L.ff2_w, D, dst, D, M, D, FF, L.ff2_b, other, D, L.bn2_scale, L.bn2_shift, st))) return rc;
    }
    cur = dst;
  }
  if (w->n_layers == 0) EVRP_CUDA_OK(cudaMemcpyAsync(emb, ha, (size_t)M * D * 4, cudaMemcpyDeviceToDevice, st));
  if (tcm) {
    if ((rc = tc_gemm(0, emb, D, w->node_proj_hi, w->node_proj_lo, kvl, 384, M, 384, D, nullptr, nullptr, 0, nullptr, nullptr, sm_count, st))) return rc;
  } else {
    if ((rc = launch_sgemm<0>(emb, D, w->node_proj, 384, kvl, 384, M, 384, D, nullptr, nullptr, 0, nullptr, nullptr, st))) return rc;
  }
  if (step_tab) {                                        // nets/AM.py:336-337: emb @ project_step_context.weight[:, :128]^T
    if (tcm) {
      if (!w->step_proj_hi || !w->step_proj_lo) return -EVRP_E_BADARG;
      if ((rc = tc_gemm(0, emb, D, w->step_proj_hi, w->step_proj_lo, step_tab, D, M, D, D, nullptr, nullptr, 0, nullptr, nullptr, sm_count, st))) return rc;
    } else {
      if (!w->step_proj) return -EVRP_E_BADARG;
      if ((rc = launch_sgemm<0>(emb, D, w->step_proj, D, step_tab, D, M, D, D, nullptr, nullptr, 0, nullptr, nullptr, st))) return rc;
    }
  }
  fixed_context_kernel<<<B, 128, 0, st>>>(emb, S, w->fixed_ctx_w, fixed_ctx);
  EVRP_LAUNCH_OK();
  return 0;
}

// init embedding alone (training forward): weights in nn.Linear layout (depot [128,2], station [128,2], customer [128,3])
int evrp_init_embed_forward(const float* static_xy, const float* dynamic, int B, int S, int F, const float* depot_w,
                            const float* depot_b, const float* station_w, const float* station_b, const float* cust_w,
                            const float* cust_b, float* h, void* stream) {
  using namespace evrp;
  if (!static_xy || !dynamic || !depot_w || !depot_b || !station_w || !station_b || !cust_w || !cust_b || !h || B < 1 || S < 2 ||
      F < 1 || F >= S)
    return -EVRP_E_BADARG;
  int dev = 0, sms = 148;
  EVRP_CUDA_OK(cudaGetDevice(&dev));
  EVRP_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long want = ((long long)B * S + 7) / 8;
  const unsigned grid = (unsigned)(want < (long long)sms * 16 ? want : (long long)sms * 16);
  init_embed_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(static_xy, dynamic, B, S, F, depot_w, depot_b, station_w, station_b,
                                                            cust_w, cust_b, h);
  EVRP_LAUNCH_OK();
  return 0;
}

size_t evrp_init_embed_backward_workspace_bytes(void) { return (size_t)592 * 10 * 128 * sizeof(float); }

// gradients of the six init-embedding parameters from grad_h [B,S,128]
int evrp_init_embed_backward(const float* static_xy, const float* dynamic, const float* grad_h, int B, int S, int F,
                             float* g_depot_w, float* g_depot_b, float* g_station_w, float* g_station_b, float* g_cust_w,
                             float* g_cust_b, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace evrp;
  if (!static_xy || !dynamic || !grad_h || !g_depot_w || !g_depot_b || !g_station_w || !g_station_b || !g_cust_w || !g_cust_b ||
      !workspace || B < 1 || S < 2 || F < 1 || F >= S)
    return -EVRP_E_BADARG;
  const long long rows = (long long)B * S;
  const int grid = (int)((rows + 3) / 4 < 592 ? (rows + 3) / 4 : 592);
  if (workspace_bytes < (size_t)grid * 10 * 128 * sizeof(float)) return -EVRP_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  init_embed_bwd_kernel<<<grid, 512, 0, st>>>(static_xy, dynamic, grad_h, B, S, F, (float*)workspace);
  EVRP_LAUNCH_OK();
  init_embed_bwd_reduce_kernel<<<10, 128, 0, st>>>((const float*)workspace, grid, g_depot_w, g_depot_b, g_station_w, g_station_b,
                                                   g_cust_w, g_cust_b);
  EVRP_LAUNCH_OK();
  return 0;
}

}  // extern "C"
