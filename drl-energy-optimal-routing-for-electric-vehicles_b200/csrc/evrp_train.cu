// Small kernels of the REINFORCE training step (trainer.py:115-127) that used to be chains of eager
// PyTorch ops:
//   split_tf32_kernel              weight [rows, cols] (any strides) -> TF32 (hi, lo) planes for evrp_gemm_3xtf32
//   bn_*                           train-mode BatchNorm1d(128) over [M,128] rows, nets/GraphEncoder.py:142-150: batch
//                                  statistics as fp64 sums [sum x, sum x^2, n] (ONE all-reduce of 257 doubles makes them
//                                  global-batch statistics when instances are sharded over ranks), normalise + running
//                                  statistics; backward [sum g, sum g*xhat] -> one all-reduce -> input gradient
//   route_context_*                step-query assembly of the teacher-forced decoder, nets/DRLModel.py:203-237,394:
//                                  running max of the visited embeddings (zeros before the first pick) and
//                                  fixed_ctx + emb[now] for all T steps, and their backward (scatter of the gradients to
//                                  the node that holds the running max / the current node), no [B,T,128] gathers
//   relu_mask_kernel               g * (y > 0)
#include <cuda_fp16.h>
#include "evrp_common.cuh"

namespace evrp {
namespace trn {

__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }   // = evrp_tc.cuh

__global__ void split_tf32_kernel(const float* __restrict__ w, int rows, int cols, long long sr, long long sc,
                                  float* __restrict__ hi, float* __restrict__ lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i % cols;
  const float v = w[(long long)r * sr + (long long)c * sc];
  const float h = __uint_as_float(to_tf32(v));
  hi[i] = h;
  lo[i] = __uint_as_float(to_tf32(v - h));
}

__global__ void relu_mask_kernel(const float4* __restrict__ g, const float4* __restrict__ y, float4* __restrict__ out, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 a = g[i], b = y[i];
  out[i] = make_float4(b.x > 0.f ? a.x : 0.f, b.y > 0.f ? a.y : 0.f, b.z > 0.f ? a.z : 0.f, b.w > 0.f ? a.w : 0.f);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm1d(128), training mode.  256 threads: lane group c4 = tid & 31 owns channels 4*c4 .. 4*c4+3, rsub = tid >> 5
// strides the rows.  Per-thread fp32 partial sums (a few dozen rows), cross-thread / cross-block reduction in fp64.
// ------------------------------------------------------------------------------------------------
constexpr int BN_C = 128;

struct D4 { double x, y, z, w; };

__device__ __forceinline__ void block_channel_sums(const D4& a, const D4& b, double* __restrict__ out_a, double* __restrict__ out_b) {
  __shared__ D4 sa[8][32], sb[8][32];
  const int c4 = threadIdx.x & 31, rsub = threadIdx.x >> 5;
  sa[rsub][c4] = a; sb[rsub][c4] = b;
  __syncthreads();
  if (threadIdx.x < 32) {
    double ax = 0, ay = 0, az = 0, aw = 0, bx = 0, by = 0, bz = 0, bw = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const D4 u = sa[r][c4], v = sb[r][c4];
      ax += u.x; ay += u.y; az += u.z; aw += u.w; bx += v.x; by += v.y; bz += v.z; bw += v.w;
    }
    atomicAdd(out_a + 4 * c4 + 0, ax); atomicAdd(out_a + 4 * c4 + 1, ay); atomicAdd(out_a + 4 * c4 + 2, az); atomicAdd(out_a + 4 * c4 + 3, aw);
    atomicAdd(out_b + 4 * c4 + 0, bx); atomicAdd(out_b + 4 * c4 + 1, by); atomicAdd(out_b + 4 * c4 + 2, bz); atomicAdd(out_b + 4 * c4 + 3, bw);
  }
}

// sums [257] fp64, zeroed by the caller: sum x [128] | sum x^2 [128] | n.  Every addition is fp64 (products exact in fp64), so
// the statistics do not depend on how the rows are split over threads, blocks or RANKS beyond fp64 rounding: a sharded batch
// normalises bit-identically to the same batch on one GPU (a 1e-7 difference in the mean flips ReLU gates of the layer
// behind and shows up as 1e-3 noise in single weight-gradient entries)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, int M, double* __restrict__ sums) {
  const int c4 = threadIdx.x & 31, rsub = threadIdx.x >> 5;
  D4 s = {0, 0, 0, 0}, ss = {0, 0, 0, 0};
  for (int row = blockIdx.x * 8 + rsub; row < M; row += gridDim.x * 8) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)row * BN_C) + c4);
    const double a = v.x, b = v.y, c = v.z, d = v.w;
    s.x += a; s.y += b; s.z += c; s.w += d;
    ss.x += a * a; ss.y += b * b; ss.z += c * c; ss.w += d * d;
  }
  block_channel_sums(s, ss, sums, sums + BN_C);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(sums + 2 * BN_C, (double)M);
}

// y = (x - mean) * invstd * weight + bias with the (global) sums; save [257] fp32 = mean | invstd | n for the backward;
// block 0 updates the running statistics like F.batch_norm (momentum, unbiased variance)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, int M, const double* __restrict__ sums,
                                                       const float* __restrict__ weight, const float* __restrict__ bias,
                                                       float eps, float momentum, float* __restrict__ running_mean,
                                                       float* __restrict__ running_var, float* __restrict__ save,
                                                       float* __restrict__ y) {
  __shared__ float s_scale[BN_C], s_shift[BN_C];
  if (threadIdx.x < BN_C) {
    const int c = threadIdx.x;
    const double n = sums[2 * BN_C];
    const double mean = sums[c] / n;
    double var = sums[BN_C + c] / n - mean * mean;                 // biased (normalisation)
    if (var < 0) var = 0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = invstd * weight[c];
    s_scale[c] = sc;
    s_shift[c] = fmaf(-(float)mean, sc, bias[c]);
    if (blockIdx.x == 0) {
      save[c] = (float)mean; save[BN_C + c] = invstd;
      if (c == 0) save[2 * BN_C] = (float)n;
      if (running_mean) {
        const double unb = n > 1 ? var * (n / (n - 1)) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
      }
    }
  }
  __syncthreads();
  const int c4 = threadIdx.x & 31, rsub = threadIdx.x >> 5;
  const float4 sc = *reinterpret_cast<const float4*>(&s_scale[4 * c4]), sh = *reinterpret_cast<const float4*>(&s_shift[4 * c4]);
  for (int row = blockIdx.x * 8 + rsub; row < M; row += gridDim.x * 8) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)row * BN_C) + c4);
    reinterpret_cast<float4*>(y + (size_t)row * BN_C)[c4] =
        make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
  }
}

// sums [256] fp64, zeroed by the caller: sum g | sum g * xhat     (xhat recomputed from x and the saved mean / invstd)
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ g, int M,
                                                            const float* __restrict__ save, double* __restrict__ sums) {
  const int c4 = threadIdx.x & 31, rsub = threadIdx.x >> 5;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(save) + c4), is = __ldg(reinterpret_cast<const float4*>(save + BN_C) + c4);
  D4 sg = {0, 0, 0, 0}, sx = {0, 0, 0, 0};
  for (int row = blockIdx.x * 8 + rsub; row < M; row += gridDim.x * 8) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)row * BN_C) + c4);
    const float4 d = __ldg(reinterpret_cast<const float4*>(g + (size_t)row * BN_C) + c4);
    sg.x += (double)d.x; sg.y += (double)d.y; sg.z += (double)d.z; sg.w += (double)d.w;
    sx.x += (double)d.x * (double)((v.x - mu.x) * is.x); sx.y += (double)d.y * (double)((v.y - mu.y) * is.y);
    sx.z += (double)d.z * (double)((v.z - mu.z) * is.z); sx.w += (double)d.w * (double)((v.w - mu.w) * is.w);
  }
  block_channel_sums(sg, sx, sums, sums + BN_C);
}

// gx = weight * invstd * (g - sum_g / n - xhat * sum_gx / n)   with the GLOBAL sums and row count
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ g, int M,
                                                           const float* __restrict__ save, const float* __restrict__ weight,
                                                           const double* __restrict__ sums, float* __restrict__ gx) {
  __shared__ float s_a[BN_C], s_b[BN_C], s_c[BN_C], s_mu[BN_C];
  if (threadIdx.x < BN_C) {
    const int c = threadIdx.x;
    const double n = (double)save[2 * BN_C];
    const float is = save[BN_C + c];
    const float k = weight[c] * is;
    s_a[c] = k;                                            // gx = k * g - k * mg - (k * is * mgx) * (x - mu)
    s_b[c] = k * (float)(sums[c] / n);
    s_c[c] = k * is * (float)(sums[BN_C + c] / n);
    s_mu[c] = save[c];
  }
  __syncthreads();
  const int c4 = threadIdx.x & 31, rsub = threadIdx.x >> 5;
  const float4 a = *reinterpret_cast<const float4*>(&s_a[4 * c4]), b = *reinterpret_cast<const float4*>(&s_b[4 * c4]);
  const float4 cc = *reinterpret_cast<const float4*>(&s_c[4 * c4]), mu = *reinterpret_cast<const float4*>(&s_mu[4 * c4]);
  for (int row = blockIdx.x * 8 + rsub; row < M; row += gridDim.x * 8) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)row * BN_C) + c4);
    const float4 d = __ldg(reinterpret_cast<const float4*>(g + (size_t)row * BN_C) + c4);
    reinterpret_cast<float4*>(gx + (size_t)row * BN_C)[c4] =
        make_float4(fmaf(a.x, d.x, -b.x) - cc.x * (v.x - mu.x), fmaf(a.y, d.y, -b.y) - cc.y * (v.y - mu.y),
                    fmaf(a.z, d.z, -b.z) - cc.z * (v.z - mu.z), fmaf(a.w, d.w, -b.w) - cc.w * (v.w - mu.w));
  }
}

// ------------------------------------------------------------------------------------------------
// route context: one CTA per instance, thread = channel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) route_context_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ fixed,
                                                                const int64_t* __restrict__ pi, int T, int S,
                                                                float* __restrict__ run_max, float* __restrict__ base) {
  extern __shared__ int s_pi[];
  const int b = blockIdx.x, c = threadIdx.x;
  for (int t = c; t < T; t += 128) {
    const int j = (int)pi[(size_t)b * T + t];
    s_pi[t] = min(max(j, 0), S - 1);
  }
  __syncthreads();
  const float* e = emb + (size_t)b * S * D + c;
  const float fx = fixed[(size_t)b * D + c];
  float* rm = run_max + (size_t)b * T * D + c;
  float* bs = base + (size_t)b * T * D + c;
  float cur = 0.f;
  int now = 0;
#pragma unroll 4
  for (int t = 0; t < T; ++t) {
    const int j = s_pi[t];
    const float vn = __ldg(e + (size_t)now * D), vj = __ldg(e + (size_t)j * D);
    rm[(size_t)t * D] = cur;                                // max over the picks of steps < t; zeros at t = 0 (:224)
    bs[(size_t)t * D] = fx + vn;
    cur = (t == 0) ? vj : fmaxf(cur, vj);
    now = j;
  }
}

// d emb [B,S,128] (written, not accumulated) and d fixed [B,128] from d run_max and d base
__global__ void __launch_bounds__(128) route_context_bwd_kernel(const float* __restrict__ emb, const int64_t* __restrict__ pi,
                                                                const float* __restrict__ g_rm, const float* __restrict__ g_base,
                                                                int T, int S, float* __restrict__ d_emb,
                                                                float* __restrict__ d_fixed) {
  extern __shared__ unsigned char s_raw[];
  float* acc = reinterpret_cast<float*>(s_raw);            // [S][128]: thread c owns column c (no conflicts, no atomics)
  int* s_pi = reinterpret_cast<int*>(acc + (size_t)S * D);
  const int b = blockIdx.x, c = threadIdx.x;
  for (int t = c; t < T; t += 128) {
    const int j = (int)pi[(size_t)b * T + t];
    s_pi[t] = min(max(j, 0), S - 1);
  }
  for (int j = 0; j < S; ++j) acc[j * D + c] = 0.f;
  __syncthreads();
  const float* e = emb + (size_t)b * S * D + c;
  const float* gb = g_base + (size_t)b * T * D + c;
  float dfx = 0.f;
  int now = 0;
#pragma unroll 8
  for (int t = 0; t < T; ++t) {                             // base[t] = fixed + emb[now_t]
    const float g = __ldg(gb + (size_t)t * D);
    dfx += g;
    acc[now * D + c] += g;
    now = s_pi[t];
  }
  if (g_rm) {
    const float* gr = g_rm + (size_t)b * T * D + c;
    float cur = 0.f, pend = 0.f;
    int arg = -1;
#pragma unroll 8
    for (int u = 0; u + 1 < T; ++u) {                       // pick u feeds run_max[t] for t > u while it holds the maximum
      const int j = s_pi[u];
      const float v = __ldg(e + (size_t)j * D);
      if (u == 0 || v >= cur) {                             // ties go to the latest pick, like torch.cummax
        if (arg >= 0 && arg != j) { acc[arg * D + c] += pend; pend = 0.f; }
        cur = v; arg = j;
      }
      pend += __ldg(gr + (size_t)(u + 1) * D);
    }
    if (arg >= 0) acc[arg * D + c] += pend;
  }
  float* de = d_emb + (size_t)b * S * D + c;
  for (int j = 0; j < S; ++j) de[(size_t)j * D] = acc[j * D + c];
  d_fixed[(size_t)b * D + c] = dfx;
}


// ------------------------------------------------------------------------------------------------
// decoder weight packing on the device (evrp_pack_decoder)
// ------------------------------------------------------------------------------------------------
// scratch[0] = max |W1[:, 128:]|, scratch[1] = max |W2| (as float bit patterns: non-negative floats order like unsigned ints)
__global__ void pack_decoder_amax_kernel(const float* __restrict__ w1, const float* __restrict__ w2, unsigned int* __restrict__ scratch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;                 // 0 .. 65535 twice
  float a = 0.f, b = 0.f;
  if (i < 512 * 128) { a = fabsf(w1[(i >> 7) * 256 + 128 + (i & 127)]); b = fabsf(w2[i]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a = fmaxf(a, __shfl_xor_sync(FULL, a, o)); b = fmaxf(b, __shfl_xor_sync(FULL, b, o)); }
  if ((threadIdx.x & 31) == 0) { atomicMax(scratch, __float_as_uint(a)); atomicMax(scratch + 1, __float_as_uint(b)); }
}

__device__ __forceinline__ float pow2_scale(float amax, float& inv) {
  int e = 0;
  if (amax > 0.f) { int ex; frexpf(amax, &ex); e = 12 - ex; }
  e = max(-14, min(24, e));
  inv = ldexpf(1.f, -e);
  return ldexpf(1.f, e);
}

__global__ void pack_decoder_kernel(const float* __restrict__ tour_w, const float* __restrict__ w1, const float* __restrict__ w2,
                                    const float* __restrict__ scratch, float* __restrict__ w1_veh, float* __restrict__ w1_max,
                                    float* __restrict__ w2_t, __half* __restrict__ w1h, __half* __restrict__ w1l,
                                    __half* __restrict__ w2h, __half* __restrict__ w2l, float* __restrict__ inv_scales,
                                    __half* __restrict__ ffs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;                 // one thread per element of a 512 x 128 matrix
  if (i >= 512 * 128) return;
  float inv1, inv2;
  const float s1 = pow2_scale(scratch[0], inv1), s2 = pow2_scale(scratch[1], inv2);
  if (i == 0) { inv_scales[0] = inv1; inv_scales[1] = inv2; }
  {                                                                    // W1[:, 128:]  [512 out][128 in]
    const int f = i >> 7, k = i & 127;
    const float v = w1[f * 256 + 128 + k];
    w1_max[k * 512 + f] = v;                                           // k-major copy (FFMA kernel)
    const float ws = v * s1;                                           // exact (power of two)
    const __half hi = __float2half_rn(ws);
    const __half lo = __float2half_rn(ws - __half2float(hi));
    w1h[i] = hi; w1l[i] = lo;
    if (ffs) {                                                         // stage (f / 128, k / 64): [hi 8192 | lo 8192] halfs
      const int r = f & 127, kk = k & 63;
      const int off = ((f >> 7) * 2 + (k >> 6)) * 16384 + r * 64 + ((((kk >> 3) ^ (r & 7)) << 3) | (kk & 7));
      ffs[off] = hi; ffs[off + 8192] = lo;
    }
  }
  {                                                                    // W2  [128 out][512 in]
    const int o = i >> 9, k = i & 511;
    const float v = w2[i];
    w2_t[k * 128 + o] = v;
    const float ws = v * s2;
    const __half hi = __float2half_rn(ws);
    const __half lo = __float2half_rn(ws - __half2float(hi));
    w2h[i] = hi; w2l[i] = lo;
    if (ffs) {                                                         // stage 8 + k / 64
      const int kk = k & 63;
      const int off = (8 + (k >> 6)) * 16384 + o * 64 + ((((kk >> 3) ^ (o & 7)) << 3) | (kk & 7));
      ffs[off] = hi; ffs[off + 8192] = lo;
    }
  }
  if (i < 512 * 3) {                                                   // fold: (W1[:, :128] @ tour.weight)[f][c], fp64 like the host fold
    const int f = i / 3, c = i % 3;
    double acc = 0.0;
    for (int k = 0; k < 128; ++k) acc += (double)w1[f * 256 + k] * (double)tour_w[k * 3 + c];
    w1_veh[c * 512 + f] = (float)acc;
  }
}

}  // namespace trn
}  // namespace evrp

using namespace evrp;

extern "C" int evrp_split_tf32(const float* w, int rows, int cols, long long stride_r, long long stride_c, float* hi,
                               float* lo, void* stream) {
  if (!w || !hi || !lo || rows < 1 || cols < 1) return -EVRP_E_BADARG;
  const int n = rows * cols;
  trn::split_tf32_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, rows, cols, stride_r, stride_c, hi, lo);
  EVRP_LAUNCH_OK();
  return 0;
}

extern "C" int evrp_relu_mask(const float* g, const float* y, float* out, long long n, void* stream) {
  if (!g || !y || !out || n < 0 || (n & 3)) return -EVRP_E_BADARG;
  if (n == 0) return 0;
  const size_t n4 = (size_t)n >> 2;
  trn::relu_mask_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(y), reinterpret_cast<float4*>(out), n4);
  EVRP_LAUNCH_OK();
  return 0;
}

static int bn_grid(int M) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int g = (M + 7) / 8;
  const int cap = sms * 4;
  return g < cap ? (g > 0 ? g : 1) : cap;
}

extern "C" int evrp_bn_stats(const float* x, int M, double* sums /*[257], zeroed here*/, void* stream) {
  if (!x || !sums || M < 1) return -EVRP_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  EVRP_CUDA_OK(cudaMemsetAsync(sums, 0, 257 * sizeof(double), st));
  trn::bn_stats_kernel<<<bn_grid(M), 256, 0, st>>>(x, M, sums);
  EVRP_LAUNCH_OK();
  return 0;
}

extern "C" int evrp_bn_apply(const float* x, int M, const double* sums, const float* weight, const float* bias, float eps,
                             float momentum, float* running_mean, float* running_var, float* save /*[257]*/, float* y,
                             void* stream) {
  if (!x || !sums || !weight || !bias || !save || !y || M < 1 || (running_mean && !running_var)) return -EVRP_E_BADARG;
  trn::bn_apply_kernel<<<bn_grid(M), 256, 0, (cudaStream_t)stream>>>(x, M, sums, weight, bias, eps, momentum, running_mean,
                                                                     running_var, save, y);
  EVRP_LAUNCH_OK();
  return 0;
}

extern "C" int evrp_bn_backward_reduce(const float* x, const float* g, int M, const float* save, double* sums /*[256], zeroed here*/,
                                       void* stream) {
  if (!x || !g || !save || !sums || M < 1) return -EVRP_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  EVRP_CUDA_OK(cudaMemsetAsync(sums, 0, 256 * sizeof(double), st));
  trn::bn_bwd_reduce_kernel<<<bn_grid(M), 256, 0, st>>>(x, g, M, save, sums);
  EVRP_LAUNCH_OK();
  return 0;
}

extern "C" int evrp_bn_backward_apply(const float* x, const float* g, int M, const float* save, const float* weight,
                                      const double* sums, float* gx, void* stream) {
  if (!x || !g || !save || !weight || !sums || !gx || M < 1) return -EVRP_E_BADARG;
  trn::bn_bwd_apply_kernel<<<bn_grid(M), 256, 0, (cudaStream_t)stream>>>(x, g, M, save, weight, sums, gx);
  EVRP_LAUNCH_OK();
  return 0;
}

extern "C" int evrp_route_context_forward(const float* emb, const float* fixed_ctx, const int64_t* pi, int B, int T, int S,
                                          float* run_max, float* base, void* stream) {
  if (!emb || !fixed_ctx || !pi || !run_max || !base || B < 1 || T < 1 || S < 1 || S > EVRP_MAX_S) return -EVRP_E_BADARG;
  const size_t smem = (size_t)T * sizeof(int);
  if (smem > 48 * 1024) EVRP_CUDA_OK(cudaFuncSetAttribute(trn::route_context_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  trn::route_context_fwd_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(emb, fixed_ctx, pi, T, S, run_max, base);
  EVRP_LAUNCH_OK();
  return 0;
}

extern "C" int evrp_route_context_backward(const float* emb, const int64_t* pi, const float* grad_run_max /*nullable*/,
                                           const float* grad_base, int B, int T, int S, float* grad_emb, float* grad_fixed,
                                           void* stream) {
  if (!emb || !pi || !grad_base || !grad_emb || !grad_fixed || B < 1 || T < 1 || S < 1 || S > EVRP_MAX_S) return -EVRP_E_BADARG;
  const size_t smem = (size_t)S * D * sizeof(float) + (size_t)T * sizeof(int);
  EVRP_CUDA_OK(cudaFuncSetAttribute(trn::route_context_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  trn::route_context_bwd_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(emb, pi, grad_run_max, grad_base, T, S, grad_emb,
                                                                        grad_fixed);
  EVRP_LAUNCH_OK();
  return 0;
}

extern "C" int evrp_pack_decoder(const float* tour_w, const float* w1, const float* w2, float* w1_veh, float* w1_max, float* w2_t,
                                 void* w1max_hi, void* w1max_lo, void* w2_hi, void* w2_lo, float* inv_scales, float* scratch,
                                 void* ff_stream, void* stream) {
  if (!tour_w || !w1 || !w2 || !w1_veh || !w1_max || !w2_t || !w1max_hi || !w1max_lo || !w2_hi || !w2_lo || !inv_scales || !scratch)
    return -EVRP_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  EVRP_CUDA_OK(cudaMemsetAsync(scratch, 0, 2 * sizeof(float), st));
  trn::pack_decoder_amax_kernel<<<256, 256, 0, st>>>(w1, w2, reinterpret_cast<unsigned int*>(scratch));
  EVRP_LAUNCH_OK();
  trn::pack_decoder_kernel<<<256, 256, 0, st>>>(tour_w, w1, w2, scratch, w1_veh, w1_max, w2_t, (__half*)w1max_hi, (__half*)w1max_lo,
                                              (__half*)w2_hi, (__half*)w2_lo, inv_scales, (__half*)ff_stream);
  EVRP_LAUNCH_OK();
  return 0;
}
