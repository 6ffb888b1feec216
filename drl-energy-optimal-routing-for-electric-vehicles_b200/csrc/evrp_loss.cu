// Fused REINFORCE loss / advantage kernel (trainer.py:116-121):
//   reward[b]    = sum_t R[b,t]
//   advantage[b] = reward[b] - baseline[b]
//   loss         = (1/B_global) * sum_b advantage[b] * sum_t ll[b,t]
//   grad_ll[b,t] = advantage[b] / B_global                      (d loss / d ll, used by the backward pass)
// One warp per rollout row (coalesced row reads), then one block reduces the per-row terms in a fixed order so
// the loss is deterministic.  HBM-bound: reads 2*B*T floats, writes B*T.
#include "evrp_common.cuh"

namespace evrp {

__global__ void __launch_bounds__(256)
reinforce_rows_kernel(const float* __restrict__ ll, const float* __restrict__ R, const float* __restrict__ baseline,
                      int baseline_is_scalar, int B, int T, float inv_global_batch, const float* __restrict__ n_global_dev,
                      float* __restrict__ reward, float* __restrict__ advantage, float* __restrict__ grad_ll,
                      float* __restrict__ term) {
  if (n_global_dev) inv_global_batch = __fdiv_rn(1.f, *n_global_dev);     // the global batch size stays on the device (multi-rank)
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* lrow = ll + (size_t)b * T;
  const float* rrow = R + (size_t)b * T;
  float rs = 0.f, ls = 0.f;
  for (int t = lane; t < T; t += 32) { rs += rrow[t]; ls += lrow[t]; }
  rs = warp_sum(rs);
  ls = warp_sum(ls);
  const float adv = rs - (baseline_is_scalar ? baseline[0] : baseline[b]);
  const float g = adv * inv_global_batch;
  for (int t = lane; t < T; t += 32) grad_ll[(size_t)b * T + t] = g;
  if (lane == 0) { reward[b] = rs; advantage[b] = adv; term[b] = adv * ls; }
}

__global__ void __launch_bounds__(1024)
reinforce_reduce_kernel(const float* __restrict__ term, int B, float inv_global_batch, const float* __restrict__ n_global_dev,
                        float* __restrict__ loss) {
  if (n_global_dev) inv_global_batch = __fdiv_rn(1.f, *n_global_dev);
  __shared__ double part[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < B; i += 1024) s += (double)term[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = part[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (threadIdx.x == 0) loss[0] = (float)(s * (double)inv_global_batch);
  }
}

}  // namespace evrp

extern "C" int evrp_reinforce_loss(const float* ll, const float* R, const float* baseline, int baseline_is_scalar, int B,
                                   int T, float inv_global_batch, const float* global_batch_dev, float* reward, float* advantage,
                                   float* grad_ll, float* loss, float* scratch, void* stream) {
  if (!ll || !R || !baseline || !reward || !advantage || !grad_ll || !loss || !scratch || B < 1 || T < 1)
    return -EVRP_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  evrp::reinforce_rows_kernel<<<(B + 7) / 8, 256, 0, st>>>(ll, R, baseline, baseline_is_scalar, B, T, inv_global_batch,
                                                          global_batch_dev, reward, advantage, grad_ll, scratch);
  EVRP_LAUNCH_OK();
  evrp::reinforce_reduce_kernel<<<1, 1024, 0, st>>>(scratch, B, inv_global_batch, global_batch_dev, loss);
  EVRP_LAUNCH_OK();
  return 0;
}
