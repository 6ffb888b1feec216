// Shared device code: the EVRP environment step (a7 update_dynamic + a8 update_mask of SURVEY.md §8)
// as scalar, separately-rounded fp32 arithmetic.  Every op is an explicit __f*_rn intrinsic so that nvcc
// cannot contract a*b+c into an FMA: the recipe must reproduce the reference's op-by-op rounding bit for bit
// (problems/EVRP.py:105-280).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/evrp_b200.h"

#define EVRP_CUDA_OK(expr)                                 \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return -(int)_e;                \
  } while (0)

#define EVRP_LAUNCH_OK()                                   \
  do {                                                     \
    cudaError_t _e = cudaGetLastError();                   \
    if (_e != cudaSuccess) return -(int)_e;                \
  } while (0)

namespace evrp {

constexpr int D = EVRP_D;      // 128
constexpr int H = EVRP_H;      // 8
constexpr int DK = D / H;      // 16
constexpr int FF = EVRP_FF;    // 512
constexpr unsigned FULL = 0xffffffffu;

// x / d for a loop-invariant divisor.  mode 1: x * fl(1/d) (ATen CUDA convention).  mode 0: IEEE division.
// For the two divisors of the reference configuration (velocity = 50, 3600 s/h) the correctly rounded quotient is
// obtained with one FMA refinement of x * fl(1/d) (Markstein): q0 = x*r; q = fma(fma(-d, q0, x), r, q0).  This was
// checked EXHAUSTIVELY against x / d for every finite fp32 x (tools/divcheck.c): identical for all |x| >= 2e-37;
// smaller non-zero magnitudes (never produced by this model) take the generic division.
// generic IEEE division, kept out of line: it is the rare path of div_scalar, and one shared copy keeps the
// instruction footprint of the rollout kernels' row loop small (the loop is instruction-cache sensitive)
static __device__ __noinline__ float div_generic(float x, float d) { return __fdiv_rn(x, d); }

__device__ __forceinline__ float div_scalar(float x, float d, float inv_d, int mode) {
  if (mode) return __fmul_rn(x, inv_d);
  if ((d == 50.f || d == 3600.f) && (x == 0.f || fabsf(x) >= 1e-30f)) {
    const float q0 = __fmul_rn(x, inv_d);
    return __fmaf_rn(__fmaf_rn(-d, q0, x), inv_d, q0);
  }
  return div_generic(x, d);
}

// soc consumption of one leg (problems/EVRP.py:246-252 and the four copies inside update_mask :155-211)
__device__ __forceinline__ float leg_energy(const evrp_phys& p, float mass, float slope, float tc) {
  const float mg = __fmul_rn(mass, p.g);
  const float Pm = __fmul_rn(__fadd_rn(__fadd_rn(p.aero, __fmul_rn(mg, slope)), __fmul_rn(mg, p.Cr)), p.velocity);
  // drive (Pm > 0) and recuperation (Pm < 0) differ only in the efficiency factor: one division site (same ops, same bits)
  const float k = (Pm > 0.f) ? p.kd : p.kr;
  const float e = div_scalar(__fmul_rn(__fmul_rn(k, Pm), tc), 3600.f, p.inv_3600, p.div_mode);
  return (Pm > 0.f || Pm < 0.f) ? e : 0.f;           // Pm == 0 (and NaN) contribute nothing, as in the reference
}

__device__ __forceinline__ float vehicle_mass(const evrp_phys& p, float load) {
  // self.mc + loads*self.max_load*self.w   (problems/EVRP.py:155,246)
  return __fadd_rn(p.mc, __fmul_rn(__fmul_rn(load, p.max_load), p.w));
}

__device__ __forceinline__ float travel_time(const evrp_phys& p, float dist) {
  return div_scalar(dist, p.velocity, p.inv_velocity, p.div_mode);
}

// Feasibility of node j after arriving at node c (rules 5-7 of problems/EVRP.py:150-218) given the
// post-move scalars.  d0/g0 = D[c,j], G[c,j]; D0j/G0j = D[0,j], G[0,j]; d3/s3/d4 = rule-7 constants.
// load_j / dem_j are the (row) load value at column j and the demand of j (loads - demands, :177).
__device__ __forceinline__ bool reachable(const evrp_phys& p, int j, int F, float load0, float load_j, float dem_j,
                                          float soc_j, float time_j, float d0, float g0, float D0j, float G0j,
                                          float d3, float s3, float d4) {
  const float tc0 = travel_time(p, d0);
  const float soc0 = leg_energy(p, vehicle_mass(p, load0), g0, tc0);
  if (soc_j < soc0 || time_j < tc0) return false;                                   // :216
  const float mass2 = vehicle_mass(p, __fsub_rn(load_j, dem_j));                    // :177
  const float soc2 = leg_energy(p, mass2, G0j, travel_time(p, D0j));                // :178-186
  float soc3 = __fadd_rn(soc0, soc2);                                               // :188
  if (j <= F) soc3 = 0.f;                                                           // :189
  float tc3 = travel_time(p, __fadd_rn(d0, D0j));                                   // :190
  if (j >= 1 && j <= F) tc3 = __fadd_rn(tc3, p.charging_time);                      // :191
  else if (j > F) tc3 = __fadd_rn(tc3, p.serve_time);                               // :192
  const float soc4 = leg_energy(p, mass2, s3, travel_time(p, d3));                  // :203-211
  const float soc5 = __fadd_rn(soc0, soc4);                                         // :212
  float tc5 = travel_time(p, __fadd_rn(__fadd_rn(d0, d3), d4));                     // :213
  if (j > F) tc5 = __fadd_rn(tc5, p.charge_plus_serve);                             // :214
  const bool via_depot_bad = (soc_j < soc3) || (time_j < tc3);
  const bool via_station_bad = (soc_j < soc5) || (time_j < tc5);
  return !(via_depot_bad && via_station_bad);                                       // :218
}

// ---- call-free variant for the rollout kernels' row loop (instruction-cache footprint) ----
// div_fast is the FMA-refined quotient of div_scalar WITHOUT the range guard: it equals IEEE x / d for d in {50, 3600}
// and every x with |x| >= 2e-37 or x == 0 (tools/divcheck.c); `bad` collects the (never observed) inputs outside that
// range, for which the caller re-evaluates with the generic out-of-line version below.
__device__ __forceinline__ float div_fast(float x, float d, float inv_d, bool& bad) {
  bad |= (x != 0.f) && (fabsf(x) < 1e-30f);
  const float q0 = __fmul_rn(x, inv_d);
  return __fmaf_rn(__fmaf_rn(-d, q0, x), inv_d, q0);
}
__device__ __forceinline__ float leg_energy_fast(const evrp_phys& p, float mass, float slope, float tc, bool& bad) {
  const float mg = __fmul_rn(mass, p.g);
  const float Pm = __fmul_rn(__fadd_rn(__fadd_rn(p.aero, __fmul_rn(mg, slope)), __fmul_rn(mg, p.Cr)), p.velocity);
  const float k = (Pm > 0.f) ? p.kd : p.kr;
  const float e = div_fast(__fmul_rn(__fmul_rn(k, Pm), tc), 3600.f, p.inv_3600, bad);
  return (Pm > 0.f || Pm < 0.f) ? e : 0.f;
}
// same statements as reachable(); legal when div_mode == 0 and velocity == 50 (the caller checks once per kernel)
__device__ __forceinline__ bool reachable_fast(const evrp_phys& p, int j, int F, float load0, float load_j, float dem_j,
                                               float soc_j, float time_j, float d0, float g0, float D0j, float G0j,
                                               float d3, float s3, float d4, bool& bad) {
  const float v = p.velocity, iv = p.inv_velocity;
  const float tc0 = div_fast(d0, v, iv, bad);
  const float soc0 = leg_energy_fast(p, vehicle_mass(p, load0), g0, tc0, bad);
  const float mass2 = vehicle_mass(p, __fsub_rn(load_j, dem_j));
  const float soc2 = leg_energy_fast(p, mass2, G0j, div_fast(D0j, v, iv, bad), bad);
  float soc3 = __fadd_rn(soc0, soc2);
  if (j <= F) soc3 = 0.f;
  float tc3 = div_fast(__fadd_rn(d0, D0j), v, iv, bad);
  if (j >= 1 && j <= F) tc3 = __fadd_rn(tc3, p.charging_time);
  else if (j > F) tc3 = __fadd_rn(tc3, p.serve_time);
  const float soc4 = leg_energy_fast(p, mass2, s3, div_fast(d3, v, iv, bad), bad);
  const float soc5 = __fadd_rn(soc0, soc4);
  float tc5 = div_fast(__fadd_rn(__fadd_rn(d0, d3), d4), v, iv, bad);
  if (j > F) tc5 = __fadd_rn(tc5, p.charge_plus_serve);
  const bool direct_bad = (soc_j < soc0) || (time_j < tc0);
  const bool via_depot_bad = (soc_j < soc3) || (time_j < tc3);
  const bool via_station_bad = (soc_j < soc5) || (time_j < tc5);
  return !direct_bad && !(via_depot_bad && via_station_bad);
}
static __device__ __noinline__ bool reachable_generic(const evrp_phys& p, int j, int F, float load0, float load_j,
                                                      float dem_j, float soc_j, float time_j, float d0, float g0,
                                                      float D0j, float G0j, float d3, float s3, float d4) {
  return reachable(p, j, F, load0, load_j, dem_j, soc_j, time_j, d0, g0, D0j, G0j, d3, s3, d4);
}

// Pairwise geometry of one instance.  Either the caller's [S,S] distance / slope matrices, or -- 3-tuple input
// (static, dynamic, Elevations), nets/DRLModel.py:124-133 -- the coordinate and elevation rows themselves: the entries
// are then recomputed where they are needed, with the op order of pairwise_build_kernel (evrp_step.cu), bit for bit:
//   D[i,j] = sqrt((x_i - x_j)^2 + (y_i - y_j)^2)      G[i,j] = clamp((E_i - E_j) / (D[i,j] + 1e-6), -0.1, 0.1)
// so that no [B,S,S] tensor (736 MB at B = 8192, N = 100) has to exist at all.
template <bool COORDS>                   // compile-time: the matrix-mode kernels carry none of the recompute code
struct Geo {
  const float* D; const float* G;        // [S,S] of this instance, or nullptr
  const float* x; const float* y; const float* e;   // [S] rows of this instance (coordinate mode)
  int S;
  __device__ __forceinline__ static constexpr bool coords() { return COORDS; }
  __device__ __forceinline__ void pair(float xi, float yi, float ei, int j, float& d, float& g) const {
    const float dx = __fsub_rn(xi, __ldg(x + j)), dy = __fsub_rn(yi, __ldg(y + j));
    d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    g = fminf(fmaxf(__fdiv_rn(__fsub_rn(ei, __ldg(e + j)), __fadd_rn(d, 0.000001f)), -0.10f), 0.10f);
  }
  __device__ __forceinline__ void at(int i, int j, float& d, float& g) const {
    if (coords()) pair(__ldg(x + i), __ldg(y + i), __ldg(e + i), j, d, g);
    else { d = __ldg(D + (size_t)i * S + j); g = __ldg(G + (size_t)i * S + j); }
  }
};
template <bool COORDS>
__device__ __forceinline__ Geo<COORDS> make_geo(const float* Dm, const float* Gm, const float* xy, const float* elev, int inst, int S) {
  Geo<COORDS> g;
  g.S = S;
  g.D = COORDS ? nullptr : Dm + (size_t)inst * S * S;
  g.G = COORDS ? nullptr : Gm + (size_t)inst * S * S;
  g.x = COORDS ? xy + (size_t)inst * 2 * S : nullptr;
  g.y = COORDS ? xy + (size_t)inst * 2 * S + S : nullptr;
  g.e = COORDS ? elev + (size_t)inst * S : nullptr;
  return g;
}

// Counter-based RNG for sampling: splitmix64 finaliser over (seed, row, step) -> uniform in [0,1).
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t row, uint32_t step) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (row * 1024ull + step + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);
}

// Packed fp32 FMA (Blackwell FFMA2, PTX fma.rn.f32x2): d.{x,y} = a.{x,y} * b.{x,y} + d.{x,y}; each lane is an
// ordinary IEEE fp32 fused multiply-add, so results are bit-identical to two scalar fmaf calls at half the issue cost.
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  uint64_t& dd = reinterpret_cast<uint64_t&>(d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd) : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
}
__device__ __forceinline__ void ffma2s(float2& d, const float2& a, float s) { ffma2(d, a, make_float2(s, s)); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

}  // namespace evrp
