// Leapfrog updates: pmwd/nbody.py:39-46 (drift), :70-77 (kick) and their adjoint companions
// :49-67 (drift_adj), :80-99 (kick_adj), fused so that one pass over the particle arrays does
// the half-kick and the drift (60 B/particle instead of 2 x 36 B), and in the adjoint also the
// cotangent updates and the two global dot products that scale the cosmology cotangents.
//
// Pure streaming, HBM-bound: float4 vectorised over the flattened (N*3) arrays, grid-stride,
// grid = 148 SMs x 8.  Arithmetic is mul-then-add in float32 exactly as the reference
// (`vel + acc * factor`), no FMA contraction.
#include "common.cuh"

namespace pmwd {

__device__ __forceinline__ float4 axpy4(float4 y, float4 x, float a) {
  y.x = __fadd_rn(y.x, __fmul_rn(x.x, a));
  y.y = __fadd_rn(y.y, __fmul_rn(x.y, a));
  y.z = __fadd_rn(y.z, __fmul_rn(x.z, a));
  y.w = __fadd_rn(y.w, __fmul_rn(x.w, a));
  return y;
}
__device__ __forceinline__ float4 axmy4(float4 y, float4 x, float a) {  // y - x*a
  y.x = __fsub_rn(y.x, __fmul_rn(x.x, a));
  y.y = __fsub_rn(y.y, __fmul_rn(x.y, a));
  y.z = __fsub_rn(y.z, __fmul_rn(x.z, a));
  y.w = __fsub_rn(y.w, __fmul_rn(x.w, a));
  return y;
}

template <bool KICK, bool DRIFT>
__global__ void __launch_bounds__(256)
kick_drift_kernel(int64_t n, float* __restrict__ disp, float* __restrict__ vel,
                  const float* __restrict__ acc, float K, float D) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  for (int64_t i = tid; i < n4; i += stride) {
    float4 v = reinterpret_cast<float4*>(vel)[i];
    if (KICK) {
      v = axpy4(v, reinterpret_cast<const float4*>(acc)[i], K);
      reinterpret_cast<float4*>(vel)[i] = v;
    }
    if (DRIFT) {
      float4 x = reinterpret_cast<float4*>(disp)[i];
      reinterpret_cast<float4*>(disp)[i] = axpy4(x, v, D);
    }
  }
  for (int64_t i = (n4 << 2) + tid; i < n; i += stride) {
    float v = vel[i];
    if (KICK) { v = __fadd_rn(v, __fmul_rn(acc[i], K)); vel[i] = v; }
    if (DRIFT) disp[i] = __fadd_rn(disp[i], __fmul_rn(v, D));
  }
}

__device__ __forceinline__ double dot4(float4 a, float4 b) {
  return (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
}

// PRE: a kick_adj with factor K0 (the previous step's trailing half-kick, same acc / alpha) is applied
// first in the same pass; its sum(pi . acc) equals this kick's and is added to sums_pre[0] as well.
template <bool KICK, bool DRIFT, bool PRE = false>
__global__ void __launch_bounds__(256)
kick_drift_adj_kernel(int64_t n, float* __restrict__ disp, float* __restrict__ vel,
                      const float* __restrict__ acc, float* __restrict__ xi,
                      float* __restrict__ pi, const float* __restrict__ alpha, float K, float D,
                      double* __restrict__ sums, float K0 = 0.f, double* __restrict__ sums_pre = nullptr) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  double s_pa = 0.0, s_xv = 0.0;   // sum(pi*acc), sum(xi*vel) accumulated in float64
  for (int64_t i = tid; i < n4; i += stride) {
    float4 v = reinterpret_cast<float4*>(vel)[i];
    float4 x = reinterpret_cast<float4*>(xi)[i];
    float4 p = reinterpret_cast<float4*>(pi)[i];
    if (KICK) {
      float4 a = reinterpret_cast<const float4*>(acc)[i];
      const float4 al = reinterpret_cast<const float4*>(alpha)[i];
      if (PRE) {
        v = axpy4(v, a, K0);
        x = axmy4(x, al, K0);
      }
      v = axpy4(v, a, K);                                              // nbody.py:87
      x = axmy4(x, al, K);                                             // nbody.py:91
      s_pa += dot4(p, a);                                              // nbody.py:95
      reinterpret_cast<float4*>(vel)[i] = v;
      reinterpret_cast<float4*>(xi)[i] = x;
    }
    if (DRIFT) {
      float4 d = reinterpret_cast<float4*>(disp)[i];
      reinterpret_cast<float4*>(disp)[i] = axpy4(d, v, D);             // nbody.py:56
      p = axmy4(p, x, D);                                              // nbody.py:60
      s_xv += dot4(x, v);                                              // nbody.py:64
      reinterpret_cast<float4*>(pi)[i] = p;
    }
  }
  for (int64_t i = (n4 << 2) + tid; i < n; i += stride) {
    float v = vel[i], x = xi[i], p = pi[i];
    if (KICK) {
      float a = acc[i];
      if (PRE) {
        v = __fadd_rn(v, __fmul_rn(a, K0));
        x = __fsub_rn(x, __fmul_rn(alpha[i], K0));
      }
      v = __fadd_rn(v, __fmul_rn(a, K));
      x = __fsub_rn(x, __fmul_rn(alpha[i], K));
      s_pa += (double)p * a;
      vel[i] = v; xi[i] = x;
    }
    if (DRIFT) {
      disp[i] = __fadd_rn(disp[i], __fmul_rn(v, D));
      p = __fsub_rn(p, __fmul_rn(x, D));
      s_xv += (double)x * v;
      pi[i] = p;
    }
  }
  // block reduction in float64, then one atomic per block per sum
  __shared__ double red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_pa += __shfl_xor_sync(0xffffffffu, s_pa, o);
    s_xv += __shfl_xor_sync(0xffffffffu, s_xv, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = s_pa; red[1][w] = s_xv; }
  __syncthreads();
  if (w == 0) {
    double a = l < (blockDim.x >> 5) ? red[0][l] : 0.0;
    double b = l < (blockDim.x >> 5) ? red[1][l] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (l == 0) {
      if (KICK) atomicAdd(sums + 0, a);
      if (PRE) atomicAdd(sums_pre + 0, a);
      if (DRIFT) atomicAdd(sums + 1, b);
    }
  }
}

}  // namespace pmwd

using namespace pmwd;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int pmwd_kick_drift(void* stream, int64_t n, float* disp, float* vel, const float* acc,
                               float K, float D, int do_kick, int do_drift) {
  PMWD_REQUIRE(n >= 0, "negative length");
  PMWD_REQUIRE(vel != nullptr, "null vel");
  PMWD_REQUIRE(!do_kick || acc, "kick needs acc");
  PMWD_REQUIRE(!do_drift || disp, "drift needs disp");
  PMWD_REQUIRE(aligned16(disp) && aligned16(vel) && aligned16(acc), "arrays must be 16-byte aligned");
  if (n == 0 || (!do_kick && !do_drift)) return PMWD_OK;
  cudaStream_t st = as_stream(stream);
  StageTimer timer(ST_KICK_DRIFT, st);
  int grid = grid_for((n + 3) / 4, 256, 8);
  if (do_kick && do_drift) kick_drift_kernel<true, true><<<grid, 256, 0, st>>>(n, disp, vel, acc, K, D);
  else if (do_kick) kick_drift_kernel<true, false><<<grid, 256, 0, st>>>(n, disp, vel, acc, K, D);
  else kick_drift_kernel<false, true><<<grid, 256, 0, st>>>(n, disp, vel, acc, K, D);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

extern "C" int pmwd_kick_drift_adj(void* stream, int64_t n, float* disp, float* vel,
                                   const float* acc, float* xi, float* pi, const float* alpha,
                                   float K, float D, int do_kick, int do_drift, double* sums) {
  PMWD_REQUIRE(n >= 0, "negative length");
  PMWD_REQUIRE(vel && xi && pi && sums, "null buffer");
  PMWD_REQUIRE(!do_kick || (acc && alpha), "kick_adj needs acc and alpha");
  PMWD_REQUIRE(!do_drift || disp, "drift_adj needs disp");
  PMWD_REQUIRE(aligned16(disp) && aligned16(vel) && aligned16(acc) && aligned16(xi) &&
               aligned16(pi) && aligned16(alpha), "arrays must be 16-byte aligned");
  if (n == 0 || (!do_kick && !do_drift)) return PMWD_OK;
  cudaStream_t st = as_stream(stream);
  StageTimer timer(ST_KICK_DRIFT_ADJ, st);
  int grid = grid_for((n + 3) / 4, 256, 8);
  if (do_kick && do_drift)
    kick_drift_adj_kernel<true, true><<<grid, 256, 0, st>>>(n, disp, vel, acc, xi, pi, alpha, K, D, sums);
  else if (do_kick)
    kick_drift_adj_kernel<true, false><<<grid, 256, 0, st>>>(n, disp, vel, acc, xi, pi, alpha, K, D, sums);
  else
    kick_drift_adj_kernel<false, true><<<grid, 256, 0, st>>>(n, disp, vel, acc, xi, pi, alpha, K, D, sums);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

// Two consecutive adjoint updates in one pass: kick_adj(K0) [the trailing half-kick of one step], then
// kick_adj(K) + drift_adj(D) [the leading half-kick and drift of the next step, pmwd/nbody.py:143-162];
// same float32 operation sequence as the two separate calls.  sums_pre[0] += sum(pi . acc) (of the first
// kick), sums[0] += the same sum (second kick), sums[1] += sum(xi . vel).
extern "C" int pmwd_kick_kick_drift_adj(void* stream, int64_t n, float* disp, float* vel, const float* acc,
                                        float* xi, float* pi, const float* alpha, float K0, float K, float D,
                                        double* sums_pre, double* sums) {
  PMWD_REQUIRE(n >= 0, "negative length");
  PMWD_REQUIRE(disp && vel && xi && pi && acc && alpha && sums && sums_pre, "null buffer");
  PMWD_REQUIRE(aligned16(disp) && aligned16(vel) && aligned16(acc) && aligned16(xi) &&
               aligned16(pi) && aligned16(alpha), "arrays must be 16-byte aligned");
  if (n == 0) return PMWD_OK;
  cudaStream_t st = as_stream(stream);
  StageTimer timer(ST_KICK_DRIFT_ADJ, st);
  int grid = grid_for((n + 3) / 4, 256, 8);
  kick_drift_adj_kernel<true, true, true><<<grid, 256, 0, st>>>(n, disp, vel, acc, xi, pi, alpha, K, D, sums, K0,
                                                               sums_pre);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}
