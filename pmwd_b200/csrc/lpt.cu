// Fused elementwise passes of the LPT initial conditions (pmwd/lpt.py:40-76, 190-208) and their VJPs.
//
//   pmwd_lpt_source2      : the 2LPT source L = sum_{i<j} (s_ii s_jj - s_ij^2) from the six strain fields in
//                           ONE pass, same terms in the same order as lpt.py:47-74 (m == n);
//   pmwd_lpt_displace     : disp_i = (disp0_i + D1 g1_i) + D2 g2_i, vel_i likewise with a^2 H D' (lpt.py:203-208),
//                           all orders and axes in ONE pass straight into the (N, 3) particle arrays;
//   *_vjp                 : their cotangents (the float64 sums are the cotangents of the growth factors).
// The reference leaves these to XLA fusion; a literal torch transcription launches ~40 elementwise kernels.
#include "common.cuh"

namespace pmwd {

__global__ void __launch_bounds__(256)
lpt_source2_kernel(int64_t n, const float* __restrict__ s00, const float* __restrict__ s11,
                   const float* __restrict__ s22, const float* __restrict__ s01, const float* __restrict__ s02,
                   const float* __restrict__ s12, float* __restrict__ L) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = s00[i], b = s11[i], c = s22[i], d = s01[i], e = s02[i], f = s12[i];
    float l = __fadd_rn(0.f, __fmul_rn(a, c));          // i = 0: j = 2, then j = 1 (lpt.py:50-54)
    l = __fadd_rn(l, __fmul_rn(a, b));
    l = __fadd_rn(l, __fmul_rn(b, c));                  // i = 1: j = 2
    l = __fsub_rn(l, __fmul_rn(d, d));                  // (0,1), (0,2), (1,2)  (lpt.py:66-74)
    l = __fsub_rn(l, __fmul_rn(e, e));
    l = __fsub_rn(l, __fmul_rn(f, f));
    L[i] = l;
  }
}

__global__ void __launch_bounds__(256)
lpt_source2_vjp_kernel(int64_t n, const float* __restrict__ s00, const float* __restrict__ s11,
                       const float* __restrict__ s22, const float* __restrict__ s01, const float* __restrict__ s02,
                       const float* __restrict__ s12, const float* __restrict__ Lc, float* __restrict__ c00,
                       float* __restrict__ c11, float* __restrict__ c22, float* __restrict__ c01,
                       float* __restrict__ c02, float* __restrict__ c12) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = s00[i], b = s11[i], c = s22[i], g = Lc[i];
    c00[i] = (c + b) * g;
    c11[i] = (a + c) * g;
    c22[i] = (a + b) * g;
    c01[i] = -2.f * s01[i] * g;
    c02[i] = -2.f * s02[i] * g;
    c12[i] = -2.f * s12[i] * g;
  }
}

struct Ptr3 { const float* p[3]; };
struct MPtr3 { float* p[3]; };

__global__ void __launch_bounds__(256)
lpt_displace_kernel(int64_t n, const float* __restrict__ disp0, const float* __restrict__ vel0, Ptr3 g1, Ptr3 g2,
                    float D1, float V1, float D2, float V2, float* __restrict__ disp, float* __restrict__ vel) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < 3 * n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = q / 3;
    const int a = (int)(q - 3 * p);
    float d = disp0[q], v = vel0[q];
    const float ga = g1.p[a][p];
    d = __fadd_rn(d, __fmul_rn(D1, ga));
    v = __fadd_rn(v, __fmul_rn(V1, ga));
    if (g2.p[0]) {
      const float gb = g2.p[a][p];
      d = __fadd_rn(d, __fmul_rn(D2, gb));
      v = __fadd_rn(v, __fmul_rn(V2, gb));
    }
    disp[q] = d;
    vel[q] = v;
  }
}

// g1c_a = D1 dc_a + V1 vc_a, g2c_a = D2 dc_a + V2 vc_a; sums[0..3] += (sum dc g1, sum vc g1, sum dc g2, sum vc g2)
__global__ void __launch_bounds__(256)
lpt_displace_vjp_kernel(int64_t n, const float* __restrict__ dc, const float* __restrict__ vc, Ptr3 g1, Ptr3 g2,
                        float D1, float V1, float D2, float V2, MPtr3 g1c, MPtr3 g2c, double* __restrict__ sums) {
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < 3 * n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = q / 3;
    const int a = (int)(q - 3 * p);
    const float d = dc[q], v = vc[q];
    g1c.p[a][p] = D1 * d + V1 * v;
    const float ga = g1.p[a][p];
    s[0] += (double)d * ga;
    s[1] += (double)v * ga;
    if (g2.p[0]) {
      g2c.p[a][p] = D2 * d + V2 * v;
      const float gb = g2.p[a][p];
      s[2] += (double)d * gb;
      s[3] += (double)v * gb;
    }
  }
  __shared__ double red[4][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double t = s[k];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) red[k][warp] = t;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, t);
  }
}

}  // namespace pmwd

using namespace pmwd;

extern "C" int pmwd_lpt_source2(void* stream, int64_t n, const float* const* s, float* L) {
  PMWD_REQUIRE(s && L && n >= 0, "null buffer");
  for (int k = 0; k < 6; ++k) PMWD_REQUIRE(s[k], "null strain field");
  if (n == 0) return PMWD_OK;
  lpt_source2_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(n, s[0], s[1], s[2], s[3], s[4], s[5], L);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

extern "C" int pmwd_lpt_source2_vjp(void* stream, int64_t n, const float* const* s, const float* L_cot,
                                    float* const* s_cot) {
  PMWD_REQUIRE(s && s_cot && L_cot && n >= 0, "null buffer");
  for (int k = 0; k < 6; ++k) PMWD_REQUIRE(s[k] && s_cot[k], "null strain field");
  if (n == 0) return PMWD_OK;
  lpt_source2_vjp_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
      n, s[0], s[1], s[2], s[3], s[4], s[5], L_cot, s_cot[0], s_cot[1], s_cot[2], s_cot[3], s_cot[4], s_cot[5]);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

extern "C" int pmwd_lpt_displace(void* stream, int64_t n, const float* disp0, const float* vel0,
                                 const float* const* g1, const float* const* g2, float D1, float V1, float D2,
                                 float V2, float* disp, float* vel) {
  PMWD_REQUIRE(disp0 && vel0 && g1 && disp && vel && n >= 0, "null buffer");
  Ptr3 a, b;
  for (int k = 0; k < 3; ++k) {
    PMWD_REQUIRE(g1[k] && (!g2 || g2[k]), "null gradient field");
    a.p[k] = g1[k];
    b.p[k] = g2 ? g2[k] : nullptr;
  }
  if (n == 0) return PMWD_OK;
  lpt_displace_kernel<<<grid_for(3 * n, 256, 8), 256, 0, as_stream(stream)>>>(n, disp0, vel0, a, b, D1, V1, D2, V2,
                                                                              disp, vel);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

// sums: device double[4], accumulated into (the caller zero-fills)
extern "C" int pmwd_lpt_displace_vjp(void* stream, int64_t n, const float* disp_cot, const float* vel_cot,
                                     const float* const* g1, const float* const* g2, float D1, float V1, float D2,
                                     float V2, float* const* g1_cot, float* const* g2_cot, double* sums) {
  PMWD_REQUIRE(disp_cot && vel_cot && g1 && g1_cot && sums && n >= 0, "null buffer");
  PMWD_REQUIRE((g2 == nullptr) == (g2_cot == nullptr), "g2 and g2_cot go together");
  Ptr3 a, b;
  MPtr3 ac, bc;
  for (int k = 0; k < 3; ++k) {
    PMWD_REQUIRE(g1[k] && g1_cot[k] && (!g2 || (g2[k] && g2_cot[k])), "null gradient field");
    a.p[k] = g1[k]; ac.p[k] = g1_cot[k];
    b.p[k] = g2 ? g2[k] : nullptr; bc.p[k] = g2 ? g2_cot[k] : nullptr;
  }
  if (n == 0) return PMWD_OK;
  lpt_displace_vjp_kernel<<<grid_for(3 * n, 256, 8), 256, 0, as_stream(stream)>>>(n, disp_cot, vel_cot, a, b, D1, V1,
                                                                                  D2, V2, ac, bc, sums);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}
