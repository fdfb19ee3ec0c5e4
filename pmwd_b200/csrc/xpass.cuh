// Shared declarations of the fused x-pass kernels (xpass.cu: radix-4 shared-memory Stockham for
// any supported length; xpass16.cu: register-resident radix-16 for nx in {256, 512, 1024}).
#pragma once
#include "common.cuh"

namespace pmwd {

struct XParams {
  int nx, ny_l, nzc;            // array extents
  int ny_g, nz_g;               // global (real-space) sizes of axes 1, 2
  int y0;                       // global index of local row 0
  double period;                // 2 pi / spacing
  float nyq, eps, scale;
  const float2* in[3];
  float2* out[3];
};

__device__ __forceinline__ float xkval(int i, int n, double period, bool last) {
  int f = last ? i : (i < (n + 1) / 2 ? i : i - n);
  return (float)(((double)f / (double)n) * period);
}
__device__ __forceinline__ bool xnyq(float k, float nyq, float eps) {
  return fabsf(__fsub_rn(fabsf(k), nyq)) <= eps;
}
// xpass16.cu
bool xpass16_supported(const XParams& P, bool adjoint);
int xpass16_launch(cudaStream_t st, const XParams& P, bool adjoint);

}  // namespace pmwd
