// Per-mode arithmetic of the power-spectrum estimator (pmwd/spec_util.py:50-147), shared by the
// CUDA kernel (powspec.cu) and the CPU emulation test (tests/host/powspec_emul.cc).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PMWD_PS_HD __host__ __device__ __forceinline__
#else
#define PMWD_PS_HD inline
#endif

namespace pmwd {
namespace ps {

// fftfreq(n)[i] (cycles per grid unit) in float64, cast to float32 (spec_util.py:117, dtype of P);
// last axis: rfftfreq
PMWD_PS_HD float freq(int i, int n, bool last) {
  const int f = last ? i : (i < (n + 1) / 2 ? i : i - n);
  return (float)((double)f / (double)n);
}

// np.sinc(x) ** -p in float32 (spec_util.py:119-121)
PMWD_PS_HD float sinc_pow(float x, float p) {
  const double pi = 3.141592653589793238462643383279502884;
  const float y = (float)(pi * (double)(x == 0.f ? 1.0e-20f : x));
  const float s = sinf(y) / y;
  return powf(s, -p);
}

// np.digitize(k, edges, right): number of edges e with e <= k (right = false) or e < k (right = true)
PMWD_PS_HD int digitize(double k, const double* edges, int nedges, bool right) {
  int lo = 0, hi = nedges;           // first index whose edge is > k (or >= k when right)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const bool below = right ? (edges[mid] < k) : (edges[mid] <= k);
    if (below) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct Mode {
  int bin;
  double kN, pr, pi, N;              // k * N, Re P * N, Im P * N, N
};

// one Fourier mode (i, j, l) of the half spectrum; g = nullptr-equivalent via has_g
PMWD_PS_HD Mode mode(int i, int j, int l, int nx, int ny, int nz, float fr, float fi, bool has_g, float gr,
                     float gi, bool has_deconv, float deconv, const double* edges, int nedges, bool right) {
  const float kx = freq(i, nx, false), ky = freq(j, ny, false), kz = freq(l, nz, true);
  const float k = sqrtf((kx * kx + ky * ky) + kz * kz);
  float pr, pi = 0.f;
  if (!has_g) {
    pr = fr * fr + fi * fi;                                   // f.real ** 2 + f.imag ** 2
  } else {
    pr = fr * gr + fi * gi;                                   // f * conj(g)
    pi = fi * gr - fr * gi;
  }
  if (has_deconv) {
    const float w[3] = {sinc_pow(kx, deconv), sinc_pow(ky, deconv), sinc_pow(kz, deconv)};
    for (int a = 0; a < 3; ++a) {
      pr = pr * w[a];
      pi = pi * w[a];
    }
  }
  // Hermitian multiplicity of the half spectrum (spec_util.py:123-126)
  const double N = (l == 0 || ((nz & 1) == 0 && l == nz / 2)) ? 1.0 : 2.0;
  Mode m;
  m.bin = digitize((double)k, edges, nedges, right);
  m.kN = (double)k * N;
  m.pr = (double)pr * N;
  m.pi = (double)pi * N;
  m.N = N;
  return m;
}

// VJP of the auto spectrum w.r.t. the field: with L = sum_b Pbar_b P_b and
// wbin[b] = Pbar_b * (spacing^3 / N_total) / N_b (0 outside the returned bins),
//   dL/df(x) = 2 * sum_{k in full spectrum} w_k f_k e^{+ikx},  w_k = wbin[bin(k)] * prod_a sinc(k_a)^-deconv,
// i.e. twice the unnormalised C2R transform of w_k f_k.  This returns w_k for one mode.
PMWD_PS_HD float weight(int i, int j, int l, int nx, int ny, int nz, bool has_deconv, float deconv,
                        const double* edges, int nedges, bool right, const double* wbin) {
  const float kx = freq(i, nx, false), ky = freq(j, ny, false), kz = freq(l, nz, true);
  const float k = sqrtf((kx * kx + ky * ky) + kz * kz);
  float w = (float)wbin[digitize((double)k, edges, nedges, right)];
  if (has_deconv) w = ((w * sinc_pow(kx, deconv)) * sinc_pow(ky, deconv)) * sinc_pow(kz, deconv);
  return w;
}

}  // namespace ps
}  // namespace pmwd
