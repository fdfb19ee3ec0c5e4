// Fused k-space kernels between the R2C and C2R transforms.
//
// Replaces the reference's chain of separate XLA elementwise passes over the half-spectrum
// (pmwd/gravity.py:12-14 laplace, :38-42 neg_grad once per axis, pmwd/lpt.py:22-32 strain) by
// single passes: the forward force spectrum reads rho_k once and writes the three gradient
// spectra (16 B/cell instead of ~48), the adjoint reads three and writes one.
//
// HBM-bound, one complex64 per thread-iteration.  Wavenumbers follow fftfreq
// (pmwd/pm_util.py:159-199): computed in float64 (index / n * 2 pi / spacing), cast to float32;
// the per-axis tables for the two fastest axes live in shared memory so that no float64
// division is issued per element.  k^2 is summed in float32 in axis order (gravity.py:12).
#include "common.cuh"

namespace pmwd {

enum { KS_LAPLACE = 0, KS_NEG_GRAD = 1, KS_FORCE = 2, KS_FORCE_ADJ = 3, KS_STRAIN = 4 };

struct KsParams {
  int rank;
  int n[3];        // real-space shape padded on the left with 1s
  int nc;          // n[2]/2 + 1
  double period;   // 2 pi / spacing
  float nyq, eps;  // pi / spacing and nyq * eps(float32)   (gravity.py:38-39)
  int n1_global;   // full length of axis 1 (== n[1] unless the spectrum is a y-slab)
  int i1_off;      // global index of local row 0 along axis 1
  int ax_i, ax_j;  // padded axis ids for NEG_GRAD / STRAIN
  float scale;
  const float2* in[3];
  float2* out[3];
};

__device__ __forceinline__ float kval(int i, int n, double period, bool last) {
  // fftfreq(n)[i] = (i or i-n)/n ; rfftfreq(n)[i] = i/n   (float64, then cast)
  int f = last ? i : (i < (n + 1) / 2 ? i : i - n);
  return (float)(((double)f / (double)n) * period);
}

__device__ __forceinline__ bool is_nyq(float k, float nyq, float eps) {
  return fabsf(__fsub_rn(fabsf(k), nyq)) <= eps;
}

template <int MODE>
__global__ void __launch_bounds__(256) kspace_kernel(KsParams P) {
  extern __shared__ float tab[];
  float* k2tab = tab;            // axis 2 (last, rfftfreq), nc entries
  float* k1tab = tab + P.nc;     // axis 1, n[1] entries
  for (int i = threadIdx.x; i < P.nc; i += blockDim.x) k2tab[i] = kval(i, P.n[2], P.period, true);
  for (int i = threadIdx.x; i < P.n[1]; i += blockDim.x)
    k1tab[i] = kval(i + P.i1_off, P.n1_global, P.period, false);
  __syncthreads();

  const int first = 3 - P.rank;  // first real axis among the padded three
  const int64_t rows = (int64_t)P.n[0] * P.n[1];
  // one warp per (i0, i1) row; lanes stride the contiguous last axis
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < rows; row += nwarp) {
    const int i0 = (int)(row / P.n[1]);
    const int i1 = (int)(row - (int64_t)i0 * P.n[1]);
    const float k0 = first <= 0 ? kval(i0, P.n[0], P.period, false) : 0.f;
    const float k1 = first <= 1 ? k1tab[i1] : 0.f;
    const int64_t base = row * P.nc;
    // 4 independent loads in flight per lane before any dependent work (latency-bound otherwise)
    constexpr int U = 4;
    constexpr int NIN = MODE == KS_FORCE_ADJ ? 3 : 1;
    for (int i2b = lane; i2b < P.nc; i2b += 32 * U) {
      float2 in[U][NIN];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i2 = i2b + 32 * u;
#pragma unroll
        for (int q = 0; q < NIN; ++q) {
          in[u][q] = make_float2(0.f, 0.f);
          if (i2 < P.nc && (NIN == 1 || q + first < 3)) in[u][q] = __ldcs(P.in[q] + base + i2);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i2 = i2b + 32 * u;
        if (i2 >= P.nc) break;
        const float k2 = k2tab[i2];
        const float kv[3] = {k0, k1, k2};
        // k^2 = ((0 + k_a^2) + k_b^2) + k_c^2 over the real axes in order
        float ksq = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
          if (a >= first) ksq = __fadd_rn(ksq, __fmul_rn(kv[a], kv[a]));
        const int64_t e = base + i2;
        const float2 s = in[u][0];

        if (MODE == KS_LAPLACE) {
          float2 o = make_float2(0.f, 0.f);
          if (ksq != 0.f) o = make_float2(__fdiv_rn(-s.x, ksq), __fdiv_rn(-s.y, ksq));
          __stcs(P.out[0] + e, o);
        } else if (MODE == KS_NEG_GRAD) {
          const float k = P.ax_i == 0 ? k0 : (P.ax_i == 1 ? k1 : k2);
          float2 o = make_float2(0.f, 0.f);
          if (!is_nyq(k, P.nyq, P.eps)) o = make_float2(__fmul_rn(k, s.y), -__fmul_rn(k, s.x));
          __stcs(P.out[0] + e, o);
        } else if (MODE == KS_FORCE) {
          float2 pot = make_float2(0.f, 0.f);
          if (ksq != 0.f)
            pot = make_float2(__fdiv_rn(-__fmul_rn(P.scale, s.x), ksq),
                              __fdiv_rn(-__fmul_rn(P.scale, s.y), ksq));
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            if (a < first) continue;
            const float k = kv[a];
            float2 o = make_float2(0.f, 0.f);
            if (!is_nyq(k, P.nyq, P.eps)) o = make_float2(__fmul_rn(k, pot.y), -__fmul_rn(k, pot.x));
            __stcs(P.out[a - first] + e, o);
          }
        } else if (MODE == KS_FORCE_ADJ) {
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            if (a < first) continue;
            const float k = kv[a];
            if (ksq != 0.f && !is_nyq(k, P.nyq, P.eps)) {
              const float2 v = in[u][a - first];
              float2 phi = make_float2(__fdiv_rn(-v.x, ksq), __fdiv_rn(-v.y, ksq));
              // -( -i k phi ) = +i k phi = (-k phi.y, k phi.x)
              acc.x = __fsub_rn(acc.x, __fmul_rn(k, phi.y));
              acc.y = __fadd_rn(acc.y, __fmul_rn(k, phi.x));
            }
          }
          __stcs(P.out[0] + e, make_float2(__fmul_rn(P.scale, acc.x), __fmul_rn(P.scale, acc.y)));
        } else if (MODE == KS_STRAIN) {
          float ki = P.ax_i == 0 ? k0 : (P.ax_i == 1 ? k1 : k2);
          float kj = P.ax_j == 0 ? k0 : (P.ax_j == 1 ? k1 : k2);
          if (P.ax_i != P.ax_j) {
            if (is_nyq(ki, P.nyq, P.eps)) ki = 0.f;
            if (is_nyq(kj, P.nyq, P.eps)) kj = 0.f;
          }
          const float m = __fmul_rn(-ki, kj);
          __stcs(P.out[0] + e, make_float2(__fmul_rn(m, s.x), __fmul_rn(m, s.y)));
        }
      }
    }
  }
}

static int ks_setup(KsParams* P, int rank, const int32_t* shape, double spacing) {
  PMWD_REQUIRE(rank >= 1 && rank <= 3, "rank must be 1, 2 or 3");
  PMWD_REQUIRE(shape != nullptr, "null shape");
  PMWD_REQUIRE(spacing > 0, "spacing must be positive");
  memset(P, 0, sizeof(*P));
  P->rank = rank;
  for (int a = 0; a < 3; ++a) P->n[a] = 1;
  for (int a = 0; a < rank; ++a) {
    PMWD_REQUIRE(shape[a] > 0, "non-positive shape");
    P->n[3 - rank + a] = shape[a];
  }
  P->nc = P->n[2] / 2 + 1;
  P->n1_global = P->n[1];
  P->i1_off = 0;
  const double pi = 3.141592653589793238462643383279502884;
  P->period = 2.0 * pi / spacing;
  double nyq = pi / spacing;
  P->nyq = (float)nyq;
  // eps = nyquist * finfo(float32).eps evaluated in float64 then used against float32 values
  P->eps = (float)(nyq * 1.1920928955078125e-07);
  P->scale = 1.f;
  return PMWD_OK;
}

template <int MODE>
static int ks_launch(cudaStream_t st, const KsParams& P) {
  const int block = 256;
  int64_t rows = (int64_t)P.n[0] * P.n[1];
  int64_t warps_needed = rows;
  int64_t blocks = (warps_needed * 32 + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * 8;
  int grid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
  size_t smem = (size_t)(P.nc + P.n[1]) * sizeof(float);
  PMWD_REQUIRE(smem <= 48 * 1024, "mesh axis too long for the k tables");
  kspace_kernel<MODE><<<grid, block, smem, st>>>(P);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

}  // namespace pmwd

using namespace pmwd;

extern "C" int pmwd_laplace(void* stream, int rank, const int32_t* shape, double spacing,
                            const void* src, void* pot) {
  KsParams P;
  int rc = ks_setup(&P, rank, shape, spacing);
  if (rc) return rc;
  PMWD_REQUIRE(src && pot, "null buffer");
  P.in[0] = (const float2*)src;
  P.out[0] = (float2*)pot;
  return ks_launch<KS_LAPLACE>(as_stream(stream), P);
}

extern "C" int pmwd_neg_grad(void* stream, int rank, const int32_t* shape, double spacing,
                             int axis, const void* pot, void* out) {
  KsParams P;
  int rc = ks_setup(&P, rank, shape, spacing);
  if (rc) return rc;
  PMWD_REQUIRE(pot && out, "null buffer");
  PMWD_REQUIRE(axis >= 0 && axis < rank, "axis out of range");
  P.ax_i = 3 - rank + axis;
  P.in[0] = (const float2*)pot;
  P.out[0] = (float2*)out;
  return ks_launch<KS_NEG_GRAD>(as_stream(stream), P);
}

extern "C" int pmwd_kspace_force(void* stream, int rank, const int32_t* shape, double spacing,
                                 float scale, const void* rho, void* const* g) {
  KsParams P;
  int rc = ks_setup(&P, rank, shape, spacing);
  if (rc) return rc;
  PMWD_REQUIRE(rho && g, "null buffer");
  P.scale = scale;
  P.in[0] = (const float2*)rho;
  for (int a = 0; a < rank; ++a) {
    PMWD_REQUIRE(g[a] != nullptr, "null output spectrum");
    P.out[a] = (float2*)g[a];
  }
  return ks_launch<KS_FORCE>(as_stream(stream), P);
}

extern "C" int pmwd_kspace_force_adj(void* stream, int rank, const int32_t* shape, double spacing,
                                     float scale, const void* const* v, void* out) {
  KsParams P;
  int rc = ks_setup(&P, rank, shape, spacing);
  if (rc) return rc;
  PMWD_REQUIRE(v && out, "null buffer");
  P.scale = scale;
  for (int a = 0; a < rank; ++a) {
    PMWD_REQUIRE(v[a] != nullptr, "null input spectrum");
    P.in[a] = (const float2*)v[a];
  }
  P.out[0] = (float2*)out;
  return ks_launch<KS_FORCE_ADJ>(as_stream(stream), P);
}

// Slab (multi-GPU) variants: the spectrum is the transposed slab [n0][ny_local][n2/2+1] holding
// global rows y0 .. y0+ny_local-1 of axis 1 (after the distributed FFT's all-to-all).
extern "C" int pmwd_kspace_force_slab(void* stream, const int32_t* shape, int y0, int ny_local,
                                      double spacing, float scale, const void* rho, void* const* g) {
  KsParams P;
  int rc = ks_setup(&P, 3, shape, spacing);
  if (rc) return rc;
  PMWD_REQUIRE(rho && g && g[0] && g[1] && g[2], "null buffer");
  PMWD_REQUIRE(y0 >= 0 && ny_local > 0 && y0 + ny_local <= shape[1], "bad y slab");
  P.n[1] = ny_local;
  P.i1_off = y0;
  P.scale = scale;
  P.in[0] = (const float2*)rho;
  for (int a = 0; a < 3; ++a) P.out[a] = (float2*)g[a];
  return ks_launch<KS_FORCE>(as_stream(stream), P);
}

extern "C" int pmwd_kspace_force_adj_slab(void* stream, const int32_t* shape, int y0, int ny_local,
                                          double spacing, float scale, const void* const* v,
                                          void* out) {
  KsParams P;
  int rc = ks_setup(&P, 3, shape, spacing);
  if (rc) return rc;
  PMWD_REQUIRE(v && out && v[0] && v[1] && v[2], "null buffer");
  PMWD_REQUIRE(y0 >= 0 && ny_local > 0 && y0 + ny_local <= shape[1], "bad y slab");
  P.n[1] = ny_local;
  P.i1_off = y0;
  P.scale = scale;
  for (int a = 0; a < 3; ++a) P.in[a] = (const float2*)v[a];
  P.out[0] = (float2*)out;
  return ks_launch<KS_FORCE_ADJ>(as_stream(stream), P);
}

extern "C" int pmwd_strain(void* stream, int rank, const int32_t* shape, double spacing, int i,
                           int j, const void* pot, void* out) {
  KsParams P;
  int rc = ks_setup(&P, rank, shape, spacing);
  if (rc) return rc;
  PMWD_REQUIRE(pot && out, "null buffer");
  PMWD_REQUIRE(i >= 0 && i < rank && j >= 0 && j < rank, "axis out of range");
  P.ax_i = 3 - rank + i;
  P.ax_j = 3 - rank + j;
  P.in[0] = (const float2*)pot;
  P.out[0] = (float2*)out;
  return ks_launch<KS_STRAIN>(as_stream(stream), P);
}
