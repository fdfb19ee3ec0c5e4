// Spherically binned power spectrum (pmwd/spec_util.py:50-147): |f_k|^2 (or f_k conj(g_k)),
// optional sinc deconvolution, Hermitian multiplicities and np.digitize binning of one half
// spectrum in ONE pass, accumulated in float64.  A warp walks one (kx, ky) row; every lane keeps
// the sums of its current bin in registers (k grows along the row, so a lane changes bin only a
// handful of times) and flushes them to the CTA's shared-memory histogram, which is added to the
// global one at the end: a few thousand float64 atomics per CTA instead of one per mode.
#include "common.cuh"
#include "powspec.cuh"

namespace pmwd {

constexpr int PS_MAX_EDGES = 512;

struct PsParams {
  int nx, ny, nz, nzc;
  int nedges, right, has_g, has_deconv;
  float deconv;
  const float2* f;
  const float2* g;
  const double* edges;
  double* out;                       // [4][nedges + 1]: k*N, Re P*N, Im P*N, N
};

__device__ __forceinline__ void ps_flush(double* hist, int nb, int bin, double kN, double pr, double pi, double N) {
  if (bin < 0 || N == 0.0) return;
  atomicAdd(hist + bin, kN);
  atomicAdd(hist + nb + bin, pr);
  atomicAdd(hist + 2 * nb + bin, pi);
  atomicAdd(hist + 3 * nb + bin, N);
}

__global__ void __launch_bounds__(256) powspec_kernel(PsParams P) {
  extern __shared__ double ps_smem[];
  const int nb = P.nedges + 1;
  double* edges = ps_smem;                       // [nedges]
  double* hist = ps_smem + P.nedges;             // [4][nb]
  for (int t = threadIdx.x; t < P.nedges; t += blockDim.x) edges[t] = P.edges[t];
  for (int t = threadIdx.x; t < 4 * nb; t += blockDim.x) hist[t] = 0.0;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t nrows = (int64_t)P.nx * P.ny;
  for (int64_t row = warp; row < nrows; row += nwarps) {
    const int i = (int)(row / P.ny), j = (int)(row - (int64_t)i * P.ny);
    int cur = -1;
    double kN = 0, pr = 0, pi = 0, N = 0;
    for (int l = lane; l < P.nzc; l += 32) {
      const int64_t q = row * P.nzc + l;
      const float2 f = P.f[q];
      float2 g = make_float2(0.f, 0.f);
      if (P.has_g) g = P.g[q];
      const ps::Mode m = ps::mode(i, j, l, P.nx, P.ny, P.nz, f.x, f.y, P.has_g != 0, g.x, g.y, P.has_deconv != 0,
                                  P.deconv, edges, P.nedges, P.right != 0);
      if (m.bin != cur) {
        ps_flush(hist, nb, cur, kN, pr, pi, N);
        cur = m.bin;
        kN = pr = pi = N = 0;
      }
      kN += m.kN; pr += m.pr; pi += m.pi; N += m.N;
    }
    ps_flush(hist, nb, cur, kN, pr, pi, N);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 4 * nb; t += blockDim.x)
    if (hist[t] != 0.0) atomicAdd(P.out + t, hist[t]);
}

// out_k = w_k f_k (see ps::weight): the k-space half of powspec's VJP
__global__ void __launch_bounds__(256) powspec_weight_kernel(PsParams P, const double* wbin, float2* out) {
  extern __shared__ double ps_smem[];
  double* edges = ps_smem;                       // [nedges]
  double* wb = ps_smem + P.nedges;               // [nedges + 1]
  for (int t = threadIdx.x; t < P.nedges; t += blockDim.x) edges[t] = P.edges[t];
  for (int t = threadIdx.x; t <= P.nedges; t += blockDim.x) wb[t] = wbin[t];
  __syncthreads();
  const int64_t n = (int64_t)P.nx * P.ny * P.nzc;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(q % P.nzc);
    const int64_t row = q / P.nzc;
    const int j = (int)(row % P.ny), i = (int)(row / P.ny);
    const float w = ps::weight(i, j, l, P.nx, P.ny, P.nz, P.has_deconv != 0, P.deconv, edges, P.nedges,
                               P.right != 0, wb);
    const float2 f = P.f[q];
    out[q] = make_float2(f.x * w, f.y * w);
  }
}

}  // namespace pmwd

using namespace pmwd;

// VJP helper of the auto spectrum: out_c64 = w_k f_k with w_k = wbin[bin(k)] * prod_a sinc(k_a)^-deconv
// (wbin: device float64[nedges + 1], indexed like pmwd_powspec_bin's bins).  The field cotangent is
// twice the unnormalised C2R transform of out_c64.  out_c64 may alias f_c64.
extern "C" int pmwd_powspec_weight(void* stream, const int32_t* shape, const void* f_c64, int has_deconv,
                                   double deconv, const double* edges_dev, int nedges, int right,
                                   const double* wbin_dev, void* out_c64) {
  PMWD_REQUIRE(shape && f_c64 && edges_dev && wbin_dev && out_c64, "null buffer");
  PMWD_REQUIRE(shape[0] > 0 && shape[1] > 0 && shape[2] > 0, "bad shape");
  PMWD_REQUIRE(nedges > 0 && nedges <= PS_MAX_EDGES, "1 <= nedges <= 512");
  PsParams P;
  P.nx = shape[0]; P.ny = shape[1]; P.nz = shape[2]; P.nzc = shape[2] / 2 + 1;
  P.nedges = nedges; P.right = right; P.has_g = 0; P.has_deconv = has_deconv;
  P.deconv = (float)deconv;
  P.f = (const float2*)f_c64; P.g = nullptr; P.edges = edges_dev; P.out = nullptr;
  const int64_t n = (int64_t)P.nx * P.ny * P.nzc;
  const int64_t want = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(want < cap ? want : cap);
  const size_t smem = (size_t)(2 * nedges + 1) * sizeof(double);
  powspec_weight_kernel<<<grid, 256, smem, as_stream(stream)>>>(P, wbin_dev, (float2*)out_c64);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

// Accumulates (+=) into out[4][nedges + 1] (device float64: sum k N, sum Re P N, sum Im P N, sum N
// per np.digitize bin 0..nedges); the caller zeroes it, and may call again for further fields
// (spec_util.py:112-113 sums the leading axes).  f_c64 / g_c64: half spectra [nx][ny][nz/2+1] of
// rfftn (pm_util.py:236-289), g_c64 NULL for the auto spectrum; edges: device float64[nedges],
// ascending, in cycles per grid unit (spec_util.py:10-47).
extern "C" int pmwd_powspec_bin(void* stream, const int32_t* shape, const void* f_c64, const void* g_c64,
                                int has_deconv, double deconv, const double* edges_dev, int nedges, int right,
                                double* out_dev) {
  PMWD_REQUIRE(shape && f_c64 && edges_dev && out_dev, "null buffer");
  PMWD_REQUIRE(shape[0] > 0 && shape[1] > 0 && shape[2] > 0, "bad shape");
  PMWD_REQUIRE(nedges > 0 && nedges <= PS_MAX_EDGES, "1 <= nedges <= 512");
  PsParams P;
  P.nx = shape[0]; P.ny = shape[1]; P.nz = shape[2]; P.nzc = shape[2] / 2 + 1;
  P.nedges = nedges; P.right = right; P.has_g = g_c64 != nullptr; P.has_deconv = has_deconv;
  P.deconv = (float)deconv;
  P.f = (const float2*)f_c64; P.g = (const float2*)g_c64; P.edges = edges_dev; P.out = out_dev;
  const int64_t nrows = (int64_t)P.nx * P.ny;
  const int64_t want = (nrows + 7) / 8;                     // 8 warps per CTA
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(want < cap ? want : cap);
  const size_t smem = (size_t)(nedges + 4 * (nedges + 1)) * sizeof(double);
  powspec_kernel<<<grid, 256, smem, as_stream(stream)>>>(P);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}
