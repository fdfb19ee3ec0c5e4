// Fused x-pass of the FFT Poisson solve.
//
// The 3-D transforms of gravity() (pmwd/gravity.py:56,64) are split as  (y,z) 2-D FFTs (cuFFT,
// batched over x planes)  +  1-D FFTs along x done HERE, in shared memory, fused with the
// k-space algebra of laplace / neg_grad (gravity.py:9-16, 37-44):
//
//   forward :  S = FFT_x(rho2d);  pot = -(scale S)/k^2;
//              G_x = IFFT_x(-i kx pot);  P = IFFT_x(pot);  G_y = -i ky P;  G_z = -i kz P
//              (ky, kz are constant along x, so two inverse transforms serve three outputs)
//   adjoint :  W = i ky V_y + i kz V_z;  S = (FFT_x(W) + i kx FFT_x(V_x)) * (-scale / k^2);
//              rho_cot2d = IFFT_x(S)                    [A_i^T = -A_i, pmwd/nbody.py:111-116]
//
// One pass reads the spectrum once and writes the outputs once (16 N_m bytes forward) instead
// of 4 strided cuFFT passes + the k-space kernel (5 x 16 N_m bytes): HBM-bound work whose
// transforms stay on chip.  Layout: data[nx][ny_l][nzc] (x slowest); a CTA owns a tile of
// T = 8 adjacent kz columns for one y (64-byte segments per x row, two full sectors).
// The FFT is a one-buffer Stockham autosort (radix 4 (+2)), 1024 threads per CTA,
// twiddles from a shared-memory table built in float64.  Works on the slab-decomposed
// layout too (ny_l rows starting at global row y0).
#include <cuda_pipeline.h>
#include <stdlib.h>

#include "common.cuh"
#include "xpass.cuh"

namespace pmwd {


#ifndef PMWD_XTHREADS
#define PMWD_XTHREADS 1024
#endif
#ifndef PMWD_XT
#define PMWD_XT 8     // measured on B200 at 1024^3: T=8 (64-byte rows) 11.5 ms, T=4 (32-byte rows) 23.4 ms
#endif
template <int NX>
struct XCfg {
  // T kz-columns per tile (T*8-byte row segments).  T=8 runs 1024 threads, one CTA per SM;
  // T=4 runs 512 threads, two CTAs per SM (slower: 32-byte rows waste L2/DRAM requests).
  static constexpr int T = PMWD_XT;
  static constexpr int BIG = PMWD_XTHREADS;
  static constexpr int CTAS = BIG == 1024 ? 1 : 2;
  static constexpr int THREADS = NX * T / 2 >= BIG ? BIG : NX * T / 2;
  static constexpr int EPT = NX * T / THREADS;
  static constexpr int XSTEP = THREADS / T;
  static constexpr int LOG2 = NX == 64 ? 6 : NX == 128 ? 7 : NX == 256 ? 8 : NX == 512 ? 9 : NX == 1024 ? 10 : 11;
  // q = -i kx pot is parked in registers across the P transform, except for the longest
  // columns where that would spill: there it is parked in the G_x output array itself
  // (written, then re-read -- an L2 hit -- transformed and overwritten)
  static constexpr bool QGLOBAL = (NX * T / THREADS) >= 16;
  // two tile buffers (cp.async prefetch of the next tile) when they fit in shared memory
  static constexpr bool DBUF = 2 * NX * T * 8 + NX * 12 <= 200 * 1024;
  static constexpr int N4 = LOG2 / 2;
  static constexpr bool HAS2 = (LOG2 & 1) != 0;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// One Stockham autosort stage of radix R over the tile buffer buf[x][XT] (p = product of the
// radices already applied).  Every thread owns butterflies b = tid + THREADS*m with column
// c = b % XT and index i = b / XT; all reads complete before any write (one buffer).
template <int NX, int R, bool INV>
__device__ __forceinline__ void fft_stage(float2* buf, const float2* tw, const int p) {
  constexpr int THREADS = XCfg<NX>::THREADS;
  constexpr int XT = XCfg<NX>::T;
  constexpr int t = NX / R;                     // butterflies per column
  constexpr int nb = XT * t;
  constexpr int MAXB = (nb + THREADS - 1) / THREADS;
  float2 v[MAXB][R];                            // the whole tile lives in registers across the barrier
  const int tstep = NX / (p * R);               // twiddle table stride: angle = 2 pi r k / (p R)
#pragma unroll
  for (int m = 0; m < MAXB; ++m) {
    const int b = threadIdx.x + THREADS * m;
    if (nb % THREADS == 0 || b < nb) {
      const int c = b % XT, i = b / XT;
      const int k = i & (p - 1);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float2 x = buf[(i + r * t) * XT + c];
        if (r > 0 && p > 1) {
          float2 w = tw[(r * k * tstep) & (NX - 1)];
          if (INV) w.y = -w.y;
          x = cmul(x, w);
        }
        v[m][r] = x;
      }
      if (R == 4) {
        const float2 a0 = make_float2(v[m][0].x + v[m][2].x, v[m][0].y + v[m][2].y);
        const float2 a1 = make_float2(v[m][0].x - v[m][2].x, v[m][0].y - v[m][2].y);
        const float2 a2 = make_float2(v[m][1].x + v[m][R - 1].x, v[m][1].y + v[m][R - 1].y);
        const float2 d = make_float2(v[m][1].x - v[m][R - 1].x, v[m][1].y - v[m][R - 1].y);
        // forward: multiply by -i ; inverse: by +i
        const float2 a3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
        v[m][0] = make_float2(a0.x + a2.x, a0.y + a2.y);
        v[m][1] = make_float2(a1.x + a3.x, a1.y + a3.y);
        v[m][R - 2] = make_float2(a0.x - a2.x, a0.y - a2.y);
        v[m][R - 1] = make_float2(a1.x - a3.x, a1.y - a3.y);
      } else {
        const float2 s0 = v[m][0], s1 = v[m][1];
        v[m][0] = make_float2(s0.x + s1.x, s0.y + s1.y);
        v[m][1] = make_float2(s0.x - s1.x, s0.y - s1.y);
      }
    }
  }
  __syncthreads();                              // all reads of this stage done
#pragma unroll
  for (int m = 0; m < MAXB; ++m) {
    const int b = threadIdx.x + THREADS * m;
    if (nb % THREADS == 0 || b < nb) {
      const int c = b % XT, i = b / XT;
      const int k = i & (p - 1);
      const int dst = ((i - k) * R + k) * XT + c;
#pragma unroll
      for (int r = 0; r < R; ++r) buf[dst + r * p * XT] = v[m][r];
    }
  }
  __syncthreads();
}

template <int NX, bool INV>
__device__ __forceinline__ void fft_tile(float2* buf, const float2* tw) {
  int p = 1;
#pragma unroll
  for (int s = 0; s < XCfg<NX>::N4; ++s) {
    fft_stage<NX, 4, INV>(buf, tw, p);
    p *= 4;
  }
  if (XCfg<NX>::HAS2) fft_stage<NX, 2, INV>(buf, tw, p);
}

template <int NX>
__device__ __forceinline__ void build_tables(const XParams& P, float2* tw, float* kx) {
  for (int n = threadIdx.x; n < NX; n += XCfg<NX>::THREADS) {
    double s, c;
    sincospi(-2.0 * (double)n / (double)NX, &s, &c);
    tw[n] = make_float2((float)c, (float)s);
    kx[n] = xkval(n, NX, P.period, false);
  }
}

// -------------------------------------------------------------------------- forward force
// Tile input is double-buffered: the next tile's column data is fetched with cp.async (LDGSTS)
// into the second buffer while the current tile goes through its three transforms.
template <int NX>
__device__ __forceinline__ void prefetch_tile(const XParams& P, float2* dst, int64_t tile, int ztiles,
                                              int64_t plane) {
  constexpr int EPT = XCfg<NX>::EPT;
  constexpr int XT = XCfg<NX>::T;
  constexpr int XSTEP = XCfg<NX>::XSTEP;
  const int c = threadIdx.x % XT;
  const int x0 = threadIdx.x / XT;
  const int iy = (int)(tile / ztiles);
  const int kz0 = (int)(tile - (int64_t)iy * ztiles) * XT;
  const bool live = kz0 + c < P.nzc;
  const int64_t col = (int64_t)iy * P.nzc + kz0 + c;
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int x = x0 + XSTEP * e;
    if (live) __pipeline_memcpy_async(dst + x * XT + c, P.in[0] + (int64_t)x * plane + col, sizeof(float2));
    else dst[x * XT + c] = make_float2(0.f, 0.f);
  }
  __pipeline_commit();
}

template <int NX>
__global__ void __launch_bounds__(XCfg<NX>::THREADS, XCfg<NX>::CTAS) xfused_force_kernel(XParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int XT = XCfg<NX>::T;
  float2* buf = reinterpret_cast<float2*>(smem_raw);           // [NX][T]   current tile
  constexpr bool DB = XCfg<NX>::DBUF;
  float2* nxt = DB ? buf + NX * XT : buf;                       // [NX][T]   next tile (prefetch)
  float2* tw = nxt + NX * XT;                                   // [NX]
  float* kx = reinterpret_cast<float*>(tw + NX);                // [NX]
  build_tables<NX>(P, tw, kx);

  constexpr int EPT = XCfg<NX>::EPT;                            // elements per thread
  const int ztiles = (P.nzc + XT - 1) / XT;
  const int64_t ntiles = (int64_t)P.ny_l * ztiles;
  const int c = threadIdx.x % XT;
  const int x0 = threadIdx.x / XT;
  constexpr int XSTEP = XCfg<NX>::XSTEP;
  const int64_t plane = (int64_t)P.ny_l * P.nzc;

  __syncthreads();
  if (DB && (int64_t)blockIdx.x < ntiles) prefetch_tile<NX>(P, buf, blockIdx.x, ztiles, plane);

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int iy = (int)(tile / ztiles);
    const int kz0 = (int)(tile - (int64_t)iy * ztiles) * XT;
    const bool live = kz0 + c < P.nzc;
    const int64_t col = (int64_t)iy * P.nzc + kz0 + c;
    const float ky = xkval(iy + P.y0, P.ny_g, P.period, false);
    const float kz = xkval(kz0 + c, P.nz_g, P.period, true);

    // ---- this tile's data has been requested one iteration ago: wait, then start the next one
    if (!DB) {
      __syncthreads();                                           // previous tile's stores from buf done
      prefetch_tile<NX>(P, buf, tile, ztiles, plane);
    }
    __pipeline_wait_prior(0);
    __syncthreads();
    if (DB && tile + gridDim.x < ntiles) prefetch_tile<NX>(P, nxt, tile + gridDim.x, ztiles, plane);
    fft_tile<NX, false>(buf, tw);

    // ---- pot = -(scale S)/k^2 in place; q = -i kx pot kept in registers
    float2 q[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int x = x0 + XSTEP * e;
      const float k0 = kx[x];
      float ksq = __fadd_rn(__fadd_rn(__fmul_rn(k0, k0), __fmul_rn(ky, ky)), __fmul_rn(kz, kz));
      float2 s = buf[x * XT + c];
      float2 pot = make_float2(0.f, 0.f);
      if (ksq != 0.f)
        pot = make_float2(__fdiv_rn(-__fmul_rn(P.scale, s.x), ksq), __fdiv_rn(-__fmul_rn(P.scale, s.y), ksq));
      buf[x * XT + c] = pot;
      q[e] = xnyq(k0, P.nyq, P.eps) ? make_float2(0.f, 0.f)
                                    : make_float2(__fmul_rn(k0, pot.y), -__fmul_rn(k0, pot.x));
    }
    __syncthreads();

    // ---- P = IFFT_x(pot): G_y = -i ky P, G_z = -i kz P
    fft_tile<NX, true>(buf, tw);
    const bool zy = xnyq(ky, P.nyq, P.eps), zz = xnyq(kz, P.nyq, P.eps);
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int x = x0 + XSTEP * e;
      const float2 p = buf[x * XT + c];
      if (live) {
        const int64_t g = (int64_t)x * plane + col;
        __stcs(P.out[1] + g, zy ? make_float2(0.f, 0.f) : make_float2(__fmul_rn(ky, p.y), -__fmul_rn(ky, p.x)));
        __stcs(P.out[2] + g, zz ? make_float2(0.f, 0.f) : make_float2(__fmul_rn(kz, p.y), -__fmul_rn(kz, p.x)));
      }
    }
    __syncthreads();

    // ---- G_x = IFFT_x(q)
#pragma unroll
    for (int e = 0; e < EPT; ++e) buf[(x0 + XSTEP * e) * XT + c] = q[e];
    __syncthreads();
    fft_tile<NX, true>(buf, tw);
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int x = x0 + XSTEP * e;
      if (live) __stcs(P.out[0] + (int64_t)x * plane + col, buf[x * XT + c]);
    }
    // swap the roles of the two tile buffers
    float2* tmp = buf; buf = nxt; nxt = tmp;
  }
}

// -------------------------------------------------------------------------- adjoint force
template <int NX>
__global__ void __launch_bounds__(XCfg<NX>::THREADS, XCfg<NX>::CTAS) xfused_force_adj_kernel(XParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* buf = reinterpret_cast<float2*>(smem_raw);
  float2* tw = buf + NX * XCfg<NX>::T;
  float* kx = reinterpret_cast<float*>(tw + NX);
  build_tables<NX>(P, tw, kx);
  __syncthreads();

  constexpr int EPT = XCfg<NX>::EPT;
  constexpr int XT = XCfg<NX>::T;
  const int ztiles = (P.nzc + XT - 1) / XT;
  const int64_t ntiles = (int64_t)P.ny_l * ztiles;
  const int c = threadIdx.x % XT;
  const int x0 = threadIdx.x / XT;
  constexpr int XSTEP = XCfg<NX>::XSTEP;
  const int64_t plane = (int64_t)P.ny_l * P.nzc;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int iy = (int)(tile / ztiles);
    const int kz0 = (int)(tile - (int64_t)iy * ztiles) * XT;
    const bool live = kz0 + c < P.nzc;
    const int64_t col = (int64_t)iy * P.nzc + kz0 + c;
    const float ky = xkval(iy + P.y0, P.ny_g, P.period, false);
    const float kz = xkval(kz0 + c, P.nz_g, P.period, true);
    const float kym = xnyq(ky, P.nyq, P.eps) ? 0.f : ky;
    const float kzm = xnyq(kz, P.nyq, P.eps) ? 0.f : kz;

    // ---- FFT_x(V_x), kept in registers after the transform
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int x = x0 + XSTEP * e;
      buf[x * XT + c] = live ? __ldcs(P.in[0] + (int64_t)x * plane + col) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    fft_tile<NX, false>(buf, tw);
    float2 sx[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) sx[e] = buf[(x0 + XSTEP * e) * XT + c];
    __syncthreads();

    // ---- W = i ky V_y + i kz V_z (ky, kz constant along x), FFT_x(W)
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int x = x0 + XSTEP * e;
      float2 w = make_float2(0.f, 0.f);
      if (live) {
        const int64_t g = (int64_t)x * plane + col;
        const float2 vy = __ldcs(P.in[1] + g), vz = __ldcs(P.in[2] + g);
        w = make_float2(-(kym * vy.y) - kzm * vz.y, kym * vy.x + kzm * vz.x);
      }
      buf[x * XT + c] = w;
    }
    __syncthreads();
    fft_tile<NX, false>(buf, tw);

    // ---- S = (FFT(W) + i kx FFT(V_x)) * (-scale / k^2)
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int x = x0 + XSTEP * e;
      const float k0 = kx[x];
      const float k0m = xnyq(k0, P.nyq, P.eps) ? 0.f : k0;
      float ksq = __fadd_rn(__fadd_rn(__fmul_rn(k0, k0), __fmul_rn(ky, ky)), __fmul_rn(kz, kz));
      float2 w = buf[x * XT + c];
      float2 t = make_float2(w.x - k0m * sx[e].y, w.y + k0m * sx[e].x);
      float2 s = make_float2(0.f, 0.f);
      if (ksq != 0.f) s = make_float2(__fdiv_rn(-(P.scale * t.x), ksq), __fdiv_rn(-(P.scale * t.y), ksq));
      buf[x * XT + c] = s;
    }
    __syncthreads();
    fft_tile<NX, true>(buf, tw);
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int x = x0 + XSTEP * e;
      if (live) __stcs(P.out[0] + (int64_t)x * plane + col, buf[x * XT + c]);
    }
    __syncthreads();
  }
}

template <int NX>
static int launch_x(cudaStream_t st, const XParams& P, bool adjoint) {
  constexpr int XT = XCfg<NX>::T;
  size_t smem = (size_t)NX * XT * sizeof(float2) * ((adjoint || !XCfg<NX>::DBUF) ? 1 : 2) + (size_t)NX * sizeof(float2) +
                (size_t)NX * sizeof(float);
  const int ztiles = (P.nzc + XT - 1) / XT;
  int64_t ntiles = (int64_t)P.ny_l * ztiles;
  int64_t cap = (int64_t)sm_count() * (XCfg<NX>::THREADS >= 1024 ? 1 : 2);
  int grid = (int)(ntiles < cap ? ntiles : cap);
  if (adjoint) {
    PMWD_CUDA_TRY(cudaFuncSetAttribute(xfused_force_adj_kernel<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xfused_force_adj_kernel<NX><<<grid, XCfg<NX>::THREADS, smem, st>>>(P);
  } else {
    PMWD_CUDA_TRY(cudaFuncSetAttribute(xfused_force_kernel<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xfused_force_kernel<NX><<<grid, XCfg<NX>::THREADS, smem, st>>>(P);
  }
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

// which implementation the last pmwd_xpass_* call of this process used (tests assert on it)
static volatile int g_last_variant = 0;

bool xpass_supported(int nx) {
  return nx == 64 || nx == 128 || nx == 256 || nx == 512 || nx == 1024 || nx == 2048;
}

int xpass_run(cudaStream_t st, const int32_t* shape, int y0, int ny_l, double spacing, float scale,
              const void* const* in, void* const* out, bool adjoint) {
  PMWD_REQUIRE(shape && in && out, "null buffer");
  PMWD_REQUIRE(xpass_supported(shape[0]), "fused x-pass supports nx in {64,128,256,512,1024,2048}");
  PMWD_REQUIRE(y0 >= 0 && ny_l > 0 && y0 + ny_l <= shape[1], "bad y slab");
  XParams P;
  memset(&P, 0, sizeof(P));
  P.nx = shape[0]; P.ny_l = ny_l; P.nzc = shape[2] / 2 + 1;
  P.ny_g = shape[1]; P.nz_g = shape[2]; P.y0 = y0;
  const double pi = 3.141592653589793238462643383279502884;
  P.period = 2.0 * pi / spacing;
  P.nyq = (float)(pi / spacing);
  P.eps = (float)((pi / spacing) * 1.1920928955078125e-07);
  P.scale = scale;
  const int nin = adjoint ? 3 : 1, nout = adjoint ? 1 : 3;
  for (int a = 0; a < nin; ++a) { PMWD_REQUIRE(in[a], "null input"); P.in[a] = (const float2*)in[a]; }
  for (int a = 0; a < nout; ++a) { PMWD_REQUIRE(out[a], "null output"); P.out[a] = (float2*)out[a]; }
  // nx in {256, 512, 1024}: register-resident radix-16 transforms (xpass16.cu);
  // PMWD_XPASS16=0 keeps the radix-4 shared-memory kernels below for comparison
  static const bool use16 = [] {
    const char* e = getenv("PMWD_XPASS16");
    return !(e && e[0] == '0');
  }();
  if (use16 && xpass16_supported(P, adjoint)) {
    g_last_variant = 2;
    return xpass16_launch(st, P, adjoint);
  }
  g_last_variant = 1;
  switch (shape[0]) {
    case 64: return launch_x<64>(st, P, adjoint);
    case 128: return launch_x<128>(st, P, adjoint);
    case 256: return launch_x<256>(st, P, adjoint);
    case 512: return launch_x<512>(st, P, adjoint);
    case 1024: return launch_x<1024>(st, P, adjoint);
    default: return launch_x<2048>(st, P, adjoint);
  }
}

}  // namespace pmwd

using namespace pmwd;

// Fused x-pass on data[nx][ny_local][nz/2+1] that has already been transformed over (y, z):
// see the header of this file.  `shape` is the GLOBAL real-space mesh shape.
extern "C" int pmwd_xpass_force(void* stream, const int32_t* shape, int y0, int ny_local,
                                double spacing, float scale, const void* rho2d, void* const* g2d) {
  const void* in[3] = {rho2d, nullptr, nullptr};
  StageTimer t(ST_KSPACE, as_stream(stream));
  return xpass_run(as_stream(stream), shape, y0, ny_local, spacing, scale, in, g2d, false);
}

extern "C" int pmwd_xpass_force_adj(void* stream, const int32_t* shape, int y0, int ny_local,
                                    double spacing, float scale, const void* const* v2d, void* out2d) {
  void* out[3] = {out2d, nullptr, nullptr};
  StageTimer t(ST_KSPACE_ADJ, as_stream(stream));
  return xpass_run(as_stream(stream), shape, y0, ny_local, spacing, scale, v2d, out, true);
}

extern "C" int pmwd_xpass_supported(int nx) { return xpass_supported(nx) ? 1 : 0; }

extern "C" int pmwd_xpass_last_variant(void) { return g_last_variant; }
