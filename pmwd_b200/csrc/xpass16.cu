// Fused x-pass of the FFT Poisson solve, register-resident variant for nx in {256, 512, 1024}
// (same math and interface as xpass.cu, see its header; gravity: pmwd/gravity.py:9-16,37-44,56-64,
// its VJP: pmwd/nbody.py:108-118).
//
// The transforms along x keep 16 points per thread in registers (xfft16.cuh): radix 16 x 16 x R3
// with two shared-memory exchanges per transform.  Global loads, the k-space algebra and the
// stores work on the registers directly, because a thread owns the same 16 points before and
// after every transform.  Two thread-private staging arrays in shared memory (slot e of thread t
// at [e * THREADS + t], never shared between threads, so they need no barrier) hold
//   forward:  A = next tile's spectrum (cp.async prefetch),  B = q = -i kx pot while P is transformed
//   adjoint:  A = V_y, then FFT_x(V_x);  B = V_z, then next tile's V_x (cp.async prefetch)
// Shared memory at nx = 1024: 68 KB exchange + 2 x 64 KB staging + 12 KB tables = 208 KB, one
// 512-thread CTA per SM with up to 128 registers per thread.
#include <cuda_pipeline.h>

#include "xfft16.cuh"
#include "xpass.cuh"

namespace pmwd {

using r16::Cfg;

template <int NX>
struct K16 {
  static constexpr int THREADS = Cfg<NX>::THREADS;
  static constexpr int CTAS = NX == 1024 ? 1 : NX == 512 ? 2 : 4;
  static constexpr size_t SMEM = (size_t)Cfg<NX>::ROWS * r16::T * sizeof(float2) +
                                 2 * (size_t)NX * r16::T * sizeof(float2) + (size_t)NX * sizeof(float2) +
                                 (size_t)NX * sizeof(float);
};

template <int NX, bool INV>
__device__ __forceinline__ void fft16(float2 (&v)[16], float2* ex, const float2* tw, int j, int c) {
  r16::dft16<INV>(v);
  __syncthreads();                       // every earlier read of the exchange buffer is done
  r16::ex_write1<NX>(ex, j, c, v);
  __syncthreads();
  r16::ex_read<NX>(ex, j, c, v);
  r16::twiddle2<NX, INV>(v, tw, j);
  r16::dft16<INV>(v);
  __syncthreads();
  r16::ex_write2<NX>(ex, j, c, v);
  __syncthreads();
  r16::ex_read<NX>(ex, j, c, v);
  r16::stage3<NX, INV>(v, tw, j);
}

struct Tile {
  int iy, kz0;
  bool live;
  int64_t col;
};

__device__ __forceinline__ Tile tile_of(const XParams& P, int64_t tile, int ztiles, int c) {
  Tile t;
  t.iy = (int)(tile / ztiles);
  t.kz0 = (int)(tile - (int64_t)t.iy * ztiles) * r16::T;
  t.live = t.kz0 + c < P.nzc;
  t.col = (int64_t)t.iy * P.nzc + t.kz0 + c;
  return t;
}

// cp.async this thread's 16 points of `src` for tile t into its private staging slots
template <int NX>
__device__ __forceinline__ void stage_in(float2* st, const float2* src, const Tile& t, int j, int64_t plane) {
  if (t.live) {
    const float2* g = src + (int64_t)j * plane + t.col;
#pragma unroll
    for (int e = 0; e < 16; ++e)
      __pipeline_memcpy_async(st + e * Cfg<NX>::THREADS + threadIdx.x, g + (int64_t)(Cfg<NX>::J * e) * plane,
                              sizeof(float2));
  }
}

template <int NX>
__device__ __forceinline__ void build_tables16(const XParams& P, float2* tw, float* kx) {
  for (int n = threadIdx.x; n < NX; n += Cfg<NX>::THREADS) {
    double s, c;
    sincospi(-2.0 * (double)n / (double)NX, &s, &c);
    tw[n] = make_float2((float)c, (float)s);
    kx[n] = xkval(n, NX, P.period, false);
  }
}

// -------------------------------------------------------------------------- forward force
template <int NX>
__global__ void __launch_bounds__(K16<NX>::THREADS, K16<NX>::CTAS) xr16_force_kernel(XParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int XT = r16::T, THREADS = Cfg<NX>::THREADS, J = Cfg<NX>::J;
  float2* ex = reinterpret_cast<float2*>(smem_raw);              // [ROWS][T]  exchange buffer
  float2* stA = ex + Cfg<NX>::ROWS * XT;                          // [16][THREADS]  next tile's input
  float2* stB = stA + NX * XT;                                    // [16][THREADS]  q
  float2* tw = stB + NX * XT;                                     // [NX]
  float* kx = reinterpret_cast<float*>(tw + NX);                  // [NX]
  build_tables16<NX>(P, tw, kx);

  const int ztiles = (P.nzc + XT - 1) / XT;
  const int64_t ntiles = (int64_t)P.ny_l * ztiles;
  const int c = threadIdx.x % XT;
  const int j = threadIdx.x / XT;
  const int64_t plane = (int64_t)P.ny_l * P.nzc;
  float2* mineA = stA + threadIdx.x;
  float2* mineB = stB + threadIdx.x;

  if ((int64_t)blockIdx.x < ntiles) stage_in<NX>(stA, P.in[0], tile_of(P, blockIdx.x, ztiles, c), j, plane);
  __pipeline_commit();
  __syncthreads();                                                // tables

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const Tile t = tile_of(P, tile, ztiles, c);
    const float ky = xkval(t.iy + P.y0, P.ny_g, P.period, false);
    const float kz = xkval(t.kz0 + c, P.nz_g, P.period, true);

    float2 v[16];
    __pipeline_wait_prior(0);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = t.live ? mineA[e * THREADS] : make_float2(0.f, 0.f);
    if (tile + gridDim.x < ntiles) stage_in<NX>(stA, P.in[0], tile_of(P, tile + gridDim.x, ztiles, c), j, plane);
    __pipeline_commit();

    fft16<NX, false>(v, ex, tw, j, c);

    // ---- pot = -(scale S)/k^2 stays in registers; q = -i kx pot is parked in staging B
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float k0 = kx[j + J * e];
      const float ksq = __fadd_rn(__fadd_rn(__fmul_rn(k0, k0), __fmul_rn(ky, ky)), __fmul_rn(kz, kz));
      const float2 s = v[e];
      float2 pot = make_float2(0.f, 0.f);
      if (ksq != 0.f)
        pot = make_float2(__fdiv_rn(-__fmul_rn(P.scale, s.x), ksq), __fdiv_rn(-__fmul_rn(P.scale, s.y), ksq));
      v[e] = pot;
      mineB[e * THREADS] = xnyq(k0, P.nyq, P.eps) ? make_float2(0.f, 0.f)
                                                 : make_float2(__fmul_rn(k0, pot.y), -__fmul_rn(k0, pot.x));
    }

    // ---- P = IFFT_x(pot): G_y = -i ky P, G_z = -i kz P
    fft16<NX, true>(v, ex, tw, j, c);
    if (t.live) {
      const bool zy = xnyq(ky, P.nyq, P.eps), zz = xnyq(kz, P.nyq, P.eps);
      const int64_t g0 = (int64_t)j * plane + t.col;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int64_t g = g0 + (int64_t)(J * e) * plane;
        const float2 p = v[e];
        __stcs(P.out[1] + g, zy ? make_float2(0.f, 0.f) : make_float2(__fmul_rn(ky, p.y), -__fmul_rn(ky, p.x)));
        __stcs(P.out[2] + g, zz ? make_float2(0.f, 0.f) : make_float2(__fmul_rn(kz, p.y), -__fmul_rn(kz, p.x)));
      }
    }

    // ---- G_x = IFFT_x(q)
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = mineB[e * THREADS];
    fft16<NX, true>(v, ex, tw, j, c);
    if (t.live) {
      const int64_t g0 = (int64_t)j * plane + t.col;
#pragma unroll
      for (int e = 0; e < 16; ++e) __stcs(P.out[0] + g0 + (int64_t)(J * e) * plane, v[e]);
    }
  }
  __pipeline_wait_prior(0);
}

// -------------------------------------------------------------------------- adjoint force
template <int NX>
__global__ void __launch_bounds__(K16<NX>::THREADS, K16<NX>::CTAS) xr16_force_adj_kernel(XParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int XT = r16::T, THREADS = Cfg<NX>::THREADS, J = Cfg<NX>::J;
  float2* ex = reinterpret_cast<float2*>(smem_raw);
  float2* stA = ex + Cfg<NX>::ROWS * XT;                          // V_y, then FFT_x(V_x)
  float2* stB = stA + NX * XT;                                    // V_z, then the next tile's V_x
  float2* tw = stB + NX * XT;
  float* kx = reinterpret_cast<float*>(tw + NX);
  build_tables16<NX>(P, tw, kx);

  const int ztiles = (P.nzc + XT - 1) / XT;
  const int64_t ntiles = (int64_t)P.ny_l * ztiles;
  const int c = threadIdx.x % XT;
  const int j = threadIdx.x / XT;
  const int64_t plane = (int64_t)P.ny_l * P.nzc;
  float2* mineA = stA + threadIdx.x;
  float2* mineB = stB + threadIdx.x;

  if ((int64_t)blockIdx.x < ntiles) stage_in<NX>(stB, P.in[0], tile_of(P, blockIdx.x, ztiles, c), j, plane);
  __pipeline_commit();
  __syncthreads();                                                // tables

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const Tile t = tile_of(P, tile, ztiles, c);
    const float ky = xkval(t.iy + P.y0, P.ny_g, P.period, false);
    const float kz = xkval(t.kz0 + c, P.nz_g, P.period, true);
    const float kym = xnyq(ky, P.nyq, P.eps) ? 0.f : ky;
    const float kzm = xnyq(kz, P.nyq, P.eps) ? 0.f : kz;

    // ---- FFT_x(V_x) while V_y, V_z stream into the staging arrays
    float2 v[16];
    __pipeline_wait_prior(0);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = t.live ? mineB[e * THREADS] : make_float2(0.f, 0.f);
    stage_in<NX>(stA, P.in[1], t, j, plane);
    stage_in<NX>(stB, P.in[2], t, j, plane);
    __pipeline_commit();
    fft16<NX, false>(v, ex, tw, j, c);

    // ---- W = i ky V_y + i kz V_z (ky, kz constant along x); FFT_x(V_x) is parked in staging A
    __pipeline_wait_prior(0);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float2 w = make_float2(0.f, 0.f);
      if (t.live) {
        const float2 vy = mineA[e * THREADS], vz = mineB[e * THREADS];
        w = make_float2(-(kym * vy.y) - kzm * vz.y, kym * vy.x + kzm * vz.x);
      }
      mineA[e * THREADS] = v[e];
      v[e] = w;
    }
    if (tile + gridDim.x < ntiles) stage_in<NX>(stB, P.in[0], tile_of(P, tile + gridDim.x, ztiles, c), j, plane);
    __pipeline_commit();
    fft16<NX, false>(v, ex, tw, j, c);

    // ---- S = (FFT(W) + i kx FFT(V_x)) * (-scale / k^2)
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float k0 = kx[j + J * e];
      const float k0m = xnyq(k0, P.nyq, P.eps) ? 0.f : k0;
      const float ksq = __fadd_rn(__fadd_rn(__fmul_rn(k0, k0), __fmul_rn(ky, ky)), __fmul_rn(kz, kz));
      const float2 w = v[e], sx = mineA[e * THREADS];
      const float2 tt = make_float2(w.x - k0m * sx.y, w.y + k0m * sx.x);
      float2 s = make_float2(0.f, 0.f);
      if (ksq != 0.f) s = make_float2(__fdiv_rn(-(P.scale * tt.x), ksq), __fdiv_rn(-(P.scale * tt.y), ksq));
      v[e] = s;
    }
    fft16<NX, true>(v, ex, tw, j, c);
    if (t.live) {
      const int64_t g0 = (int64_t)j * plane + t.col;
#pragma unroll
      for (int e = 0; e < 16; ++e) __stcs(P.out[0] + g0 + (int64_t)(J * e) * plane, v[e]);
    }
  }
  __pipeline_wait_prior(0);
}

template <int NX>
static int launch16(cudaStream_t st, const XParams& P, bool adjoint) {
  const int ztiles = (P.nzc + r16::T - 1) / r16::T;
  const int64_t ntiles = (int64_t)P.ny_l * ztiles;
  const int64_t cap = (int64_t)sm_count() * K16<NX>::CTAS;
  const int grid = (int)(ntiles < cap ? ntiles : cap);
  const int smem = (int)K16<NX>::SMEM;
  if (adjoint) {
    PMWD_CUDA_TRY(cudaFuncSetAttribute(xr16_force_adj_kernel<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    xr16_force_adj_kernel<NX><<<grid, K16<NX>::THREADS, smem, st>>>(P);
  } else {
    PMWD_CUDA_TRY(cudaFuncSetAttribute(xr16_force_kernel<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    xr16_force_kernel<NX><<<grid, K16<NX>::THREADS, smem, st>>>(P);
  }
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

bool xpass16_supported(int nx) { return nx == 256 || nx == 512 || nx == 1024; }

int xpass16_launch(cudaStream_t st, const XParams& P, bool adjoint) {
  switch (P.nx) {
    case 256: return launch16<256>(st, P, adjoint);
    case 512: return launch16<512>(st, P, adjoint);
    case 1024: return launch16<1024>(st, P, adjoint);
    default: PMWD_REQUIRE(false, "register x-pass supports nx in {256,512,1024}");
  }
  return PMWD_EINVAL;
}

}  // namespace pmwd
