// Fused x-pass of the FFT Poisson solve, register-resident variant for nx in {256, 512, 1024}
// (same math and interface as xpass.cu, see its header; gravity: pmwd/gravity.py:9-16,37-44,56-64,
// its VJP: pmwd/nbody.py:108-118).
//
// Memory pattern first.  The x stride of data[nx][ny_l][nzc] is a whole plane (4.2 MB at 1024^3),
// so a column tile is read and written as one short row segment per x.  Measured on B200
// (tools/lab/xpattern.cu, 1 read + 3 write streams at 1024^3): 64-byte segments cap at 2.26 TB/s,
// 128-byte segments issued as 16-byte words reach 3.9 TB/s.  So:
//   * the (y, kz) plane is treated as a flat list of ny_l * nzc independent columns, tiled in
//     runs of 16 columns = 128 bytes that are 128-byte ALIGNED in memory whatever nzc is
//     (nzc = nz/2 + 1 is odd, per-row tiling would misalign every other row);
//   * a thread owns TWO adjacent columns (one float4 per x) and 16 points of each, in registers:
//     x = j + (nx/16) e, e = 0..15, before and after every transform (xfft16.cuh), so the loads,
//     the k-space algebra and the stores work on registers, as 16-byte accesses;
//   * the transforms are radix 16 x 16 x R3 with two shared-memory exchanges each; the two
//     columns of a thread go through a one-column exchange buffer one after the other, which
//     leaves room for
//   * a tile-sized staging array in shared memory (thread-private 16-byte slots, no barrier
//     needed).  It parks the one tile-sized temporary (q = -i kx pot while P is transformed;
//     FFT_x(V_x) in the adjoint) and, once that has been read back, receives the NEXT tile's
//     input by cp.async under the tile's last transform.  (Parking q in the G_x array instead
//     cost 27 % of the kernel: the L2 re-read queued behind the tile's own 256 KB of stores.)
// Shared memory at nx = 1024: 68 KB exchange + 128 KB staging + 12 KB tables; one 512-thread CTA
// per SM with up to 128 registers per thread.
//
// nx = 2048 (the 8-GPU mesh) does not fit one CTA's registers: the same kernels run on a CLUSTER of
// two CTAs (cudaLaunchKernelEx, cluster dimension 2), 1024 threads per tile, each CTA keeping half
// of the butterflies.  The exchange buffer is split between the two CTAs' shared memories so that
// every read is local and the writes go to either CTA through distributed shared memory
// (cluster.map_shared_rank); __syncthreads becomes cluster.sync (xfft16.cuh, last section; index
// arithmetic emulated on the CPU).  Validated on B200 in round 2 (parity test green; the y-slab of
// one of 8 GPUs at 2048^3: 9.0 ms forward / 9.6 ms adjoint against 15.5 / 12.3 ms for the radix-4
// kernel of xpass.cu, profiles/r02_validate_pending.txt); PMWD_XPASS16_2048=0 selects the latter.
#include <cooperative_groups.h>
#include <cuda_pipeline.h>
#include <stdlib.h>

#include "xfft16.cuh"
#include "xpass.cuh"

namespace pmwd {

using r16::Cfg;
using r16::TP;

template <int NX>
struct K16 {
  static constexpr bool CL = NX == 2048;                       // two-CTA cluster per tile
  static constexpr int THREADS = CL ? 512 : Cfg<NX>::J * TP;   // per CTA
  static constexpr int CTAS = NX >= 1024 ? 1 : NX == 512 ? 2 : 4;
  static constexpr int LROWS = CL ? Cfg<1024>::ROWS : Cfg<NX>::ROWS;            // rows of this CTA's exchange buffer
  static constexpr size_t EX = (size_t)LROWS * TP * sizeof(float2);             // one-column exchange (padded rows)
  static constexpr size_t STG = (size_t)16 * THREADS * sizeof(float4);         // [16][THREADS] staging slots
  static constexpr size_t SMEM = EX + STG + (size_t)NX * sizeof(float2) + (size_t)NX * sizeof(float);
};

namespace cg = cooperative_groups;

// Who a thread is within its tile, and where the exchange buffers are.  One CTA: j = jl and all
// three pointers are the CTA's own buffer.  Cluster: j = 64 rank + jl, ex0 / ex1 = the buffers of
// CTA 0 / 1 (one of them through distributed shared memory), self = this CTA's own.
struct XCtx {
  int j, jl, cp, rank;
  float2 *ex0, *ex1, *self;
};

template <int NX>
__device__ __forceinline__ XCtx make_ctx(float2* ex_local) {
  XCtx c;
  c.cp = threadIdx.x % TP;
  c.jl = threadIdx.x / TP;
  if (K16<NX>::CL) {
    cg::cluster_group cl = cg::this_cluster();
    c.rank = (int)cl.block_rank();
    c.j = 64 * c.rank + c.jl;
    float2* peer = cl.map_shared_rank(ex_local, c.rank ^ 1);
    c.ex0 = c.rank == 0 ? ex_local : peer;
    c.ex1 = c.rank == 0 ? peer : ex_local;
  } else {
    c.rank = 0;
    c.j = c.jl;
    c.ex0 = c.ex1 = ex_local;
  }
  c.self = ex_local;
  return c;
}

// barrier over everything that shares the exchange buffer(s)
template <int NX>
__device__ __forceinline__ void xsync() {
  if (K16<NX>::CL) cg::this_cluster().sync(); else __syncthreads();
}
template <int NX>
__device__ __forceinline__ void xwrite1(const XCtx& c, const float2 (&v)[16]) {
  if constexpr (K16<NX>::CL) r16::c2k_write1(c.ex0, c.ex1, c.j, c.cp, v); else r16::ex_write1<NX>(c.self, c.j, c.cp, v);
}
template <int NX>
__device__ __forceinline__ void xwrite2(const XCtx& c, const float2 (&v)[16]) {
  if constexpr (K16<NX>::CL) r16::c2k_write2(c.ex0, c.ex1, c.j, c.cp, v); else r16::ex_write2<NX>(c.self, c.j, c.cp, v);
}
template <int NX>
__device__ __forceinline__ void xread(const XCtx& c, float2 (&v)[16]) {
  if constexpr (K16<NX>::CL) r16::ex_read<1024>(c.self, c.jl, c.cp, v); else r16::ex_read<NX>(c.self, c.j, c.cp, v);
}

// Both columns of a thread through the one-column exchange buffer, a then b; the register
// arithmetic of one column sits between the barriers of the other column's exchange.
template <int NX, bool INV>
__device__ __forceinline__ void fft16x2(float2 (&a)[16], float2 (&b)[16], const XCtx& c, const float2* tw) {
  r16::dft16<INV>(a);
  xsync<NX>();                           // every earlier read of the exchange buffer is done
  xwrite1<NX>(c, a);
  r16::dft16<INV>(b);
  xsync<NX>();
  xread<NX>(c, a);
  xsync<NX>();
  xwrite1<NX>(c, b);
  r16::twiddle2<NX, INV>(a, tw, c.j);
  r16::dft16<INV>(a);
  xsync<NX>();
  xread<NX>(c, b);
  xsync<NX>();
  xwrite2<NX>(c, a);
  r16::twiddle2<NX, INV>(b, tw, c.j);
  r16::dft16<INV>(b);
  xsync<NX>();
  xread<NX>(c, a);
  xsync<NX>();
  xwrite2<NX>(c, b);
  r16::stage3<NX, INV>(a, tw, c.j);
  xsync<NX>();
  xread<NX>(c, b);
  r16::stage3<NX, INV>(b, tw, c.j);
}

// cp.async this thread's 16 float4 words (two columns x 16 points) of `src` into its staging slots
template <int NX>
__device__ __forceinline__ void stage_in(float4* mine, const float2* src, int64_t g0, int64_t plane, bool live) {
  if (live) {
#pragma unroll
    for (int e = 0; e < 16; ++e)
      __pipeline_memcpy_async(mine + e * K16<NX>::THREADS, src + g0 + (int64_t)(Cfg<NX>::J * e) * plane, sizeof(float4));
  }
}

// wavenumbers of flat column m = iy * nzc + kz (m < ny_l * nzc fits 32 bits)
struct ColK {
  float ky, kz;
};
__device__ __forceinline__ ColK col_k(const XParams& P, unsigned m) {
  const unsigned iy = m / (unsigned)P.nzc;
  const unsigned iz = m - iy * (unsigned)P.nzc;
  ColK k;
  k.ky = xkval((int)iy + P.y0, P.ny_g, P.period, false);
  k.kz = xkval((int)iz, P.nz_g, P.period, true);
  return k;
}

template <int NX>
__device__ __forceinline__ void build_tables16(const XParams& P, float2* tw, float* kx) {
  for (int n = threadIdx.x; n < NX; n += K16<NX>::THREADS) {
    double s, c;
    sincospi(-2.0 * (double)n / (double)NX, &s, &c);
    tw[n] = make_float2((float)c, (float)s);
    kx[n] = xkval(n, NX, P.period, false);
  }
}

__device__ __forceinline__ float ksq_of(float k0, float ky, float kz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(k0, k0), __fmul_rn(ky, ky)), __fmul_rn(kz, kz));
}
// -scale / k^2 (0 at k = 0): one reciprocal per mode; the transforms around it are accurate to
// ~1e-7 relative anyway, so the correctly rounded quotient of the standalone k-space kernel
// (kspace.cu, bit-exact with the oracle) would buy nothing here
__device__ __forceinline__ float green_of(float ksq, float scale) {
  return ksq != 0.f ? -scale * __frcp_rn(ksq) : 0.f;
}
// -i k p, zero on a Nyquist plane
__device__ __forceinline__ float2 neg_ik(float k, bool zero, float2 p) {
  return zero ? make_float2(0.f, 0.f) : make_float2(__fmul_rn(k, p.y), -__fmul_rn(k, p.x));
}

// Schedule of one tile (both kernels): the staging slots hold this tile's input (prefetched
// during the previous tile's last transform), then the parked temporary, then -- as soon as
// that has been read back -- the next tile's input, streaming in under the last transform.
// No global load ever has to be waited for behind the tile's own stores.

// -------------------------------------------------------------------------- forward force
template <int NX>
__global__ void __launch_bounds__(K16<NX>::THREADS, K16<NX>::CTAS) xr16_force_kernel(XParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int J = Cfg<NX>::J, THREADS = K16<NX>::THREADS;
  float2* ex = reinterpret_cast<float2*>(smem_raw);               // [ROWS][TP]  one-column exchange buffer
  float4* stg = reinterpret_cast<float4*>(smem_raw + K16<NX>::EX); // [16][THREADS]  staging slots
  float2* tw = reinterpret_cast<float2*>(stg + 16 * THREADS);     // [NX]
  float* kx = reinterpret_cast<float*>(tw + NX);                  // [NX]
  build_tables16<NX>(P, tw, kx);

  const int64_t plane = (int64_t)P.ny_l * P.nzc;                  // columns; even (checked by the host)
  const int64_t ntiles = (plane + 2 * TP - 1) / (2 * TP);
  const XCtx cx = make_ctx<NX>(ex);
  const int cp = cx.cp, j = cx.j;
  float4* mine = stg + threadIdx.x;
  // one tile per CTA, or per cluster of two
  const int64_t tile0 = K16<NX>::CL ? blockIdx.x >> 1 : blockIdx.x;
  const int64_t tstep = K16<NX>::CL ? gridDim.x >> 1 : gridDim.x;

  {
    const int64_t m = tile0 * (2 * TP) + 2 * cp;
    stage_in<NX>(mine, P.in[0], (int64_t)j * plane + m, plane, m < plane);
    __pipeline_commit();
  }
  xsync<NX>();                                                    // tables; peer CTA is resident

  for (int64_t tile = tile0; tile < ntiles; tile += tstep) {
    const int64_t m = tile * (2 * TP) + 2 * cp;                   // this thread's columns m, m + 1
    const bool live = m < plane;
    const ColK ka = col_k(P, live ? (unsigned)m : 0u), kb = col_k(P, live ? (unsigned)m + 1u : 0u);
    const int64_t g0 = (int64_t)j * plane + m;                    // float2 index of (x = j, column m)

    float2 a[16], b[16];
    __pipeline_wait_prior(0);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) q = mine[e * THREADS];
      a[e] = make_float2(q.x, q.y);
      b[e] = make_float2(q.z, q.w);
    }
    fft16x2<NX, false>(a, b, cx, tw);

    // ---- pot = -(scale S)/k^2 stays in registers; q = -i kx pot is parked in the staging slots
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float k0 = kx[j + J * e];
      const bool zx = xnyq(k0, P.nyq, P.eps);
      const float ga = green_of(ksq_of(k0, ka.ky, ka.kz), P.scale), gb = green_of(ksq_of(k0, kb.ky, kb.kz), P.scale);
      a[e] = make_float2(a[e].x * ga, a[e].y * ga);
      b[e] = make_float2(b[e].x * gb, b[e].y * gb);
      mine[e * THREADS] = r16::pack4(neg_ik(k0, zx, a[e]), neg_ik(k0, zx, b[e]));
    }

    // ---- P = IFFT_x(pot): G_y = -i ky P, G_z = -i kz P
    fft16x2<NX, true>(a, b, cx, tw);
    if (live) {
      const bool zya = xnyq(ka.ky, P.nyq, P.eps), zza = xnyq(ka.kz, P.nyq, P.eps);
      const bool zyb = xnyq(kb.ky, P.nyq, P.eps), zzb = xnyq(kb.kz, P.nyq, P.eps);
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int64_t g = g0 + (int64_t)(J * e) * plane;
        __stcs(reinterpret_cast<float4*>(P.out[1] + g), r16::pack4(neg_ik(ka.ky, zya, a[e]), neg_ik(kb.ky, zyb, b[e])));
        __stcs(reinterpret_cast<float4*>(P.out[2] + g), r16::pack4(neg_ik(ka.kz, zza, a[e]), neg_ik(kb.kz, zzb, b[e])));
      }
    }

    // ---- G_x = IFFT_x(q); the next tile's spectrum streams into the freed slots meanwhile
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float4 q = mine[e * THREADS];
      a[e] = make_float2(q.x, q.y);
      b[e] = make_float2(q.z, q.w);
    }
    {
      const int64_t mn = m + tstep * (2 * TP);       // same columns of this CTA's next tile
      stage_in<NX>(mine, P.in[0], (int64_t)j * plane + mn, plane, mn < plane);
      __pipeline_commit();
    }
    fft16x2<NX, true>(a, b, cx, tw);
    if (live) {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        __stcs(reinterpret_cast<float4*>(P.out[0] + g0 + (int64_t)(J * e) * plane), r16::pack4(a[e], b[e]));
    }
  }
  __pipeline_wait_prior(0);
  xsync<NX>();                           // (cluster) the peer may still be writing into this CTA's buffer
}

// -------------------------------------------------------------------------- adjoint force
template <int NX>
__global__ void __launch_bounds__(K16<NX>::THREADS, K16<NX>::CTAS) xr16_force_adj_kernel(XParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int J = Cfg<NX>::J, THREADS = K16<NX>::THREADS;
  float2* ex = reinterpret_cast<float2*>(smem_raw);
  float4* stg = reinterpret_cast<float4*>(smem_raw + K16<NX>::EX); // V_x, then FFT_x(V_x), then the next tile's V_x
  float2* tw = reinterpret_cast<float2*>(stg + 16 * THREADS);
  float* kx = reinterpret_cast<float*>(tw + NX);
  build_tables16<NX>(P, tw, kx);

  const int64_t plane = (int64_t)P.ny_l * P.nzc;
  const int64_t ntiles = (plane + 2 * TP - 1) / (2 * TP);
  const XCtx cx = make_ctx<NX>(ex);
  const int cp = cx.cp, j = cx.j;
  float4* mine = stg + threadIdx.x;
  const int64_t tile0 = K16<NX>::CL ? blockIdx.x >> 1 : blockIdx.x;
  const int64_t tstep = K16<NX>::CL ? gridDim.x >> 1 : gridDim.x;

  {
    const int64_t m = tile0 * (2 * TP) + 2 * cp;
    stage_in<NX>(mine, P.in[0], (int64_t)j * plane + m, plane, m < plane);
    __pipeline_commit();
  }
  xsync<NX>();                                                    // tables; peer CTA is resident

  for (int64_t tile = tile0; tile < ntiles; tile += tstep) {
    const int64_t m = tile * (2 * TP) + 2 * cp;
    const bool live = m < plane;
    const ColK ka = col_k(P, live ? (unsigned)m : 0u), kb = col_k(P, live ? (unsigned)m + 1u : 0u);
    const float kyma = xnyq(ka.ky, P.nyq, P.eps) ? 0.f : ka.ky, kzma = xnyq(ka.kz, P.nyq, P.eps) ? 0.f : ka.kz;
    const float kymb = xnyq(kb.ky, P.nyq, P.eps) ? 0.f : kb.ky, kzmb = xnyq(kb.kz, P.nyq, P.eps) ? 0.f : kb.kz;
    const int64_t g0 = (int64_t)j * plane + m;

    // ---- FFT_x(V_x), parked in the staging slots
    float2 a[16], b[16];
    __pipeline_wait_prior(0);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) q = mine[e * THREADS];
      a[e] = make_float2(q.x, q.y);
      b[e] = make_float2(q.z, q.w);
    }
    fft16x2<NX, false>(a, b, cx, tw);
#pragma unroll
    for (int e = 0; e < 16; ++e) mine[e * THREADS] = r16::pack4(a[e], b[e]);

    // ---- W = i ky V_y + i kz V_z (ky, kz constant along x), FFT_x(W)
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float4 vy = make_float4(0.f, 0.f, 0.f, 0.f), vz = vy;
      if (live) {
        const int64_t g = g0 + (int64_t)(J * e) * plane;
        vy = __ldcs(reinterpret_cast<const float4*>(P.in[1] + g));
        vz = __ldcs(reinterpret_cast<const float4*>(P.in[2] + g));
      }
      a[e] = make_float2(-(kyma * vy.y) - kzma * vz.y, kyma * vy.x + kzma * vz.x);
      b[e] = make_float2(-(kymb * vy.w) - kzmb * vz.w, kymb * vy.z + kzmb * vz.z);
    }
    fft16x2<NX, false>(a, b, cx, tw);

    // ---- S = (FFT(W) + i kx FFT(V_x)) * (-scale / k^2); then the next tile's V_x streams in
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float k0 = kx[j + J * e];
      const float k0m = xnyq(k0, P.nyq, P.eps) ? 0.f : k0;
      const float4 sx = mine[e * THREADS];
      const float ga = green_of(ksq_of(k0, ka.ky, ka.kz), P.scale), gb = green_of(ksq_of(k0, kb.ky, kb.kz), P.scale);
      a[e] = make_float2((a[e].x - k0m * sx.y) * ga, (a[e].y + k0m * sx.x) * ga);
      b[e] = make_float2((b[e].x - k0m * sx.w) * gb, (b[e].y + k0m * sx.z) * gb);
    }
    {
      const int64_t mn = m + tstep * (2 * TP);
      stage_in<NX>(mine, P.in[0], (int64_t)j * plane + mn, plane, mn < plane);
      __pipeline_commit();
    }
    fft16x2<NX, true>(a, b, cx, tw);
    if (live) {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        __stcs(reinterpret_cast<float4*>(P.out[0] + g0 + (int64_t)(J * e) * plane), r16::pack4(a[e], b[e]));
    }
  }
  __pipeline_wait_prior(0);
  xsync<NX>();                           // (cluster) the peer may still be writing into this CTA's buffer
}

template <int NX>
static int launch16(cudaStream_t st, const XParams& P, bool adjoint) {
  const int64_t plane = (int64_t)P.ny_l * P.nzc;
  const int64_t ntiles = (plane + 2 * TP - 1) / (2 * TP);
  const int csize = K16<NX>::CL ? 2 : 1;                                   // CTAs per tile
  const int smem = (int)K16<NX>::SMEM;
  void (*kern)(XParams) = adjoint ? xr16_force_adj_kernel<NX> : xr16_force_kernel<NX>;
  PMWD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(K16<NX>::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = K16<NX>::CL ? 1 : 0;
  // resident tiles: every tile slot must be co-resident (persistent loop), so a cluster launch asks
  // the driver how many clusters fit at once (GPC boundaries can leave an SM without a partner)
  int64_t cap = (int64_t)sm_count() * K16<NX>::CTAS / csize;
  if (K16<NX>::CL) {
    static int max_clusters[2] = {0, 0};
    if (max_clusters[adjoint] == 0) {
      cfg.gridDim = dim3((unsigned)(cap * csize));
      int n = 0;
      PMWD_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      PMWD_REQUIRE(n > 0, "no two-CTA cluster of the x-pass kernel fits on this device");
      max_clusters[adjoint] = n;
    }
    if (cap > max_clusters[adjoint]) cap = max_clusters[adjoint];
  }
  const int grid = (int)(ntiles < cap ? ntiles : cap) * csize;
  cfg.gridDim = dim3(grid);
  PMWD_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, P));
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

// 16-byte accesses need an even number of columns per x plane and 16-byte aligned arrays.
// nx = 2048 runs on two-CTA clusters; PMWD_XPASS16_2048=0 falls back to the radix-4 kernel (A/B).
bool xpass16_supported(const XParams& P, bool adjoint) {
  static const bool allow2k = [] {
    const char* e = getenv("PMWD_XPASS16_2048");
    return !(e && e[0] == '0');
  }();
  if (!(P.nx == 256 || P.nx == 512 || P.nx == 1024 || (P.nx == 2048 && allow2k))) return false;
  if ((((int64_t)P.ny_l * P.nzc) & 1) != 0) return false;
  const int nin = adjoint ? 3 : 1, nout = adjoint ? 1 : 3;
  for (int i = 0; i < nin; ++i)
    if ((uintptr_t)P.in[i] % 16 != 0) return false;
  for (int i = 0; i < nout; ++i)
    if ((uintptr_t)P.out[i] % 16 != 0) return false;
  return true;
}

int xpass16_launch(cudaStream_t st, const XParams& P, bool adjoint) {
  switch (P.nx) {
    case 256: return launch16<256>(st, P, adjoint);
    case 512: return launch16<512>(st, P, adjoint);
    case 1024: return launch16<1024>(st, P, adjoint);
    case 2048: return launch16<2048>(st, P, adjoint);
    default: PMWD_REQUIRE(false, "register x-pass supports nx in {256,512,1024,2048}");
  }
  return PMWD_EINVAL;
}

}  // namespace pmwd
