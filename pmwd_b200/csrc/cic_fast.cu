// Fused CIC kernels for the gravity fast path (3-D, int16 pmid, offset 0, cell_size=None,
// target mesh == conf.mesh_shape): the shapes gravity() and its VJP use
// (pmwd/gravity.py:47-72, pmwd/nbody.py:102-118).
//
//  * scatter_fast<NCH>   : rho (NCH=1, scalar val) or the 3 cotangent meshes V_i = scatter(pi_i)
//                          (NCH=3, SoA meshes) -- pmwd/scatter.py:60-83 / gather.py:113.
//  * gather3_kernel      : acc_i = gather(F_i), i=0..2 from 3 SoA force meshes in ONE pass
//                          (the reference calls gather 3x and recomputes the stencil 3x,
//                          pmwd/gravity.py:61-69), optionally fused with the trailing half-kick
//                          vel += acc*K (pmwd/nbody.py:70-77).
//  * force_adj_gather    : alpha = disp cotangent of gravity: the disp_cot parts of _gather_bwd x3
//                          (gather.py:106-110) and of _scatter_bwd (scatter.py:112-116) in one pass.
//
// All HBM / L2-atomic bound.  Particles arrive in Lagrangian (C-order) sequence, so the 32
// lanes of a warp touch 2..4 contiguous z-rows of the mesh: loads coalesce and the float32
// reductions (RED.E.ADD.F32 / .F32x2) land in the same L2 sectors.
#include <math.h>
#include <stdlib.h>

#include "cic.cuh"

namespace pmwd {

struct FastParams {
  int64_t n;
  int nx, ny, nz;   // periodic wrap shape (conf.mesh_shape; the GLOBAL mesh on a slab)
  int nx_ext;       // x planes held by the mesh array (== nx on one GPU; slab + halos otherwise)
  int xoff;         // global x index of plane 0 of the array (offset / cell, an integer)
  float cell;
  int cs;           // particle arrays are streamed once: load / store them with the .cs (evict-first) hint so
                    // that they do not push the force-mesh lines the stencils re-use out of L2
  float inv_cell;   // 1 / cell if cell is a power of two (the product is then the same correctly rounded
                    // number as the IEEE division, at a tenth of its instructions), else 0
};

template <typename T>
__device__ __forceinline__ T ld_p(const T* p, int cs) { return cs ? __ldcs(p) : *p; }
template <typename T>
__device__ __forceinline__ void st_p(T* p, T v, int cs) { if (cs) __stcs(p, v); else *p = v; }

// x / cell in float32, bitwise what the reference's division gives (pm_util.py:133, gather.py:108-110)
__device__ __forceinline__ float div_cell(float x, float cell, float inv_cell) {
  return inv_cell != 0.f ? __fmul_rn(x, inv_cell) : __fdiv_rn(x, cell);
}

int slab_xoff(const pmwd_cic_desc* d);
bool cic_is_fast(const pmwd_cic_desc* d);

struct Stencil3 {
  int ix[2], iy[2], iz[2];
  float wx[2], wy[2], wz[2];
};

template <bool GRAD>
struct Stencil3G : Stencil3 {
  float sx[2], sy[2], sz[2];
};

__device__ __forceinline__ void axis_fast(int pm, float disp, float cell, float inv_cell, int n, int* idx, float* w,
                                          float* s) {
  float t = div_cell(disp, cell, inv_cell);
  int i0 = (int)floorf(t);
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    float d = __fsub_rn(t, (float)(i0 + b));
    w[b] = __fsub_rn(1.f, fabsf(d));
    if (s) s[b] = sign_neg(d);
  }
  int i = wrap_index(pm + i0, n);
  idx[0] = i;
  idx[1] = (i + 1 == n) ? 0 : i + 1;
}

// Global (wrapped) x index -> plane of the local array, or -1 if the array does not hold it
// (enmesh's s2 drop, pm_util.py:140-141 + jax `mode='drop'`).
__device__ __forceinline__ void localize_x(const FastParams& P, int* ix) {
  if (P.xoff == 0 && P.nx_ext == P.nx) return;
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    int l = ix[b] - P.xoff;
    if (l < 0) l += P.nx;
    ix[b] = l < P.nx_ext ? l : -1;
  }
}

__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }

__device__ __forceinline__ void red_add2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// ---------------------------------------------------------------------------------------
// Warp-aggregated reductions: lanes of a warp whose particles sit in the same base cell (adjacent
// in the column-ordered storage of a clustered run) first add their 8 x NCH contributions with a
// log-step shuffle reduction over the peer group (__match_any_sync), then only the group's
// first lane issues the RED.E.ADD.F32(x2)s.  No peers (early times) costs one match + one vote.
template <int NV>
__device__ __forceinline__ bool reduce_peers(unsigned peers, float (&v)[NV]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int rel_pos = __popc(peers << (31 - lane) << 1);        // peers with a lower lane id
  const bool first = rel_pos == 0;
  peers &= (0xfffffffeu << lane);                          // peers with a higher lane id
  while (__any_sync(full, peers)) {
    const int next = __ffs(peers);
    const int src = next ? next - 1 : lane;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float t = __shfl_sync(full, v[i], src);
      if (next) v[i] += t;
    }
    const int done = rel_pos & 1;
    peers &= __ballot_sync(full, !done);
    rel_pos >>= 1;
  }
  return first;
}

template <int NCH>
__global__ void __launch_bounds__(256)
scatter_fast_kernel(FastParams P, const short* __restrict__ pmid, const float* __restrict__ disp,
                    const float* __restrict__ val, float val_scalar,
                    float* __restrict__ m0, float* __restrict__ m1, float* __restrict__ m2) {
  // whole warps iterate together (the shuffles below need all 32 lanes)
  const int64_t n_round = (P.n + 31) & ~(int64_t)31;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_round;
       p += (int64_t)gridDim.x * blockDim.x) {
    const bool active = p < P.n;
    const int64_t pp = active ? p : P.n - 1;
    Stencil3 s;
    axis_fast(pmid[3 * pp + 0], disp[3 * pp + 0], P.cell, P.inv_cell, P.nx, s.ix, s.wx, nullptr);
    axis_fast(pmid[3 * pp + 1], disp[3 * pp + 1], P.cell, P.inv_cell, P.ny, s.iy, s.wy, nullptr);
    axis_fast(pmid[3 * pp + 2], disp[3 * pp + 2], P.cell, P.inv_cell, P.nz, s.iz, s.wz, nullptr);
    // group key: the (global) base cell; inactive tail lanes get unique keys
    const unsigned long long key =
        active ? (unsigned long long)(((int64_t)s.ix[0] * P.ny + s.iy[0]) * P.nz + s.iz[0])
               : ~0ull - (unsigned long long)(threadIdx.x & 31);
    localize_x(P, s.ix);
    // contributions c[(bx*2+by)*2 + bz][ch] = val[ch] * ((wx*wy)*wz)
    float c[8 * NCH];
#pragma unroll
    for (int bx = 0; bx < 2; ++bx)
#pragma unroll
      for (int by = 0; by < 2; ++by) {
        const float wxy = __fmul_rn(s.wx[bx], s.wy[by]);
#pragma unroll
        for (int bz = 0; bz < 2; ++bz) {
          const float w = __fmul_rn(wxy, s.wz[bz]);
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            const float v = val ? val[NCH * pp + ch] : val_scalar;
            c[((bx * 2 + by) * 2 + bz) * NCH + ch] = active ? __fmul_rn(v, w) : 0.f;
          }
        }
      }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const bool first = reduce_peers<8 * NCH>(peers, c);
    if (!first || !active) continue;
    // z-neighbours are adjacent in memory: one 8-byte reduction when the pair is aligned
    const bool pair = (s.iz[1] == s.iz[0] + 1) && ((s.iz[0] & 1) == 0) && ((P.nz & 1) == 0);
#pragma unroll
    for (int bx = 0; bx < 2; ++bx) {
      if (s.ix[bx] < 0) continue;
#pragma unroll
      for (int by = 0; by < 2; ++by) {
        const int64_t row = ((int64_t)s.ix[bx] * P.ny + s.iy[by]) * P.nz;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          float* m = ch == 0 ? m0 : (ch == 1 ? m1 : m2);
          const float a = c[((bx * 2 + by) * 2 + 0) * NCH + ch], b = c[((bx * 2 + by) * 2 + 1) * NCH + ch];
          if (pair) {
            red_add2(m + row + s.iz[0], a, b);
          } else {
            red_add(m + row + s.iz[0], a);
            red_add(m + row + s.iz[1], b);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// KICK: 0 = acc only; 1 = + trailing half-kick vel += acc*K; 2 = additionally the NEXT step's
// leading half-kick and drift (vel += acc*K1n; disp += vel*Dn): the whole per-particle update of
// a KDK step in the gather pass, same float32 operation sequence as the separate kernels.
template <int KICK>
__global__ void __launch_bounds__(256)
gather3_kernel(FastParams P, const short* __restrict__ pmid, const float* disp,
               const float* __restrict__ f0, const float* __restrict__ f1,
               const float* __restrict__ f2, float* __restrict__ acc, float* __restrict__ vel,
               float K, float K1n, float Dn, float* disp_rw) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P.n;
       p += (int64_t)gridDim.x * blockDim.x) {
    Stencil3 s;
    const int cs = P.cs;
    const float dp0 = ld_p(disp + 3 * p + 0, cs), dp1 = ld_p(disp + 3 * p + 1, cs), dp2 = ld_p(disp + 3 * p + 2, cs);
    axis_fast(ld_p(pmid + 3 * p + 0, cs), dp0, P.cell, P.inv_cell, P.nx, s.ix, s.wx, nullptr);
    axis_fast(ld_p(pmid + 3 * p + 1, cs), dp1, P.cell, P.inv_cell, P.ny, s.iy, s.wy, nullptr);
    axis_fast(ld_p(pmid + 3 * p + 2, cs), dp2, P.cell, P.inv_cell, P.nz, s.iz, s.wz, nullptr);
    localize_x(P, s.ix);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    // neighbour order n = bx + 2 by + 4 bz (axis 0 = LSB, pm_util.py:95-97), summed sequentially
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int bx = n & 1, by = (n >> 1) & 1, bz = n >> 2;
      float w = __fmul_rn(__fmul_rn(s.wx[bx], s.wy[by]), s.wz[bz]);
      const bool ok = s.ix[bx] >= 0;           // dropped neighbour reads 0 (fill_value=0)
      int64_t lin = ((int64_t)(ok ? s.ix[bx] : 0) * P.ny + s.iy[by]) * P.nz + s.iz[bz];
      a0 = __fadd_rn(a0, __fmul_rn(ok ? __ldg(f0 + lin) : 0.f, w));
      a1 = __fadd_rn(a1, __fmul_rn(ok ? __ldg(f1 + lin) : 0.f, w));
      a2 = __fadd_rn(a2, __fmul_rn(ok ? __ldg(f2 + lin) : 0.f, w));
    }
    st_p(acc + 3 * p + 0, a0, cs);
    st_p(acc + 3 * p + 1, a1, cs);
    st_p(acc + 3 * p + 2, a2, cs);
    if (KICK) {
      float v0 = __fadd_rn(ld_p(vel + 3 * p + 0, cs), __fmul_rn(a0, K));
      float v1 = __fadd_rn(ld_p(vel + 3 * p + 1, cs), __fmul_rn(a1, K));
      float v2 = __fadd_rn(ld_p(vel + 3 * p + 2, cs), __fmul_rn(a2, K));
      if (KICK == 2) {
        v0 = __fadd_rn(v0, __fmul_rn(a0, K1n));
        v1 = __fadd_rn(v1, __fmul_rn(a1, K1n));
        v2 = __fadd_rn(v2, __fmul_rn(a2, K1n));
        st_p(disp_rw + 3 * p + 0, __fadd_rn(dp0, __fmul_rn(v0, Dn)), cs);
        st_p(disp_rw + 3 * p + 1, __fadd_rn(dp1, __fmul_rn(v1, Dn)), cs);
        st_p(disp_rw + 3 * p + 2, __fadd_rn(dp2, __fmul_rn(v2, Dn)), cs);
      }
      st_p(vel + 3 * p + 0, v0, cs);
      st_p(vel + 3 * p + 1, v1, cs);
      st_p(vel + 3 * p + 2, v2, cs);
    }
  }
}

// ---------------------------------------------------------------------------------------
// alpha_j = [ sum_i sum_n pi_i F_i[n] g_nj + sum_n (val * rho_cot[n]) g_nj ] / cell
// evaluated as the reference does: four separate (gather-VJP x3, scatter-VJP) neighbour sums,
// each divided by cell, then added (gather.py:108-110, scatter.py:114-116, JAX sums the
// cotangent contributions to ptcl.disp).
__global__ void __launch_bounds__(256)
force_adj_gather_kernel(FastParams P, const short* __restrict__ pmid, const float* __restrict__ disp,
                        const float* __restrict__ f0, const float* __restrict__ f1,
                        const float* __restrict__ f2, const float* __restrict__ rho_cot,
                        const float* __restrict__ pi, float val, float* __restrict__ alpha,
                        float* __restrict__ acc) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P.n;
       p += (int64_t)gridDim.x * blockDim.x) {
    Stencil3G<true> s;
    const int cs = P.cs;
    axis_fast(ld_p(pmid + 3 * p + 0, cs), ld_p(disp + 3 * p + 0, cs), P.cell, P.inv_cell, P.nx, s.ix, s.wx, s.sx);
    axis_fast(ld_p(pmid + 3 * p + 1, cs), ld_p(disp + 3 * p + 1, cs), P.cell, P.inv_cell, P.ny, s.iy, s.wy, s.sy);
    axis_fast(ld_p(pmid + 3 * p + 2, cs), ld_p(disp + 3 * p + 2, cs), P.cell, P.inv_cell, P.nz, s.iz, s.wz, s.sz);
    localize_x(P, s.ix);
    const float p0 = ld_p(pi + 3 * p + 0, cs), p1 = ld_p(pi + 3 * p + 1, cs), p2 = ld_p(pi + 3 * p + 2, cs);
    float d[4][3];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;     // acc = gather3 of the same force values (same arithmetic as gather3_kernel)
#pragma unroll
    for (int q = 0; q < 4; ++q) d[q][0] = d[q][1] = d[q][2] = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int bx = n & 1, by = (n >> 1) & 1, bz = n >> 2;
      // g_j = sign(-d_j) * prod_{m != j} w_m in axis order (j+1.., 0..j-1)
      float gx = __fmul_rn(s.sx[bx], __fmul_rn(s.wy[by], s.wz[bz]));
      float gy = __fmul_rn(s.sy[by], __fmul_rn(s.wz[bz], s.wx[bx]));
      float gz = __fmul_rn(s.sz[bz], __fmul_rn(s.wx[bx], s.wy[by]));
      const bool ok = s.ix[bx] >= 0;
      int64_t lin = ((int64_t)(ok ? s.ix[bx] : 0) * P.ny + s.iy[by]) * P.nz + s.iz[bz];
      const float F0 = ok ? __ldg(f0 + lin) : 0.f, F1 = ok ? __ldg(f1 + lin) : 0.f, F2 = ok ? __ldg(f2 + lin) : 0.f;
      if (acc) {
        const float w = __fmul_rn(__fmul_rn(s.wx[bx], s.wy[by]), s.wz[bz]);
        a0 = __fadd_rn(a0, __fmul_rn(F0, w));
        a1 = __fadd_rn(a1, __fmul_rn(F1, w));
        a2 = __fadd_rn(a2, __fmul_rn(F2, w));
      }
      float t[4];
      t[0] = __fmul_rn(p0, F0);
      t[1] = __fmul_rn(p1, F1);
      t[2] = __fmul_rn(p2, F2);
      t[3] = __fmul_rn(ok ? __ldg(rho_cot + lin) : 0.f, val);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        d[q][0] = __fadd_rn(d[q][0], __fmul_rn(t[q], gx));
        d[q][1] = __fadd_rn(d[q][1], __fmul_rn(t[q], gy));
        d[q][2] = __fadd_rn(d[q][2], __fmul_rn(t[q], gz));
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a = div_cell(d[0][j], P.cell, P.inv_cell);
      a = __fadd_rn(a, div_cell(d[1][j], P.cell, P.inv_cell));
      a = __fadd_rn(a, div_cell(d[2][j], P.cell, P.inv_cell));
      a = __fadd_rn(a, div_cell(d[3][j], P.cell, P.inv_cell));
      st_p(alpha + 3 * p + j, a, cs);
    }
    if (acc) {
      st_p(acc + 3 * p + 0, a0, cs);
      st_p(acc + 3 * p + 1, a1, cs);
      st_p(acc + 3 * p + 2, a2, cs);
    }
  }
}

// ---------------------------------------------------------------------------------------
static int fast_params(const pmwd_cic_desc* d, FastParams* P) {
  PMWD_REQUIRE(d && d->dim == 3, "fast path is 3-D");
  PMWD_REQUIRE(d->pmid_bytes == 2, "fast path needs int16 pmid");
  PMWD_REQUIRE(!d->general, "fast path needs cell_size=None");
  PMWD_REQUIRE(cic_is_fast(d), "fast path needs offset (k*cell, 0, 0), mesh == conf.mesh_shape in y, z");
  P->n = d->ptcl_num;
  P->nx = d->wrap_shape[0];
  P->ny = d->wrap_shape[1];
  P->nz = d->wrap_shape[2];
  P->nx_ext = d->mesh_shape[0];
  P->cell = (float)d->cell_size;
  {
    int e = 0;
    const float m = frexpf(P->cell, &e);
    P->inv_cell = (m == 0.5f && e > -100 && e < 100 && !getenv("PMWD_CIC_DIV")) ? 1.f / P->cell : 0.f;
    const char* c = getenv("PMWD_PTCL_CS");
    P->cs = c ? atoi(c) : 1;
  }
  P->xoff = slab_xoff(d);
  return PMWD_OK;
}

// x offset in whole cells: divmod(b12, a1) with a1 the float32 cell size (pm_util.py:120)
// must have zero remainder for the fast path.
int slab_xoff(const pmwd_cic_desc* d) {
  double a1 = (double)(float)d->cell_size;
  double q = floor(d->offset[0] / a1);
  int off = (int)q % d->wrap_shape[0];
  return off < 0 ? off + d->wrap_shape[0] : off;
}

bool cic_is_fast(const pmwd_cic_desc* d) {
  if (!d || d->dim != 3 || d->pmid_bytes != 2 || d->general) return false;
  for (int a = 0; a < 3; ++a) if (d->wrap_shape[a] <= 0 || d->mesh_shape[a] <= 0) return false;
  if (d->offset[1] != 0.0 || d->offset[2] != 0.0) return false;
  if (d->mesh_shape[1] != d->wrap_shape[1] || d->mesh_shape[2] != d->wrap_shape[2]) return false;
  // mesh_shape[0] may exceed wrap_shape[0] (slab + halos wider than the box on 2 ranks): every
  // global plane then maps to its first local copy, consistently for scatter and gather
  double a1 = (double)(float)d->cell_size;
  double q = floor(d->offset[0] / a1);
  if (d->offset[0] - q * a1 != 0.0) return false;            // whole cells only
  return true;
}

bool cic_is_full_mesh(const pmwd_cic_desc* d) {
  return cic_is_fast(d) && d->offset[0] == 0.0 && d->mesh_shape[0] == d->wrap_shape[0];
}

int scatter_fast(cudaStream_t st, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                 const float* val, float val_scalar, int nch, float* m0, float* m1, float* m2) {
  FastParams P;
  int rc = fast_params(d, &P);
  if (rc) return rc;
  if (P.n == 0) return PMWD_OK;
  const int block = 256;
  int grid = grid_for(P.n, block, 8);
  if (nch == 1)
    scatter_fast_kernel<1><<<grid, block, 0, st>>>(P, (const short*)pmid, disp, val, val_scalar,
                                                   m0, nullptr, nullptr);
  else if (nch == 3)
    scatter_fast_kernel<3><<<grid, block, 0, st>>>(P, (const short*)pmid, disp, val, val_scalar,
                                                   m0, m1, m2);
  else
    PMWD_REQUIRE(false, "scatter_fast supports 1 or 3 channels");
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

int gather3_fast(cudaStream_t st, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                 const float* f0, const float* f1, const float* f2, float* acc, float* vel,
                 float K, const float* next_kd, float* disp_rw) {
  FastParams P;
  int rc = fast_params(d, &P);
  if (rc) return rc;
  if (P.n == 0) return PMWD_OK;
  const int block = 256;
  int grid = grid_for(P.n, block, 8);
  if (vel && next_kd && disp_rw)
    gather3_kernel<2><<<grid, block, 0, st>>>(P, (const short*)pmid, disp, f0, f1, f2, acc, vel, K,
                                              next_kd[0], next_kd[1], disp_rw);
  else if (vel)
    gather3_kernel<1><<<grid, block, 0, st>>>(P, (const short*)pmid, disp, f0, f1, f2, acc, vel, K, 0.f, 0.f, nullptr);
  else
    gather3_kernel<0><<<grid, block, 0, st>>>(P, (const short*)pmid, disp, f0, f1, f2, acc, nullptr, 0.f, 0.f, 0.f, nullptr);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

int force_adj_gather(cudaStream_t st, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                     const float* f0, const float* f1, const float* f2, const float* rho_cot,
                     const float* pi, float val, float* alpha, float* acc) {
  FastParams P;
  int rc = fast_params(d, &P);
  if (rc) return rc;
  if (P.n == 0) return PMWD_OK;
  const int block = 256;
  int grid = grid_for(P.n, block, 8);
  force_adj_gather_kernel<<<grid, block, 0, st>>>(P, (const short*)pmid, disp, f0, f1, f2, rho_cot,
                                                  pi, val, alpha, acc);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

}  // namespace pmwd
