// Deterministic, cell-sorted CIC scatter (PMWD_SCATTER_DETERMINISTIC).
//
// The reference's GPU scatter-add (pmwd/scatter.py:80) is an unordered float32 atomic
// accumulation and therefore not reproducible run to run (docs/papers/adjoint/adjoint.tex:1440-1480).
// This mode makes the deposit bitwise reproducible:
//   1. key[p]  = linear index of particle p's base cell (the n=0 CIC neighbour);
//   2. stable LSD radix sort of (key, p) -- ties keep ascending p  (cub::DeviceRadixSort);
//   3. run boundaries -> cell_start / cell_end;
//   4. one thread per mesh cell sums, in a FIXED order (neighbour offset n = 0..7, then
//      ascending particle id), the contributions of the particles based in its 8 lower
//      neighbours, and adds the result to the mesh with a plain store -- no atomics.
// Every mesh cell is written exactly once with coalesced stores.
#include <cub/device/device_radix_sort.cuh>

#include "cic.cuh"

namespace pmwd {

struct DetParams {
  int64_t n;
  int nx, ny, nz;
  float cell;
};

__device__ __forceinline__ int base_index(int pm, float disp, float cell, int n) {
  float t = __fdiv_rn(disp, cell);
  return wrap_index(pm + (int)floorf(t), n);
}

__global__ void __launch_bounds__(256)
det_keys_kernel(DetParams P, const short* __restrict__ pmid, const float* __restrict__ disp,
                uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P.n;
       p += (int64_t)gridDim.x * blockDim.x) {
    int ix = base_index(pmid[3 * p + 0], disp[3 * p + 0], P.cell, P.nx);
    int iy = base_index(pmid[3 * p + 1], disp[3 * p + 1], P.cell, P.ny);
    int iz = base_index(pmid[3 * p + 2], disp[3 * p + 2], P.cell, P.nz);
    keys[p] = (uint32_t)(((int64_t)ix * P.ny + iy) * P.nz + iz);
    vals[p] = (uint32_t)p;
  }
}

__global__ void __launch_bounds__(256)
det_bounds_kernel(int64_t n, const uint32_t* __restrict__ keys, uint32_t* __restrict__ cstart,
                  uint32_t* __restrict__ cend) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t k = keys[i];
    if (i == 0 || keys[i - 1] != k) cstart[k] = (uint32_t)i;
    if (i == n - 1 || keys[i + 1] != k) cend[k] = (uint32_t)(i + 1);
  }
}

template <int NCH>
__global__ void __launch_bounds__(256)
det_accum_kernel(DetParams P, const short* __restrict__ pmid, const float* __restrict__ disp,
                 const float* __restrict__ val, float val_scalar,
                 const uint32_t* __restrict__ order, const uint32_t* __restrict__ cstart,
                 const uint32_t* __restrict__ cend, float* __restrict__ m0,
                 float* __restrict__ m1, float* __restrict__ m2) {
  const int64_t ncell = (int64_t)P.nx * P.ny * P.nz;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncell;
       c += (int64_t)gridDim.x * blockDim.x) {
    int z = (int)(c % P.nz);
    int64_t r = c / P.nz;
    int y = (int)(r % P.ny);
    int x = (int)(r / P.ny);
    float acc[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) acc[ch] = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int bx = n & 1, by = (n >> 1) & 1, bz = n >> 2;
      int xb = x - bx; if (xb < 0) xb += P.nx;
      int yb = y - by; if (yb < 0) yb += P.ny;
      int zb = z - bz; if (zb < 0) zb += P.nz;
      int64_t b = ((int64_t)xb * P.ny + yb) * P.nz + zb;
      uint32_t lo = __ldg(cstart + b), hi = __ldg(cend + b);
      for (uint32_t i = lo; i < hi; ++i) {
        int64_t p = __ldg(order + i);
        float w = 1.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          float t = __fdiv_rn(__ldg(disp + 3 * p + a), P.cell);
          int i0 = (int)floorf(t);
          int bit = a == 0 ? bx : (a == 1 ? by : bz);
          float d = __fsub_rn(t, (float)(i0 + bit));
          float wa = __fsub_rn(1.f, fabsf(d));
          w = a == 0 ? wa : __fmul_rn(w, wa);
        }
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          float v = val ? __ldg(val + NCH * p + ch) : val_scalar;
          acc[ch] = __fadd_rn(acc[ch], __fmul_rn(v, w));
        }
      }
    }
    m0[c] = __fadd_rn(m0[c], acc[0]);
    if (NCH > 1) m1[c] = __fadd_rn(m1[c], acc[1]);
    if (NCH > 2) m2[c] = __fadd_rn(m2[c], acc[2]);
  }
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct DetLayout {
  size_t keys_in, vals_in, keys_out, vals_out, cstart, cend, cub, total, cub_bytes;
};

static int det_layout(int64_t n, int64_t ncell, int end_bit, DetLayout* L) {
  size_t cub_bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr,
                                                  (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                                  (uint32_t*)nullptr, n, 0, end_bit);
  if (e != cudaSuccess) { set_error("cub temp query failed: %s", cudaGetErrorString(e)); return (int)e; }
  size_t off = 0;
  L->keys_in = off;  off += align_up((size_t)n * 4);
  L->vals_in = off;  off += align_up((size_t)n * 4);
  L->keys_out = off; off += align_up((size_t)n * 4);
  L->vals_out = off; off += align_up((size_t)n * 4);
  L->cstart = off;   off += align_up((size_t)ncell * 4);
  L->cend = off;     off += align_up((size_t)ncell * 4);
  L->cub = off;      off += align_up(cub_bytes);
  L->cub_bytes = cub_bytes;
  L->total = off;
  return PMWD_OK;
}

static int ilog2_ceil(int64_t x) { int b = 0; while (((int64_t)1 << b) < x) ++b; return b; }

size_t scatter_det_scratch_bytes(const pmwd_cic_desc* d) {
  if (!d || d->dim != 3) return 0;
  int64_t ncell = (int64_t)d->mesh_shape[0] * d->mesh_shape[1] * d->mesh_shape[2];
  DetLayout L;
  if (det_layout(d->ptcl_num, ncell, ilog2_ceil(ncell), &L)) return 0;
  return L.total;
}

int scatter_det(cudaStream_t st, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                const float* val, float val_scalar, int nch, float* m0, float* m1, float* m2,
                void* scratch, size_t scratch_bytes) {
  PMWD_REQUIRE(d && d->dim == 3 && d->pmid_bytes == 2 && !d->general,
               "deterministic scatter supports the 3-D int16 fast path");
  for (int a = 0; a < 3; ++a)
    PMWD_REQUIRE(d->offset[a] == 0.0 && d->mesh_shape[a] == d->wrap_shape[a],
                 "deterministic scatter needs offset 0 and mesh == conf.mesh_shape");
  PMWD_REQUIRE(nch == 1 || nch == 3, "deterministic scatter supports 1 or 3 channels");
  DetParams P;
  P.n = d->ptcl_num;
  P.nx = d->mesh_shape[0]; P.ny = d->mesh_shape[1]; P.nz = d->mesh_shape[2];
  P.cell = (float)d->cell_size;
  int64_t ncell = (int64_t)P.nx * P.ny * P.nz;
  PMWD_REQUIRE(ncell <= ((int64_t)1 << 32) && P.n < ((int64_t)1 << 32),
               "deterministic scatter needs mesh_size <= 2^32 and ptcl_num < 2^32 per device");
  if (P.n == 0) return PMWD_OK;
  int end_bit = ilog2_ceil(ncell);
  if (end_bit < 1) end_bit = 1;
  DetLayout L;
  int rc = det_layout(P.n, ncell, end_bit, &L);
  if (rc) return rc;
  if (!scratch || scratch_bytes < L.total) {
    set_error("deterministic scatter needs %zu bytes of scratch, got %zu", L.total, scratch_bytes);
    return PMWD_ENOMEM;
  }
  char* base = (char*)scratch;
  uint32_t* keys_in = (uint32_t*)(base + L.keys_in);
  uint32_t* vals_in = (uint32_t*)(base + L.vals_in);
  uint32_t* keys_out = (uint32_t*)(base + L.keys_out);
  uint32_t* vals_out = (uint32_t*)(base + L.vals_out);
  uint32_t* cstart = (uint32_t*)(base + L.cstart);
  uint32_t* cend = (uint32_t*)(base + L.cend);

  const int block = 256;
  det_keys_kernel<<<grid_for(P.n, block, 8), block, 0, st>>>(P, (const short*)pmid, disp, keys_in, vals_in);
  PMWD_LAUNCH_CHECK();
  size_t cub_bytes = L.cub_bytes;
  PMWD_CUDA_TRY(cub::DeviceRadixSort::SortPairs(base + L.cub, cub_bytes, keys_in, keys_out, vals_in,
                                                vals_out, P.n, 0, end_bit, st));
  // cstart and cend are adjacent: one memset clears both (empty cells: start == end == 0)
  PMWD_CUDA_TRY(cudaMemsetAsync(cstart, 0, (size_t)((char*)cend - (char*)cstart) + (size_t)ncell * 4, st));
  det_bounds_kernel<<<grid_for(P.n, block, 8), block, 0, st>>>(P.n, keys_out, cstart, cend);
  PMWD_LAUNCH_CHECK();
  int grid = grid_for(ncell, block, 8);
  if (nch == 1)
    det_accum_kernel<1><<<grid, block, 0, st>>>(P, (const short*)pmid, disp, val, val_scalar, vals_out,
                                                cstart, cend, m0, nullptr, nullptr);
  else
    det_accum_kernel<3><<<grid, block, 0, st>>>(P, (const short*)pmid, disp, val, val_scalar, vals_out,
                                                cstart, cend, m0, m1, m2);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

}  // namespace pmwd
