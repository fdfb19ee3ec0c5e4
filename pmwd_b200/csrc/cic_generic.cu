// General CIC scatter / gather and their VJPs: any dim in {1,2,3}, any channel count
// (channels last, interleaved), offset / cell_size (float64 enmesh branch), int8/16/32 pmid.
// Replaces pmwd/scatter.py:33-148 and pmwd/gather.py:33-142 for every call that is not on
// the gravity fast path (which uses the fused kernels in cic_fast.cu).
//
// HBM-bound: one thread per particle, grid-stride, grid = 148 SMs x 8 CTAs.  The
// (N, 2^dim, dim) index/weight tensors of the reference never exist in memory.
#include "cic.cuh"

namespace pmwd {

int make_cic_params(const pmwd_cic_desc* d, CicParams* p) {
  PMWD_REQUIRE(d != nullptr, "null descriptor");
  PMWD_REQUIRE(d->dim >= 1 && d->dim <= 3, "dim must be 1, 2 or 3");
  PMWD_REQUIRE(d->pmid_bytes == 1 || d->pmid_bytes == 2 || d->pmid_bytes == 4,
               "pmid must be int8, int16 or int32");
  PMWD_REQUIRE(d->ptcl_num >= 0, "negative ptcl_num");
  PMWD_REQUIRE(d->nchan >= 1, "nchan must be >= 1");
  PMWD_REQUIRE(d->cell_size > 0, "cell_size must be positive");
  memset(p, 0, sizeof(*p));
  p->dim = d->dim;
  p->nchan = d->nchan;
  p->ptcl_num = d->ptcl_num;
  int64_t stride = 1;
  for (int ax = d->dim - 1; ax >= 0; --ax) {
    PMWD_REQUIRE(d->wrap_shape[ax] > 0 && d->mesh_shape[ax] > 0, "non-positive mesh shape");
    p->wrap[ax] = d->wrap_shape[ax];
    p->shape[ax] = d->mesh_shape[ax];
    p->stride[ax] = stride;
    stride *= d->mesh_shape[ax];
  }
  p->cell = (float)d->cell_size;
  p->a1 = d->cell_size;
  p->a2 = d->general ? d->cell_size2 : d->cell_size;
  if (d->general) PMWD_REQUIRE(d->cell_size2 > 0, "cell_size2 must be positive");
  double a1f = (double)p->cell;  // divmod(b12, a1) with a1 already rounded to float32
  for (int ax = 0; ax < d->dim; ++ax) {
    double b = d->offset[ax];
    double q = floor(b / a1f);
    double r = b - q * a1f;
    // python divmod keeps 0 <= r < a1
    if (r < 0) { r += a1f; q -= 1; }
    if (r >= a1f) { r -= a1f; q += 1; }
    p->ioff[ax] = (int)q;
    p->doff[ax] = (float)r;
    p->b12[ax] = b;
  }
  p->cell_out = d->general ? (float)d->cell_size2 : (float)d->cell_size;
  return PMWD_OK;
}

// mode: 0 scatter, 1 gather, 2 scatter_adj, 3 gather_adj
template <int DIM, typename PM, bool GENERAL, int MODE>
__global__ void __launch_bounds__(256)
cic_generic_kernel(CicParams P, const void* __restrict__ pmid, const float* __restrict__ disp,
                   const float* __restrict__ a_in,   // scatter: val   | gather: val      | scatter_adj: val | gather_adj: val_cot
                   float a_scalar,
                   const float* __restrict__ m_in,   // gather: mesh | scatter_adj: mesh_cot | gather_adj: mesh
                   float* __restrict__ m_out,        // scatter: mesh | gather_adj: mesh_cot (nullable)
                   float* __restrict__ p_out,        // gather: out | *_adj: disp_cot
                   float* __restrict__ p_out2) {     // scatter_adj: val_cot (nullable)
  constexpr int NN = 1 << DIM;
  constexpr bool GRAD = MODE >= 2;
  const int nchan = P.nchan;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P.ptcl_num;
       p += (int64_t)gridDim.x * blockDim.x) {
    Axis ax[DIM];
#pragma unroll
    for (int j = 0; j < DIM; ++j)
      ax[j] = cic_axis<GENERAL, GRAD>(P, j, load_pmid<PM>(pmid, p, DIM, j), disp[p * DIM + j]);

    float dcot[DIM];
#pragma unroll
    for (int j = 0; j < DIM; ++j) dcot[j] = 0.f;

    if (MODE == 1) {  // gather: out = val + sum_n mesh*frac  (sequential n, per channel)
      for (int c = 0; c < nchan; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int n = 0; n < NN; ++n) {
          int64_t lin = 0; bool ok = true; float w = 1.f;
#pragma unroll
          for (int j = 0; j < DIM; ++j) {
            int b = (n >> j) & 1;
            ok = ok && ax[j].idx[b] >= 0;
            lin += (int64_t)ax[j].idx[b] * P.stride[j];
            w = __fmul_rn(w, ax[j].w[b]);
          }
          float m = ok ? __ldg(m_in + lin * nchan + c) : 0.f;
          acc = __fadd_rn(acc, __fmul_rn(m, w));
        }
        float v = a_in ? a_in[p * nchan + c] : a_scalar;
        p_out[p * nchan + c] = __fadd_rn(v, acc);
      }
      continue;
    }

    if (MODE == 2 && p_out2) {
      for (int c = 0; c < nchan; ++c) p_out2[p * nchan + c] = 0.f;
    }

#pragma unroll
    for (int n = 0; n < NN; ++n) {
      int64_t lin = 0; bool ok = true; float w = 1.f;
#pragma unroll
      for (int j = 0; j < DIM; ++j) {
        int b = (n >> j) & 1;
        ok = ok && ax[j].idx[b] >= 0;
        lin += (int64_t)ax[j].idx[b] * P.stride[j];
        w = __fmul_rn(w, ax[j].w[b]);
      }
      if (MODE == 0) {  // scatter: mesh[ind] += val*frac
        if (ok) {
          for (int c = 0; c < nchan; ++c) {
            float v = a_in ? a_in[p * nchan + c] : a_scalar;
            atomicAdd(m_out + lin * nchan + c, __fmul_rn(v, w));
          }
        }
      } else {
        // weight gradients: g_j = sign(-d_j) * prod_{m != j} w_m in the reference's
        // axis order (j+1 .. dim-1, 0 .. j-1)  (pm_util.py:145-149)
        float g[DIM];
#pragma unroll
        for (int j = 0; j < DIM; ++j) {
          float pr = 1.f;
#pragma unroll
          for (int m = j + 1; m < DIM; ++m) pr = __fmul_rn(pr, ax[m].w[(n >> m) & 1]);
#pragma unroll
          for (int m = 0; m < j; ++m) pr = __fmul_rn(pr, ax[m].w[(n >> m) & 1]);
          g[j] = __fmul_rn(ax[j].s[(n >> j) & 1], pr);
        }
        float dot = 0.f;  // sum over channels of mesh(_cot)[ind] * val(_cot)
        for (int c = 0; c < nchan; ++c) {
          float m = ok ? __ldg(m_in + lin * nchan + c) : 0.f;
          float v = a_in ? a_in[p * nchan + c] : a_scalar;
          dot = __fadd_rn(dot, __fmul_rn(m, v));
          if (MODE == 2 && p_out2)
            p_out2[p * nchan + c] = __fadd_rn(p_out2[p * nchan + c], __fmul_rn(m, w));
          if (MODE == 3 && m_out && ok)
            atomicAdd(m_out + lin * nchan + c, __fmul_rn(v, w));
        }
#pragma unroll
        for (int j = 0; j < DIM; ++j) dcot[j] = __fadd_rn(dcot[j], __fmul_rn(dot, g[j]));
      }
    }
    if (GRAD) {
#pragma unroll
      for (int j = 0; j < DIM; ++j) p_out[p * DIM + j] = __fdiv_rn(dcot[j], P.cell_out);
    }
  }
}

template <int DIM, typename PM, bool GENERAL, int MODE>
static int launch1(cudaStream_t st, const CicParams& P, const void* pmid, const float* disp,
                   const float* a_in, float a_scalar, const float* m_in, float* m_out,
                   float* p_out, float* p_out2) {
  if (P.ptcl_num == 0) return PMWD_OK;
  const int block = 256;
  int grid = grid_for(P.ptcl_num, block, 8);
  cic_generic_kernel<DIM, PM, GENERAL, MODE><<<grid, block, 0, st>>>(
      P, pmid, disp, a_in, a_scalar, m_in, m_out, p_out, p_out2);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

template <int DIM, typename PM, int MODE>
static int launch2(cudaStream_t st, const CicParams& P, bool general, const void* pmid,
                   const float* disp, const float* a_in, float a_scalar, const float* m_in,
                   float* m_out, float* p_out, float* p_out2) {
  return general ? launch1<DIM, PM, true, MODE>(st, P, pmid, disp, a_in, a_scalar, m_in, m_out, p_out, p_out2)
                 : launch1<DIM, PM, false, MODE>(st, P, pmid, disp, a_in, a_scalar, m_in, m_out, p_out, p_out2);
}

template <int DIM, int MODE>
static int launch3(cudaStream_t st, const CicParams& P, bool general, int pmid_bytes,
                   const void* pmid, const float* disp, const float* a_in, float a_scalar,
                   const float* m_in, float* m_out, float* p_out, float* p_out2) {
  switch (pmid_bytes) {
    case 1: return launch2<DIM, int8_t, MODE>(st, P, general, pmid, disp, a_in, a_scalar, m_in, m_out, p_out, p_out2);
    case 2: return launch2<DIM, int16_t, MODE>(st, P, general, pmid, disp, a_in, a_scalar, m_in, m_out, p_out, p_out2);
    default: return launch2<DIM, int32_t, MODE>(st, P, general, pmid, disp, a_in, a_scalar, m_in, m_out, p_out, p_out2);
  }
}

template <int MODE>
int cic_generic(cudaStream_t st, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                const float* a_in, float a_scalar, const float* m_in, float* m_out,
                float* p_out, float* p_out2) {
  CicParams P;
  int rc = make_cic_params(d, &P);
  if (rc) return rc;
  bool general = d->general != 0;
  switch (d->dim) {
    case 1: return launch3<1, MODE>(st, P, general, d->pmid_bytes, pmid, disp, a_in, a_scalar, m_in, m_out, p_out, p_out2);
    case 2: return launch3<2, MODE>(st, P, general, d->pmid_bytes, pmid, disp, a_in, a_scalar, m_in, m_out, p_out, p_out2);
    default: return launch3<3, MODE>(st, P, general, d->pmid_bytes, pmid, disp, a_in, a_scalar, m_in, m_out, p_out, p_out2);
  }
}

template int cic_generic<0>(cudaStream_t, const pmwd_cic_desc*, const void*, const float*, const float*, float, const float*, float*, float*, float*);
template int cic_generic<1>(cudaStream_t, const pmwd_cic_desc*, const void*, const float*, const float*, float, const float*, float*, float*, float*);
template int cic_generic<2>(cudaStream_t, const pmwd_cic_desc*, const void*, const float*, const float*, float, const float*, float*, float*, float*);
template int cic_generic<3>(cudaStream_t, const pmwd_cic_desc*, const void*, const float*, const float*, float, const float*, float*, float*, float*);

}  // namespace pmwd
