// Cloud-in-cell stencil: device-side restatement of what the reference's enmesh computes
// per particle (pmwd/pm_util.py:33-156), without materialising the (N, 2^dim, dim) index /
// weight tensors in HBM.  All float32 arithmetic is written with explicit _rn intrinsics in
// the reference's operation order so that results are bitwise those of an IEEE float32
// evaluation of the reference expressions (the library is also built with --fmad=false).
#pragma once
#include "common.cuh"

namespace pmwd {

struct CicParams {
  int dim;
  int nchan;
  int64_t ptcl_num;
  int wrap[3];        // conf.mesh_shape (periodic wrap, enmesh s1)
  int shape[3];       // target mesh spatial shape (enmesh s2)
  int64_t stride[3];  // element strides of the target mesh (without channel factor)
  // fast branch (pm_util.py:119-136)
  float cell;         // a1 as float32
  int ioff[3];        // i12 = floor(b12 / a1)
  float doff[3];      // d12 = b12 - i12 * a1 (float64 divmod, cast to float32)
  // general float64 branch (pm_util.py:99-118)
  double a1, a2;
  double b12[3];
  float inv_cell_out; // unused placeholder to keep alignment explicit
  float cell_out;     // cell size that disp_cot is divided by (cell_size2 if general else cell)
};

int make_cic_params(const pmwd_cic_desc* d, CicParams* p);

__device__ __forceinline__ float sign_neg(float d) {
  // jnp.sign(-d): 0 at d == 0 (pm_util.py:144)
  return d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f);
}

__device__ __forceinline__ int wrap_index(int i, int n) {
  i %= n;
  return i < 0 ? i + n : i;
}

// Python-style float modulo and numpy/JAX floor_divide for doubles.
__device__ __forceinline__ double py_mod(double x, double m) {
  double r = fmod(x, m);
  if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
  return r;
}
__device__ __forceinline__ double py_floordiv(double x, double m) {
  double r = fmod(x, m);
  double div = (x - r) / m;
  if (r != 0.0 && ((r < 0.0) != (m < 0.0))) div -= 1.0;
  return rint(div);  // div is integral up to rounding
}

// Per-axis stencil data for one particle: two neighbour indices (already wrapped, -1 if
// dropped), weights 1-|d| and signs sign(-d).
struct Axis {
  int idx[2];
  float w[2];
  float s[2];
};

template <bool GENERAL, bool GRAD>
__device__ __forceinline__ Axis cic_axis(const CicParams& P, int ax, int pmid, float disp) {
  Axis a;
  if (!GENERAL) {
    int i1 = pmid - P.ioff[ax];
    float d1 = __fsub_rn(disp, P.doff[ax]);
    float t = __fdiv_rn(d1, P.cell);
    int i0 = (int)floorf(t);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      float d = __fsub_rn(t, (float)(i0 + b));
      a.w[b] = __fsub_rn(1.f, fabsf(d));
      if (GRAD) a.s[b] = sign_neg(d);
      int i = wrap_index(i1 + i0 + b, P.wrap[ax]);
      a.idx[b] = i < P.shape[ax] ? i : -1;
    }
  } else {
    double Pp = (double)pmid * P.a1 + (double)disp - P.b12[ax];
    double L = (double)P.wrap[ax] * P.a1;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      double x = Pp + (double)b * P.a2;
      x = py_mod(x, L);
      double q = py_floordiv(x, P.a2);
      double d2 = Pp - q * P.a2;
      d2 -= rint(d2 / L) * L;
      float d = __fdiv_rn((float)d2, (float)P.a2);
      a.w[b] = __fsub_rn(1.f, fabsf(d));
      if (GRAD) a.s[b] = sign_neg(d);
      long long i = (long long)q;
      a.idx[b] = (i >= 0 && i < P.shape[ax]) ? (int)i : -1;
    }
  }
  return a;
}

template <typename PM>
__device__ __forceinline__ int load_pmid(const void* pmid, int64_t p, int dim, int ax) {
  return (int)reinterpret_cast<const PM*>(pmid)[p * dim + ax];
}

}  // namespace pmwd
