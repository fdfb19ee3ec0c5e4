/* C restatement of the reference's particle-mesh step for the CPU baseline: TEST INFRASTRUCTURE ONLY
 * (used by tests/, bench.py's cpu_baseline / --impl reference legs; never by the product).
 *
 * Same float32 arithmetic, operation by operation, as the NumPy oracle (oracle/pm.py,
 * oracle/gravity.py, oracle/nbody.py), which restates
 *   enmesh fast branch      pmwd/pm_util.py:119-138,154
 *   _scatter_chunk          pmwd/scatter.py:60-83
 *   _gather_chunk           pmwd/gather.py:58-77
 *   laplace / neg_grad      pmwd/gravity.py:9-16,37-44
 *   kick / drift            pmwd/nbody.py:39-46,70-77
 * and, for the reverse-time adjoint (what jax.vjp(gravity) evaluates at pmwd/nbody.py:108-118):
 *   _gather_chunk_adj       pmwd/gather.py:96-116   (cpm_scatter_val, cpm_grad_gather)
 *   _scatter_chunk_adj      pmwd/scatter.py:102-121 (cpm_grad_gather)
 *   laplace_bwd + transposed neg_grad  pmwd/gravity.py:23-44 (cpm_kspace_force_adj)
 * but with plain loops and OpenMP over particles / mesh rows, so that the CPU baseline uses the
 * host cores the way a compiled CPU backend would.  The FFTs stay in scipy (pocketfft, threaded).
 * 3-D, offset 0, default cell size only (the step's own use); the general paths live in NumPy.
 *
 * Build: gcc -O3 -fopenmp -fno-fast-math -ffp-contract=off -shared -fPIC cpm.c -o _build/libcpm.so
 * (-ffp-contract=off: no fused multiply-adds, NumPy rounds every product). */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* floor-mod for the periodic wrap (pm_util.py:135-136) */
static inline int wrap(int i, int n) {
  int r = i % n;
  return r < 0 ? r + n : r;
}

/* pm_util.py:129-138 along one axis: t = disp / cell (float32 divide); i = floor(t) + bit;
 * d = t - i; frac = 1 - |d|; index = (pmid + i) mod n */
static inline void axis(int16_t pmid, float disp, float cell, int n, int ix[2], float w[2]) {
  const float t = disp / cell;
  const int i0 = (int)floorf(t);
  for (int b = 0; b < 2; ++b) {
    const int i = i0 + b;
    const float d = t - (float)i;
    w[b] = 1.0f - fabsf(d);
    /* int16 arithmetic in the reference: the sum wraps like int16 before the modulo */
    ix[b] = wrap((int)(int16_t)(pmid + (int16_t)i), n);
  }
}

int cpm_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* mesh[ind] += val * w, neighbours in the reference's order n = bx + 2 by + 4 bz, weight
 * ((1 * wx) * wy) * wz.  threads == 1: updates applied in (particle, neighbour) order, bit-exact
 * with np.add.at; threads > 1: particles split in contiguous ranges, atomic float adds (summation
 * order, hence the last bits, differ -- exactly like the reference's XLA scatter-add). */
void cpm_scatter(int64_t n, const int16_t* pmid, const float* disp, float val, float cell,
                 const int32_t* shape, float* mesh, int threads) {
  const int nx = shape[0], ny = shape[1], nz = shape[2];
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
  for (int64_t p = 0; p < n; ++p) {
    int ix[2], iy[2], iz[2];
    float wx[2], wy[2], wz[2];
    axis(pmid[3 * p + 0], disp[3 * p + 0], cell, nx, ix, wx);
    axis(pmid[3 * p + 1], disp[3 * p + 1], cell, ny, iy, wy);
    axis(pmid[3 * p + 2], disp[3 * p + 2], cell, nz, iz, wz);
    for (int nb = 0; nb < 8; ++nb) {
      const int bx = nb & 1, by = (nb >> 1) & 1, bz = (nb >> 2) & 1;
      const float w = (wx[bx] * wy[by]) * wz[bz];
      const float upd = val * w;
      float* cell_ptr = mesh + ((int64_t)ix[bx] * ny + iy[by]) * nz + iz[bz];
      if (threads > 1) {
#pragma omp atomic
        *cell_ptr += upd;
      } else {
        *cell_ptr += upd;
      }
    }
  }
}

/* out[p][c] = sum_n mesh_c[ind] * w, neighbour sum sequential n = 0..7 from 0 (gather.py:75 with
 * val = 0), for the nmesh (1..3) meshes of a force in one pass over the particles; out has row
 * stride 3 (the stacked acc of gravity.py:70) when nmesh == 3, 1 otherwise. */
void cpm_gather(int64_t n, const int16_t* pmid, const float* disp, float cell, const int32_t* shape,
                int nmesh, const float* m0, const float* m1, const float* m2, float* out, int threads) {
  const int nx = shape[0], ny = shape[1], nz = shape[2];
  const float* m[3] = {m0, m1, m2};
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
  for (int64_t p = 0; p < n; ++p) {
    int ix[2], iy[2], iz[2];
    float wx[2], wy[2], wz[2];
    axis(pmid[3 * p + 0], disp[3 * p + 0], cell, nx, ix, wx);
    axis(pmid[3 * p + 1], disp[3 * p + 1], cell, ny, iy, wy);
    axis(pmid[3 * p + 2], disp[3 * p + 2], cell, nz, iz, wz);
    float acc[3] = {0.f, 0.f, 0.f};
    for (int nb = 0; nb < 8; ++nb) {
      const int bx = nb & 1, by = (nb >> 1) & 1, bz = (nb >> 2) & 1;
      const float w = (wx[bx] * wy[by]) * wz[bz];
      const int64_t lin = ((int64_t)ix[bx] * ny + iy[by]) * nz + iz[bz];
      for (int c = 0; c < nmesh; ++c) acc[c] = acc[c] + m[c][lin] * w;
    }
    for (int c = 0; c < nmesh; ++c) out[(int64_t)nmesh * p + c] = 0.0f + acc[c];
  }
}

/* angular wavenumber of index i along an axis of n cells: fftfreq(n) * period in float64, cast
 * to float32 (pm_util.py:159-199); last = rfftfreq */
static inline float kval(int i, int n, double period, int last) {
  const int f = last ? i : (i < (n + 1) / 2 ? i : i - n);
  return (float)(((double)f / (double)n) * period);
}

/* gravity.py:58-62 on the half spectrum spec[nx][ny][nz/2+1] (complex64, interleaved):
 * pot = where(k2 != 0, -src / k2, 0); g_a = (-i k_a, zero on the Nyquist planes) * pot.
 * k2 = (kx^2 + ky^2) + kz^2 in float32, the sum order of `sum(k**2 for k in kvec)`. */
void cpm_kspace_force(const int32_t* shape, double spacing, const float* spec, float* g0, float* g1,
                      float* g2, int threads) {
  const int nx = shape[0], ny = shape[1], nz = shape[2], nzc = nz / 2 + 1;
  const double pi = 3.141592653589793238462643383279502884;
  const double period = 2.0 * pi / spacing;
  const float nyq = (float)(pi / spacing);
  const float eps = nyq * 1.1920928955078125e-07f;
  float* g[3] = {g0, g1, g2};
#pragma omp parallel for schedule(static) collapse(2) num_threads(threads) if (threads > 1)
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      const float kx = kval(i, nx, period, 0), ky = kval(j, ny, period, 0);
      const float kxy = kx * kx + ky * ky;
      for (int l = 0; l < nzc; ++l) {
        const float kz = kval(l, nz, period, 1);
        const float k2 = kxy + kz * kz;
        const int64_t q = 2 * (((int64_t)i * ny + j) * nzc + l);
        float pr = 0.f, pim = 0.f;
        if (k2 != 0.f) {
          /* NumPy divides complex64 by the (promoted) real k2 with Smith's algorithm, which for a
           * zero imaginary divisor reduces to a multiplication by the rounded reciprocal */
          const float scl = 1.0f / k2;
          pr = -spec[q] * scl;
          pim = -spec[q + 1] * scl;
        }
        const float k[3] = {kx, ky, kz};
        for (int a = 0; a < 3; ++a) {
          /* neg_ik = where(| |k| - nyq | <= eps, 0, -1j k);  (0 - i k)(pr + i pim) */
          if (fabsf(fabsf(k[a]) - nyq) <= eps) {
            g[a][q] = 0.f;
            g[a][q + 1] = 0.f;
          } else {
            g[a][q] = k[a] * pim;
            g[a][q + 1] = -(k[a] * pr);
          }
        }
      }
    }
}

/* dens = (dens - 1) * scale (gravity.py:52-54), in place */
void cpm_contrast(int64_t nm, float* dens, float scale, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
  for (int64_t i = 0; i < nm; ++i) dens[i] = (dens[i] - 1.0f) * scale;
}

/* y += x * f (kick: vel += acc * K, nbody.py:70-77; drift: disp += vel * D, nbody.py:39-46) */
void cpm_axpy(int64_t n, float* y, const float* x, float f, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
  for (int64_t i = 0; i < n; ++i) y[i] = y[i] + x[i] * f;
}

/* ---------------------------------------------------------------------------------------------
 * Reverse-time adjoint pieces (oracle/gravity.py:gravity_vjp, oracle/pm.py:gather_adj/scatter_adj)
 * ------------------------------------------------------------------------------------------- */

/* mesh[ind] += val[p * stride] * w  (gather.py:113: mesh_cot.at[ind].add(val_cot * frac)), i.e.
 * cpm_scatter with a per-particle value; same update order / atomics rule. */
void cpm_scatter_val(int64_t n, const int16_t* pmid, const float* disp, const float* val, int stride,
                     float cell, const int32_t* shape, float* mesh, int threads) {
  const int nx = shape[0], ny = shape[1], nz = shape[2];
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
  for (int64_t p = 0; p < n; ++p) {
    int ix[2], iy[2], iz[2];
    float wx[2], wy[2], wz[2];
    axis(pmid[3 * p + 0], disp[3 * p + 0], cell, nx, ix, wx);
    axis(pmid[3 * p + 1], disp[3 * p + 1], cell, ny, iy, wy);
    axis(pmid[3 * p + 2], disp[3 * p + 2], cell, nz, iz, wz);
    const float v = val[(int64_t)stride * p];
    for (int nb = 0; nb < 8; ++nb) {
      const int bx = nb & 1, by = (nb >> 1) & 1, bz = (nb >> 2) & 1;
      const float w = (wx[bx] * wy[by]) * wz[bz];
      const float upd = v * w;
      float* cell_ptr = mesh + ((int64_t)ix[bx] * ny + iy[by]) * nz + iz[bz];
      if (threads > 1) {
#pragma omp atomic
        *cell_ptr += upd;
      } else {
        *cell_ptr += upd;
      }
    }
  }
}

/* jnp.sign(-d) (pm_util.py:144) from the weight's own d: recomputed like axis() */
static inline void axis_grad(int16_t pmid, float disp, float cell, int n, int ix[2], float w[2], float s[2]) {
  const float t = disp / cell;
  const int i0 = (int)floorf(t);
  for (int b = 0; b < 2; ++b) {
    const int i = i0 + b;
    const float d = t - (float)i;
    w[b] = 1.0f - fabsf(d);
    s[b] = d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f);
    ix[b] = wrap((int)(int16_t)(pmid + (int16_t)i), n);
  }
}

/* out[p][j] += ( sum_n (x_p * mesh[ind_n]) * g_nj ) / cell, j = 0..2, neighbour sum sequential
 * from 0; g_nj = sign(-d_j) * prod_{m != j} w_m in axis order (j+1.., 0..j-1) (pm_util.py:143-150).
 * x_p = x[p * stride] (gather.py:106-110, the val_cot of one force component) or the scalar xs
 * when x == NULL (scatter.py:112-116, val = N_m / N_p). */
void cpm_grad_gather(int64_t n, const int16_t* pmid, const float* disp, float cell, const int32_t* shape,
                     const float* mesh, const float* x, int stride, float xs, float* out, int threads) {
  const int nx = shape[0], ny = shape[1], nz = shape[2];
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
  for (int64_t p = 0; p < n; ++p) {
    int ix[2], iy[2], iz[2];
    float wx[2], wy[2], wz[2], sx[2], sy[2], sz[2];
    axis_grad(pmid[3 * p + 0], disp[3 * p + 0], cell, nx, ix, wx, sx);
    axis_grad(pmid[3 * p + 1], disp[3 * p + 1], cell, ny, iy, wy, sy);
    axis_grad(pmid[3 * p + 2], disp[3 * p + 2], cell, nz, iz, wz, sz);
    const float xp = x ? x[(int64_t)stride * p] : xs;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    for (int nb = 0; nb < 8; ++nb) {
      const int bx = nb & 1, by = (nb >> 1) & 1, bz = (nb >> 2) & 1;
      const float gx = sx[bx] * (wy[by] * wz[bz]);
      const float gy = sy[by] * (wz[bz] * wx[bx]);
      const float gz = sz[bz] * (wx[bx] * wy[by]);
      const float t = xp * mesh[((int64_t)ix[bx] * ny + iy[by]) * nz + iz[bz]];
      d0 = d0 + t * gx;
      d1 = d1 + t * gy;
      d2 = d2 + t * gz;
    }
    out[3 * p + 0] = out[3 * p + 0] + d0 / cell;
    out[3 * p + 1] = out[3 * p + 1] + d1 / cell;
    out[3 * p + 2] = out[3 * p + 2] + d2 / cell;
  }
}

/* out = ((0 - g_0(s0)) - g_1(s1)) - g_2(s2) with g_a(s) = neg_grad(k_a, laplace(kvec, s)), the
 * accumulation of oracle/gravity.py:gravity_vjp (A_a^T = -A_a; gravity.py:23-44 transposed), on
 * half spectra [nx][ny][nz/2+1] complex64 interleaved. */
void cpm_kspace_force_adj(const int32_t* shape, double spacing, const float* s0, const float* s1,
                          const float* s2, float* out, int threads) {
  const int nx = shape[0], ny = shape[1], nz = shape[2], nzc = nz / 2 + 1;
  const double pi = 3.141592653589793238462643383279502884;
  const double period = 2.0 * pi / spacing;
  const float nyq = (float)(pi / spacing);
  const float eps = nyq * 1.1920928955078125e-07f;
  const float* s[3] = {s0, s1, s2};
#pragma omp parallel for schedule(static) collapse(2) num_threads(threads) if (threads > 1)
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) {
      const float kx = kval(i, nx, period, 0), ky = kval(j, ny, period, 0);
      const float kxy = kx * kx + ky * ky;
      for (int l = 0; l < nzc; ++l) {
        const float kz = kval(l, nz, period, 1);
        const float k2 = kxy + kz * kz;
        const int64_t q = 2 * (((int64_t)i * ny + j) * nzc + l);
        const float k[3] = {kx, ky, kz};
        float re = 0.f, im = 0.f;
        for (int a = 0; a < 3; ++a) {
          float pr = 0.f, pim = 0.f;
          if (k2 != 0.f) {
            const float scl = 1.0f / k2;
            pr = -s[a][q] * scl;
            pim = -s[a][q + 1] * scl;
          }
          float gr = 0.f, gi = 0.f;
          if (!(fabsf(fabsf(k[a]) - nyq) <= eps)) {
            gr = k[a] * pim;
            gi = -(k[a] * pr);
          }
          re = re - gr;
          im = im - gi;
        }
        out[q] = re;
        out[q + 1] = im;
      }
    }
}

/* sum_i x[i] * y[i] accumulated in float64 (the reference sums in float32; this is the threaded
 * stand-in used only for the cosmology dot products of the CPU baseline) */
double cpm_dot(int64_t n, const float* x, const float* y, int threads) {
  double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc) num_threads(threads) if (threads > 1)
  for (int64_t i = 0; i < n; ++i) acc += (double)(x[i] * y[i]);
  return acc;
}
