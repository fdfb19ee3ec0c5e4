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
