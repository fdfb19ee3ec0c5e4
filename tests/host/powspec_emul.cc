// CPU emulation of the power-spectrum binning kernel (pmwd_b200/csrc/powspec.cu): the same
// per-mode function (powspec.cuh), applied to every mode of a half spectrum in turn.
//   g++ -O2 -std=c++17 -shared -fPIC -Ipmwd_b200/csrc tests/host/powspec_emul.cc -o libpowspec_emul.so
#include "powspec.cuh"

extern "C" void ps_emul(const int* shape, const float* f, const float* g, int has_deconv, double deconv,
                        const double* edges, int nedges, int right, double* out) {
  const int nx = shape[0], ny = shape[1], nz = shape[2], nzc = nz / 2 + 1, nb = nedges + 1;
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j)
      for (int l = 0; l < nzc; ++l) {
        const long q = ((long)i * ny + j) * nzc + l;
        const pmwd::ps::Mode m =
            pmwd::ps::mode(i, j, l, nx, ny, nz, f[2 * q], f[2 * q + 1], g != nullptr, g ? g[2 * q] : 0.f,
                           g ? g[2 * q + 1] : 0.f, has_deconv != 0, (float)deconv, edges, nedges, right != 0);
        out[m.bin] += m.kN;
        out[nb + m.bin] += m.pr;
        out[2 * nb + m.bin] += m.pi;
        out[3 * nb + m.bin] += m.N;
      }
}

extern "C" void ps_weight_emul(const int* shape, const float* f, int has_deconv, double deconv, const double* edges,
                               int nedges, int right, const double* wbin, float* out) {
  const int nx = shape[0], ny = shape[1], nz = shape[2], nzc = nz / 2 + 1;
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j)
      for (int l = 0; l < nzc; ++l) {
        const long q = ((long)i * ny + j) * nzc + l;
        const float w = pmwd::ps::weight(i, j, l, nx, ny, nz, has_deconv != 0, (float)deconv, edges, nedges,
                                         right != 0, wbin);
        out[2 * q] = f[2 * q] * w;
        out[2 * q + 1] = f[2 * q + 1] * w;
      }
}
