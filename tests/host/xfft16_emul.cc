// CPU emulation of the register FFT of pmwd_b200/csrc/xfft16.cuh: every barrier-separated phase
// is run for all threads of a CTA in turn (registers = per-thread arrays), and the result is
// compared with a naive float64 DFT.  Checks the Stockham index arithmetic, the exchange-buffer
// addressing (incl. that no two threads write the same word and every word read was written)
// and the ownership invariant  x = j + (NX/16) e  before and after a transform.
//   g++ -O2 -std=c++17 -I/usr/local/cuda/include -Ipmwd_b200/csrc tests/host/xfft16_emul.cc -o /tmp/xfft16_emul
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "xfft16.cuh"

using namespace pmwd::r16;

template <int NX, bool INV>
static double run() {
  using C = Cfg<NX>;
  const int NT = C::THREADS;
  std::vector<float2> tw(NX);
  for (int n = 0; n < NX; ++n) {
    const double a = -2.0 * M_PI * n / NX;
    tw[n] = mk2((float)std::cos(a), (float)std::sin(a));
  }
  std::vector<std::complex<double>> in((size_t)NX * T);
  for (auto& z : in) z = {drand48() - 0.5, drand48() - 0.5};
  std::vector<float2[16]> reg(NT);
  // load: ownership x = j + J e
  for (int t = 0; t < NT; ++t) {
    const int c = t % T, j = t / T;
    for (int e = 0; e < 16; ++e) {
      const auto z = in[(size_t)(j + C::J * e) * T + c];
      reg[t][e] = mk2((float)z.real(), (float)z.imag());
    }
  }
  const float2 poison = mk2(NAN, NAN);
  std::vector<float2> ex((size_t)C::ROWS * T, poison);
  std::vector<int> hits(ex.size(), 0);
  auto note_writes = [&](int which) {
    // replay the write addressing to count collisions
    std::fill(hits.begin(), hits.end(), 0);
    for (int t = 0; t < NT; ++t) {
      const int c = t % T, j = t / T;
      float2 mark[16];
      for (int s = 0; s < 16; ++s) mark[s] = mk2(1.f, 0.f);
      std::vector<float2> tmp(ex.size(), mk2(0.f, 0.f));
      if (which == 1) ex_write1<NX>(tmp.data(), j, c, mark); else ex_write2<NX>(tmp.data(), j, c, mark);
      for (size_t q = 0; q < tmp.size(); ++q) if (tmp[q].x != 0.f) ++hits[q];
    }
    int total = 0;
    for (int h : hits) { if (h > 1) { std::printf("collision in write%d\n", which); std::exit(1); } total += h; }
    if (total != NX * T) { std::printf("write%d covers %d of %d\n", which, total, NX * T); std::exit(1); }
  };
  note_writes(1);
  note_writes(2);
  // stage 1
  for (int t = 0; t < NT; ++t) dft16<INV>(reg[t]);
  for (int t = 0; t < NT; ++t) ex_write1<NX>(ex.data(), t / T, t % T, reg[t]);
  for (int t = 0; t < NT; ++t) { ex_read<NX>(ex.data(), t / T, t % T, reg[t]); twiddle2<NX, INV>(reg[t], tw.data(), t / T); }
  // stage 2
  std::fill(ex.begin(), ex.end(), poison);
  for (int t = 0; t < NT; ++t) dft16<INV>(reg[t]);
  for (int t = 0; t < NT; ++t) ex_write2<NX>(ex.data(), t / T, t % T, reg[t]);
  for (int t = 0; t < NT; ++t) { ex_read<NX>(ex.data(), t / T, t % T, reg[t]); stage3<NX, INV>(reg[t], tw.data(), t / T); }
  // compare with the naive DFT, at the owned positions
  double err2 = 0, ref2 = 0;
  for (int c = 0; c < T; ++c)
    for (int k = 0; k < NX; ++k) {
      std::complex<double> acc = 0;
      for (int n = 0; n < NX; ++n) {
        const double a = (INV ? 2.0 : -2.0) * M_PI * (double)((long)n * k % NX) / NX;
        acc += in[(size_t)n * T + c] * std::complex<double>(std::cos(a), std::sin(a));
      }
      const int j = k % C::J, e = k / C::J;
      const float2 g = reg[j * T + c][e];
      err2 += std::norm(acc - std::complex<double>(g.x, g.y));
      ref2 += std::norm(acc);
    }
  return std::sqrt(err2 / ref2);
}

// two columns per thread, 16-byte exchange words (the form xpass16.cu uses)
template <int NX, bool INV>
static double run2() {
  using C = Cfg<NX>;
  const int NT = C::J * TP, NC = 2 * TP;
  std::vector<float2> tw(NX);
  for (int n = 0; n < NX; ++n) {
    const double a = -2.0 * M_PI * n / NX;
    tw[n] = mk2((float)std::cos(a), (float)std::sin(a));
  }
  std::vector<std::complex<double>> in((size_t)NX * NC);
  for (auto& z : in) z = {drand48() - 0.5, drand48() - 0.5};
  std::vector<float2[16]> ra(NT), rb(NT);
  for (int t = 0; t < NT; ++t) {
    const int cp = t % TP, j = t / TP;
    for (int e = 0; e < 16; ++e) {
      const auto za = in[(size_t)(j + C::J * e) * NC + 2 * cp], zb = in[(size_t)(j + C::J * e) * NC + 2 * cp + 1];
      ra[t][e] = mk2((float)za.real(), (float)za.imag());
      rb[t][e] = mk2((float)zb.real(), (float)zb.imag());
    }
  }
  float4 poison; poison.x = poison.y = poison.z = poison.w = NAN;
  std::vector<float4> ex((size_t)NX * TP, poison);
  for (int t = 0; t < NT; ++t) { dft16<INV>(ra[t]); dft16<INV>(rb[t]); }
  for (int t = 0; t < NT; ++t) ex4_write1<NX>(ex.data(), t / TP, t % TP, ra[t], rb[t]);
  for (int t = 0; t < NT; ++t) { ex4_read<NX>(ex.data(), t / TP, t % TP, ra[t], rb[t]); twiddle2x2<NX, INV>(ra[t], rb[t], tw.data(), t / TP); }
  std::fill(ex.begin(), ex.end(), poison);
  for (int t = 0; t < NT; ++t) { dft16<INV>(ra[t]); dft16<INV>(rb[t]); }
  for (int t = 0; t < NT; ++t) ex4_write2<NX>(ex.data(), t / TP, t % TP, ra[t], rb[t]);
  for (int t = 0; t < NT; ++t) { ex4_read<NX>(ex.data(), t / TP, t % TP, ra[t], rb[t]); stage3x2<NX, INV>(ra[t], rb[t], tw.data(), t / TP); }
  double err2 = 0, ref2 = 0;
  for (int c = 0; c < NC; ++c)
    for (int k = 0; k < NX; ++k) {
      std::complex<double> acc = 0;
      for (int n = 0; n < NX; ++n) {
        const double a = (INV ? 2.0 : -2.0) * M_PI * (double)((long)n * k % NX) / NX;
        acc += in[(size_t)n * NC + c] * std::complex<double>(std::cos(a), std::sin(a));
      }
      const int j = k % C::J, e = k / C::J, t = j * TP + c / 2;
      const float2 g = (c & 1) ? rb[t][e] : ra[t][e];
      err2 += std::norm(acc - std::complex<double>(g.x, g.y));
      ref2 += std::norm(acc);
    }
  return std::sqrt(err2 / ref2);
}

// NX = 2048 on a two-CTA cluster: threads 0..1023, rank = t / 512; two exchange buffers
template <bool INV>
static double run2k() {
  constexpr int NX = 2048;
  using C = Cfg<NX>;
  const int NT = 1024, ROWS = Cfg<1024>::ROWS;
  std::vector<float2> tw(NX);
  for (int n = 0; n < NX; ++n) {
    const double a = -2.0 * M_PI * n / NX;
    tw[n] = mk2((float)std::cos(a), (float)std::sin(a));
  }
  std::vector<std::complex<double>> in((size_t)NX * T);
  for (auto& z : in) z = {drand48() - 0.5, drand48() - 0.5};
  std::vector<float2[16]> reg(NT);
  for (int t = 0; t < NT; ++t) {
    const int rank = t / 512, jl = (t % 512) / T, c = t % T, j = 64 * rank + jl;
    for (int e = 0; e < 16; ++e) {
      const auto z = in[(size_t)(j + C::J * e) * T + c];
      reg[t][e] = mk2((float)z.real(), (float)z.imag());
    }
  }
  const float2 poison = mk2(NAN, NAN);
  std::vector<float2> b0((size_t)ROWS * T, poison), b1((size_t)ROWS * T, poison);
  float2* ex[2] = {b0.data(), b1.data()};
  auto J = [](int t) { return 64 * (t / 512) + (t % 512) / T; };
  // coverage / collisions of both write patterns
  for (int which = 1; which <= 2; ++which) {
    std::vector<int> h0(b0.size(), 0), h1(b1.size(), 0);
    for (int t = 0; t < NT; ++t)
      for (int s = 0; s < 16; ++s) {
        const int j = J(t), c = t % T;
        const int own = which == 1 ? c2k_owner1(j) : c2k_owner2(s);
        const int ad = which == 1 ? c2k_addr1(j, s, c) : c2k_addr2(j, s, c);
        if (ad < 0 || ad >= ROWS * T) { std::printf("2k write%d out of range\n", which); std::exit(1); }
        ++(own ? h1 : h0)[ad];
      }
    long tot = 0;
    for (auto* h : {&h0, &h1}) for (int v : *h) { if (v > 1) { std::printf("2k collision write%d\n", which); std::exit(1); } tot += v; }
    if (tot != (long)NX * T) { std::printf("2k write%d covers %ld\n", which, tot); std::exit(1); }
  }
  for (int t = 0; t < NT; ++t) dft16<INV>(reg[t]);
  for (int t = 0; t < NT; ++t) c2k_write1(ex[0], ex[1], J(t), t % T, reg[t]);
  for (int t = 0; t < NT; ++t) { ex_read<1024>(ex[t / 512], (t % 512) / T, t % T, reg[t]); twiddle2<NX, INV>(reg[t], tw.data(), J(t)); }
  std::fill(b0.begin(), b0.end(), poison); std::fill(b1.begin(), b1.end(), poison);
  for (int t = 0; t < NT; ++t) dft16<INV>(reg[t]);
  for (int t = 0; t < NT; ++t) c2k_write2(ex[0], ex[1], J(t), t % T, reg[t]);
  for (int t = 0; t < NT; ++t) { ex_read<1024>(ex[t / 512], (t % 512) / T, t % T, reg[t]); stage3<NX, INV>(reg[t], tw.data(), J(t)); }
  double err2 = 0, ref2 = 0;
  for (int c = 0; c < T; ++c)
    for (int k = 0; k < NX; ++k) {
      std::complex<double> acc = 0;
      for (int n = 0; n < NX; ++n) {
        const double a = (INV ? 2.0 : -2.0) * M_PI * (double)((long)n * k % NX) / NX;
        acc += in[(size_t)n * T + c] * std::complex<double>(std::cos(a), std::sin(a));
      }
      const int j = k % C::J, e = k / C::J;
      const int t = 512 * (j / 64) + (j % 64) * T + c;
      const float2 g = reg[t][e];
      err2 += std::norm(acc - std::complex<double>(g.x, g.y));
      ref2 += std::norm(acc);
    }
  return std::sqrt(err2 / ref2);
}

// half-warp bank model for the cluster write patterns (16 lanes = 2 consecutive j x 8 columns)
static int conflicts2k() {
  int bad = 0;
  for (int w = 0; w < 1024 / 16; ++w)
    for (int which = 1; which <= 2; ++which)
      for (int s = 0; s < 16; ++s) {
        int seen[16] = {0}, owner = -1;
        for (int l = 0; l < 16; ++l) {
          const int t = w * 16 + l, c = t % T, j = 64 * (t / 512) + (t % 512) / T;
          const int own = which == 1 ? c2k_owner1(j) : c2k_owner2(s);
          const int ad = which == 1 ? c2k_addr1(j, s, c) : c2k_addr2(j, s, c);
          if (owner < 0) owner = own;
          if (own != owner) ++bad;            // a half-warp must target one CTA
          if (seen[ad % 16]++) ++bad;
        }
      }
  return bad;
}

// bank-conflict model for 64-bit shared accesses: a half-warp (16 lanes) is one wavefront if
// its 16 words fall in 16 distinct 8-byte bank pairs (word index mod 16)
template <int NX>
static int conflicts() {
  using C = Cfg<NX>;
  int bad = 0;
  for (int w = 0; w < C::THREADS / 16; ++w)
    for (int which = 0; which < 3; ++which)
      for (int s = 0; s < 16; ++s) {
        int seen[16] = {0};
        for (int l = 0; l < 16; ++l) {
          const int t = w * 16 + l, c = t % T, j = t / T;
          int idx;
          if (which == 0) idx = ex_own<NX>(j, c) + s * C::ESTRIDE;
          else if (which == 1) idx = (17 * j) * T + c + s * T;
          else { const int k = j & 15; idx = (17 * (j - k) + k) * T + c + s * 17 * T; }
          if (seen[idx % 16]++) ++bad;
        }
      }
  return bad;
}

int main() {
  double worst = 0;
  double e;
  e = run<256, false>(); std::printf("NX=256 fwd rel err %.3g\n", e); worst = std::fmax(worst, e);
  e = run<256, true>(); std::printf("NX=256 inv rel err %.3g\n", e); worst = std::fmax(worst, e);
  e = run<512, false>(); std::printf("NX=512 fwd rel err %.3g\n", e); worst = std::fmax(worst, e);
  e = run<512, true>(); std::printf("NX=512 inv rel err %.3g\n", e); worst = std::fmax(worst, e);
  e = run<1024, false>(); std::printf("NX=1024 fwd rel err %.3g\n", e); worst = std::fmax(worst, e);
  e = run<1024, true>(); std::printf("NX=1024 inv rel err %.3g\n", e); worst = std::fmax(worst, e);
  e = run2<256, false>(); std::printf("NX=256 x2 fwd rel err %.3g\n", e); worst = std::fmax(worst, e);
  e = run2<512, true>(); std::printf("NX=512 x2 inv rel err %.3g\n", e); worst = std::fmax(worst, e);
  e = run2<1024, false>(); std::printf("NX=1024 x2 fwd rel err %.3g\n", e); worst = std::fmax(worst, e);
  e = run2<1024, true>(); std::printf("NX=1024 x2 inv rel err %.3g\n", e); worst = std::fmax(worst, e);
  e = run2k<false>(); std::printf("NX=2048 (cluster) fwd rel err %.3g\n", e); worst = std::fmax(worst, e);
  e = run2k<true>(); std::printf("NX=2048 (cluster) inv rel err %.3g\n", e); worst = std::fmax(worst, e);
  const int bc = conflicts<256>() + conflicts<512>() + conflicts<1024>() + conflicts2k();
  std::printf("bank conflicts (64-bit half-warp model): %d\n", bc);
  if (!(worst < 1e-6) || bc != 0) { std::printf("FAIL\n"); return 1; }
  std::printf("OK\n");
  return 0;
}
