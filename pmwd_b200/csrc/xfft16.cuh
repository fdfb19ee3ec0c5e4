// Register-resident FFT along x for the fused x-pass (xpass16.cu): NX = 16 * 16 * R3 with
// R3 in {1, 2, 4}, i.e. NX in {256, 512, 1024}, by one CTA; R3 = 8 (NX = 2048) by a cluster of two
// CTAs that exchange through distributed shared memory (last section of this file).
//
// One column per thread (first half of this file; the two-columns-per-thread form that
// xpass16.cu uses is at the end): a tile is T = 8 adjacent columns.  Thread (j, c), j in [0, NX/16),
// c in [0, T), OWNS the 16 points  x = j + (NX/16) e,  e = 0..15  of column c, in registers,
// before and after every transform (so loads, the k-space algebra and stores never touch shared
// memory).  A transform is a Stockham autosort with radices 16, 16, R3:
//     stage 1: 16-point DFT in registers               -> exchange through shared memory
//     stage 2: twiddle, 16-point DFT in registers      -> exchange through shared memory
//     stage 3: twiddle, 16/R3 R3-point DFTs in registers (in place: outputs are owned again)
// Two exchanges (4 barriers) per transform instead of the 10 shared-memory passes of a radix-4
// Stockham.  The exchange buffer rows are padded, phys(x) = x + (x >> 4), which makes every
// access pattern below conflict-free for 64-bit words (half-warp = two rows of opposite parity).
//
// The phase functions are __host__ __device__ so that tests/host/xfft16_emul.cc can run the
// index arithmetic thread by thread on the CPU against a naive DFT.
#pragma once
#include <vector_types.h>
#include <math.h>

#if defined(__CUDACC__)
#define PMWD_HD __host__ __device__ __forceinline__
#else
#define PMWD_HD inline
#endif

namespace pmwd {
namespace r16 {

constexpr int T = 8;

template <int NX>
struct Cfg {
  static_assert(NX == 256 || NX == 512 || NX == 1024 || NX == 2048, "NX = 256 * R3, R3 in {1,2,4,8}");
  static constexpr int J = NX / 16;          // threads per column; x stride between owned points
  static constexpr int R3 = NX / 256;        // radix of the last stage
  static constexpr int M3 = 16 / R3;         // last-stage butterflies per thread
  static constexpr int THREADS = J * T;
  static constexpr int ROWS = NX + NX / 16;  // padded rows of the exchange buffer
  static constexpr int ESTRIDE = (J + J / 16) * T;   // exchange-buffer distance between owned points
};

PMWD_HD float2 mk2(float x, float y) {
  float2 r;
  r.x = x;
  r.y = y;
  return r;
}

PMWD_HD float fma_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}

// a * (wx + i wy)
PMWD_HD float2 cmulw(float2 a, float wx, float wy) {
  return mk2(fma_(a.x, wx, -(a.y * wy)), fma_(a.x, wy, a.y * wx));
}

// 4-point DFT, natural order in and out; forward kernel exp(-2 pi i nk/4), inverse exp(+...)
template <bool INV>
PMWD_HD void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s02 = mk2(a0.x + a2.x, a0.y + a2.y), d02 = mk2(a0.x - a2.x, a0.y - a2.y);
  const float2 s13 = mk2(a1.x + a3.x, a1.y + a3.y), d13 = mk2(a1.x - a3.x, a1.y - a3.y);
  // forward: -i d13 = (d13.y, -d13.x); inverse: +i d13 = (-d13.y, d13.x)
  const float2 r = INV ? mk2(-d13.y, d13.x) : mk2(d13.y, -d13.x);
  a0 = mk2(s02.x + s13.x, s02.y + s13.y);
  a1 = mk2(d02.x + r.x, d02.y + r.y);
  a2 = mk2(s02.x - s13.x, s02.y - s13.y);
  a3 = mk2(d02.x - r.x, d02.y - r.y);
}

template <bool INV>
PMWD_HD void dft2(float2& a0, float2& a1) {
  const float2 s = mk2(a0.x + a1.x, a0.y + a1.y), d = mk2(a0.x - a1.x, a0.y - a1.y);
  a0 = s;
  a1 = d;
}

// 8-point DFT, natural order in and out: even / odd 4-point DFTs, then X[k], X[k+4] = E[k] +- w8^k O[k]
template <bool INV>
PMWD_HD void dft8(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4, float2& a5, float2& a6, float2& a7) {
  constexpr float C2 = 0.70710678118654752f;
  constexpr float sg = INV ? 1.f : -1.f;
  dft4<INV>(a0, a2, a4, a6);                 // E[0..3] in a0, a2, a4, a6
  dft4<INV>(a1, a3, a5, a7);                 // O[0..3] in a1, a3, a5, a7
  const float2 o0 = a1;
  const float2 o1 = cmulw(a3, C2, sg * C2);                                        // w8^1
  const float2 o2 = INV ? mk2(-a5.y, a5.x) : mk2(a5.y, -a5.x);                      // w8^2 = -+i
  const float2 o3 = cmulw(a7, -C2, sg * C2);                                       // w8^3
  const float2 e0 = a0, e1 = a2, e2 = a4, e3 = a6;
  a0 = mk2(e0.x + o0.x, e0.y + o0.y);
  a1 = mk2(e1.x + o1.x, e1.y + o1.y);
  a2 = mk2(e2.x + o2.x, e2.y + o2.y);
  a3 = mk2(e3.x + o3.x, e3.y + o3.y);
  a4 = mk2(e0.x - o0.x, e0.y - o0.y);
  a5 = mk2(e1.x - o1.x, e1.y - o1.y);
  a6 = mk2(e2.x - o2.x, e2.y - o2.y);
  a7 = mk2(e3.x - o3.x, e3.y - o3.y);
}

// 16-point DFT in registers, natural order in and out (n = 4 n1 + n2, k = k1 + 4 k2):
//   B[n2][k1] = sum_n1 v[4 n1 + n2] w4^(n1 k1);  B *= w16^(n2 k1);  V[k1 + 4 k2] = sum_n2 B[n2][k1] w4^(n2 k2)
template <bool INV>
PMWD_HD void dft16(float2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, C2 = 0.70710678118654752f;
  constexpr float sg = INV ? 1.f : -1.f;     // sign of the imaginary part of the twiddles
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // now v[4 k1 + n2] = B[n2][k1]; twiddle w16^(n2 k1) = cos(pi m / 8) + i sg sin(pi m / 8), m = n2 k1
  v[4 * 1 + 1] = cmulw(v[4 * 1 + 1], C1, sg * S1);     // m = 1
  v[4 * 1 + 2] = cmulw(v[4 * 1 + 2], C2, sg * C2);     // m = 2
  v[4 * 1 + 3] = cmulw(v[4 * 1 + 3], S1, sg * C1);     // m = 3
  v[4 * 2 + 1] = cmulw(v[4 * 2 + 1], C2, sg * C2);     // m = 2
  v[4 * 2 + 2] = INV ? mk2(-v[4 * 2 + 2].y, v[4 * 2 + 2].x) : mk2(v[4 * 2 + 2].y, -v[4 * 2 + 2].x);   // m = 4
  v[4 * 2 + 3] = cmulw(v[4 * 2 + 3], -C2, sg * C2);    // m = 6
  v[4 * 3 + 1] = cmulw(v[4 * 3 + 1], S1, sg * C1);     // m = 3
  v[4 * 3 + 2] = cmulw(v[4 * 3 + 2], -C2, sg * C2);    // m = 6
  v[4 * 3 + 3] = cmulw(v[4 * 3 + 3], -C1, -sg * S1);   // m = 9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  // v[4 k1 + k2] holds V[k1 + 4 k2]: transpose the 4x4 register file into natural order
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) {
      const float2 t = v[4 * a + b];
      v[4 * a + b] = v[4 * b + a];
      v[4 * b + a] = t;
    }
}

// ---- exchange buffer addressing (float2 index); rows padded: phys(x) = x + (x >> 4)
template <int NX>
PMWD_HD int ex_own(int j, int c) {            // owned point e lives at ex_own + e * ESTRIDE
  return (j + (j >> 4)) * T + c;
}
template <int NX>
PMWD_HD void ex_read(const float2* ex, int j, int c, float2 (&v)[16]) {
  const float2* p = ex + ex_own<NX>(j, c);
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = p[e * Cfg<NX>::ESTRIDE];
}
// stage 1 (radix 16, p = 1): output s of butterfly j goes to x = 16 j + s, phys = 17 j + s
template <int NX>
PMWD_HD void ex_write1(float2* ex, int j, int c, const float2 (&v)[16]) {
  float2* p = ex + (17 * j) * T + c;
#pragma unroll
  for (int s = 0; s < 16; ++s) p[s * T] = v[s];
}
// stage 2 (radix 16, p = 16): k = j & 15, output s goes to x = 16 (j - k) + k + 16 s,
// phys = 17 (j - k) + k + 17 s
template <int NX>
PMWD_HD void ex_write2(float2* ex, int j, int c, const float2 (&v)[16]) {
  const int k = j & 15;
  float2* p = ex + (17 * (j - k) + k) * T + c;
#pragma unroll
  for (int s = 0; s < 16; ++s) p[s * 17 * T] = v[s];
}

// tw[n] = exp(-2 pi i n / NX); stage 2 input r is multiplied by exp(-+2 pi i r k / 256)
template <int NX, bool INV>
PMWD_HD void twiddle2(float2 (&v)[16], const float2* tw, int j) {
  const int k = j & 15;
#pragma unroll
  for (int r = 1; r < 16; ++r) {
    const float2 w = tw[r * k * (NX / 256)];
    v[r] = cmulw(v[r], w.x, INV ? -w.y : w.y);
  }
}

// stage 3 (radix R3, p = 256): butterfly i = j + J m takes the owned points e = m + M3 r
// (x = i + 256 r), twiddle exp(-+2 pi i r i / NX), and is in place
template <int NX, bool INV>
PMWD_HD void stage3(float2 (&v)[16], const float2* tw, int j) {
  constexpr int R3 = Cfg<NX>::R3, M3 = Cfg<NX>::M3, J = Cfg<NX>::J;
  if constexpr (R3 > 1) {
#pragma unroll
  for (int m = 0; m < M3; ++m) {
    const int i = j + J * m;
#pragma unroll
    for (int r = 1; r < R3; ++r) {
      const float2 w = tw[r * i];
      v[m + M3 * r] = cmulw(v[m + M3 * r], w.x, INV ? -w.y : w.y);
    }
    if constexpr (R3 == 8)
      dft8<INV>(v[m], v[m + M3], v[m + 2 * M3], v[m + 3 * M3], v[m + 4 * M3], v[m + 5 * M3], v[m + 6 * M3],
                v[m + 7 * M3]);
    if constexpr (R3 == 4) dft4<INV>(v[m], v[m + M3], v[m + 2 * M3], v[m + 3 * M3]);
    if constexpr (R3 == 2) dft2<INV>(v[m], v[m + M3]);
  }
  }
}

// ---------------------------------------------------------------------------------------------
// Two adjacent columns per thread (a = column 2 cp, b = column 2 cp + 1): global accesses and the
// exchange use 16-byte words (a, b), rows of TP = 8 pairs = 128 bytes.  A quarter-warp (8 lanes,
// same j) reads or writes one whole 128-byte row, so every pattern is conflict-free without
// padding.  The twiddles are loaded once for both columns.
constexpr int TP = 8;

PMWD_HD float4 pack4(float2 a, float2 b) {
  float4 q;
  q.x = a.x; q.y = a.y; q.z = b.x; q.w = b.y;
  return q;
}

template <int NX>
PMWD_HD void ex4_read(const float4* ex, int j, int cp, float2 (&a)[16], float2 (&b)[16]) {
  const float4* p = ex + j * TP + cp;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float4 q = p[e * Cfg<NX>::J * TP];
    a[e] = mk2(q.x, q.y);
    b[e] = mk2(q.z, q.w);
  }
}
template <int NX>
PMWD_HD void ex4_write1(float4* ex, int j, int cp, const float2 (&a)[16], const float2 (&b)[16]) {
  float4* p = ex + (16 * j) * TP + cp;
#pragma unroll
  for (int s = 0; s < 16; ++s) p[s * TP] = pack4(a[s], b[s]);
}
template <int NX>
PMWD_HD void ex4_write2(float4* ex, int j, int cp, const float2 (&a)[16], const float2 (&b)[16]) {
  const int k = j & 15;
  float4* p = ex + (16 * (j - k) + k) * TP + cp;
#pragma unroll
  for (int s = 0; s < 16; ++s) p[s * 16 * TP] = pack4(a[s], b[s]);
}
template <int NX, bool INV>
PMWD_HD void twiddle2x2(float2 (&a)[16], float2 (&b)[16], const float2* tw, int j) {
  const int k = j & 15;
#pragma unroll
  for (int r = 1; r < 16; ++r) {
    const float2 w = tw[r * k * (NX / 256)];
    const float wy = INV ? -w.y : w.y;
    a[r] = cmulw(a[r], w.x, wy);
    b[r] = cmulw(b[r], w.x, wy);
  }
}
template <int NX, bool INV>
PMWD_HD void stage3x2(float2 (&a)[16], float2 (&b)[16], const float2* tw, int j) {
  constexpr int R3 = Cfg<NX>::R3, M3 = Cfg<NX>::M3, J = Cfg<NX>::J;
  if constexpr (R3 > 1) {
#pragma unroll
    for (int m = 0; m < M3; ++m) {
      const int i = j + J * m;
#pragma unroll
      for (int r = 1; r < R3; ++r) {
        const float2 w = tw[r * i];
        const float wy = INV ? -w.y : w.y;
        a[m + M3 * r] = cmulw(a[m + M3 * r], w.x, wy);
        b[m + M3 * r] = cmulw(b[m + M3 * r], w.x, wy);
      }
      if constexpr (R3 == 4) {
        dft4<INV>(a[m], a[m + M3], a[m + 2 * M3], a[m + 3 * M3]);
        dft4<INV>(b[m], b[m + M3], b[m + 2 * M3], b[m + 3 * M3]);
      }
      if constexpr (R3 == 2) {
        dft2<INV>(a[m], a[m + M3]);
        dft2<INV>(b[m], b[m + M3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// NX = 2048 on a cluster of two CTAs (512 threads each).  Global butterfly index j in [0, 128),
// CTA rank r owns j in [64 r, 64 r + 64) (local jl = j & 63) and the same 16 points per column as
// above, x = j + 128 e.  The one-column exchange buffer of 2048 rows is split by bit 6 of x:
// CTA (x >> 6) & 1 holds row x at local row xl = ((x >> 7) << 6) | (x & 63), padded as before
// (phys = xl + (xl >> 4), 1088 rows = the NX = 1024 buffer).  Every READ is then local and has
// exactly the NX = 1024 pattern (x = j + 128 e  ->  xl = jl + 64 e), so ex_read<1024>(ex, jl, ...)
// serves; only the WRITES of the two exchanges go to either CTA:
//   stage 1: x = 16 j + s          -> owner (j >> 2) & 1 (a whole warp, 4 consecutive j, writes to ONE CTA),
//            xl = 64 (j >> 3) + 16 (j & 3) + s,  phys = xl + 4 (j >> 3) + (j & 3)
//   stage 2: x = 256 (j >> 4) + k + 16 s, k = j & 15  -> owner (s >> 2) & 1,
//            xl = 64 (2 (j >> 4) + (s >> 3)) + 16 (s & 3) + k,  phys = xl + 4 (2 (j >> 4) + (s >> 3)) + (s & 3)
// Both keep the half-warp pattern "two rows of opposite parity" (conflict-free).
PMWD_HD int c2k_owner1(int j) { return (j >> 2) & 1; }
PMWD_HD int c2k_addr1(int j, int s, int c) {
  const int xl = 64 * (j >> 3) + 16 * (j & 3) + s;
  return (xl + 4 * (j >> 3) + (j & 3)) * T + c;
}
PMWD_HD int c2k_owner2(int s) { return (s >> 2) & 1; }
PMWD_HD int c2k_addr2(int j, int s, int c) {
  const int hi = 2 * (j >> 4) + (s >> 3);
  const int xl = 64 * hi + 16 * (s & 3) + (j & 15);
  return (xl + 4 * hi + (s & 3)) * T + c;
}
// ex0, ex1: the two CTAs' buffers as seen from the calling thread
PMWD_HD void c2k_write1(float2* ex0, float2* ex1, int j, int c, const float2 (&v)[16]) {
  float2* p = c2k_owner1(j) ? ex1 : ex0;
#pragma unroll
  for (int s = 0; s < 16; ++s) p[c2k_addr1(j, s, c)] = v[s];
}
PMWD_HD void c2k_write2(float2* ex0, float2* ex1, int j, int c, const float2 (&v)[16]) {
#pragma unroll
  for (int s = 0; s < 16; ++s) (c2k_owner2(s) ? ex1 : ex0)[c2k_addr2(j, s, c)] = v[s];
}

}  // namespace r16
}  // namespace pmwd
