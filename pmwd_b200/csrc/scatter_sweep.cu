// Tiled CIC deposit through shared-memory mesh tiles ("tile sweep"): pmwd/scatter.py:60-83 for
// the gravity fast path (3-D, int16 pmid, offset = whole planes, cell_size=None).
//
// The per-particle global-RED kernel (cic_fast.cu) is bound by the LSU's atomic issue rate
// (~1.3 cycles per lane: 31 cycles per RED warp instruction measured, profiles/r01_final_ncu_full.txt)
// and moves 2.3x its algorithmic DRAM bytes (memset + read-modify-write of the whole mesh).  Shared
// memory atomics are no way out on sm_100a (ATOMS: 2 cycles per lane, float add = CAS loop).  This
// kernel accumulates with PLAIN shared-memory read-modify-writes that are conflict free by
// construction, and writes every mesh cell that only one warp can touch with float4 STORES (no
// memset, no read).  A first version that binned 512-particle batches by z-band inside a CTA moved
// 1.08x the algorithmic bytes but needed 948 instructions per particle and spent 47 % of its time at
// CTA barriers (12-16 ms, profiles/r02_sweep_v1_ncu.txt); this one has no CTA barrier at all:
//
//  * the integrator keeps its particle storage sorted by (y-tile, z-tile, x-plane, y, z) (reorder.cu,
//    sweep key layout) and a table gives the particle range of every (tile, plane);
//  * a work item = one (y, z) tile of TY x BW cells x one segment of LX planes, swept by ONE WARP
//    with a private RING of four planes [(TY+1) rows][BW+1 cells] in shared memory: a particle sorted
//    at plane xs now sits at xs-1 .. xs+1 (it moved since the last re-sort) and touches planes
//    xs-1 .. xs+2; when plane xs is done, plane xs-1 is complete and is flushed;
//  * per 32 particles: stencil in registers, lanes that share a base cell merged by shuffles, then
//    eight phases (one per neighbour) of plain LDS/FADD/STS -- within a phase all lanes hit distinct
//    cells, and nobody else ever touches this warp's ring;
//  * flush: cells that only this warp can reach -> float4 stores; the tile's first row, its halo row,
//    its first float4 along z and its halo column (shared with the neighbouring tiles) and the three
//    planes at either end of the segment -> red.global.add(.v4).f32 onto cells that
//    sweep_zero_kernel cleared beforehand;
//  * particles outside their window (moved more than one plane, or out of the tile in y or z) are
//    "stragglers": appended to a list and deposited by a second, per-particle RED kernel that runs
//    after this one (so that it can never race with the plain stores).
//
// Correctness never depends on how fresh the sort is; only the straggler fraction does.
#include <math.h>
#include <stdlib.h>

#include "cic.cuh"

namespace pmwd {

int slab_xoff(const pmwd_cic_desc* d);
bool cic_is_fast(const pmwd_cic_desc* d);

constexpr int SW_MAX_WARPS = 24;
constexpr int SW_STRAG_BLOCK = 256;               // straggler-list slots a warp reserves per global atomic
constexpr uint32_t SW_STRAG_EMPTY = 0xffffffffu;  // unused slot of a reserved block
constexpr int64_t SW_STRAG_SLACK = (int64_t)SW_STRAG_BLOCK * SW_MAX_WARPS * 256;   // one open block per warp, <= 256 CTAs

struct SweepGeom {
  int64_t n;
  int nx, ny, nz;     // periodic wrap shape (the GLOBAL mesh on a slab)
  int nx_ext, xoff;   // planes held by the mesh array, global index of plane 0
  int periodic;       // the array spans the whole periodic x axis
  float cell;
  float inv_cell;     // 1 / cell when that is exact (cell a power of two), else 0: multiply instead of divide
  int ty, npencil;    // tile rows, ny / ty
  int bw, nband;      // tile cells along z (multiple of 4), nz / bw
  int lx, nseg;       // planes per x segment, ceil(nx_ext / lx)
  int rs, ps;         // ring row stride (bw + 4: float4-aligned rows, halo column at bw) and plane size (ty + 1) * rs
  int nwarps;         // warps per CTA (each with its own ring)
};

__device__ __forceinline__ void red_add4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// same helper as cic_fast.cu (merge the contributions of lanes that share a key)
template <int NV>
__device__ __forceinline__ bool sweep_reduce_peers(unsigned peers, float (&v)[NV]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int rel_pos = __popc(peers << (31 - lane) << 1);
  const bool first = rel_pos == 0;
  peers &= (0xfffffffeu << lane);
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

struct PtclRegs {
  short pm[3];
  float dp[3];
  float v;
};

// periodic wrap of pmid + floor(disp / cell): almost always within one box length
__device__ __forceinline__ int wrap_fast(int i, int n) {
  if (i >= n) { i -= n; if (i >= n) i %= n; }
  else if (i < 0) { i += n; if (i < 0) { i %= n; if (i < 0) i += n; } }
  return i;
}

// disp / cell in float32 (pm_util.py:133): for a power-of-two cell the product with the exact reciprocal is the
// same correctly rounded number, without the division's ~10 instructions (and its slow path for zero numerators)
__device__ __forceinline__ float cell_units(float d, const SweepGeom& G) {
  return G.inv_cell != 0.f ? __fmul_rn(d, G.inv_cell) : __fdiv_rn(d, G.cell);
}

// TY, BW > 0: tile shape known at compile time (the flush loop's index arithmetic folds away); 0, 0: G.ty, G.bw.
// DET: bitwise reproducible deposit (PMWD_SCATTER_DETERMINISTIC inside the integrator).  What makes the plain
// kernel order dependent are the float REDs onto cells that two warps reach; here nothing is shared:
//   * one work item = one (y, z) tile over ALL planes (lx = nx, periodic mesh), so along x only the wrap-around
//     planes nx-1, 0, 1 are visited twice -- by the same lanes of the same warp, the second time as a plain
//     load-add-store;
//   * the tile's own cells (incl. its first row and first column) are written with plain stores; what it
//     deposits into its halo row / halo column / halo corner (cells of the NEXT tile along y / z) goes to
//     per-tile halo arrays (`halo`), which sweep_det_fix_kernel adds to the mesh afterwards in a fixed order.
// The arithmetic inside a tile is sequential in the storage order (chunk by chunk, lanes merged in a fixed
// tree, neighbours in a fixed sequence), and which warp sweeps which tile does not matter.  Needs freshly
// sorted storage: a straggler would be deposited by REDs; it is counted in counters[3] so that the host can refuse.
template <int TY, int BW, bool DET>
__global__ void __launch_bounds__(SW_MAX_WARPS * 32, 1)
scatter_sweep_kernel(SweepGeom G, const short* __restrict__ pmid, const float* __restrict__ disp,
                     const float* __restrict__ val, int vstride, float vscalar, float* __restrict__ mesh,
                     const uint2* __restrict__ table, unsigned* __restrict__ counters,
                     uint32_t* __restrict__ strag, int record_strag, float* __restrict__ halo) {
  extern __shared__ __align__(16) float smf[];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ty = TY ? TY : G.ty, bw = BW ? BW : G.bw;
  const int rs = bw + 4, ps = (ty + 1) * rs;            // ring row stride (halo column at bw), plane size
  const int ngr = bw >> 2;                              // float4 groups per tile row
  float* ring = smf + (size_t)warp * 4 * ps;            // [4][ty+1][bw+4], private to this warp
  uint32_t* sbuf = reinterpret_cast<uint32_t*>(smf + (size_t)G.nwarps * 4 * ps) + warp * 64;   // straggler staging
  int scount = 0;                                       // staged stragglers (< 32 between chunks)
  unsigned sres_pos = 0, stotal = 0;                    // next reserved slot of the global list, stragglers so far
  int sres_left = 0;                                    // reserved slots left
  const int nitems = G.npencil * G.nband * G.nseg;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t plane_elems = (int64_t)G.ny * G.nz;

  for (;;) {
    int item = 0;
    if (lane == 0) item = (int)atomicAdd(&counters[0], 1u);
    item = __shfl_sync(full, item, 0);
    if (item >= nitems) break;
    const int band = item % G.nband;
    const int t1 = item / G.nband;
    const int pencil = t1 % G.npencil, sgi = t1 / G.npencil;
    const int xa = sgi * G.lx, xb = min(xa + G.lx, G.nx_ext);
    const int y0 = pencil * ty, z0 = band * bw;
    const int wrap_row = (y0 + ty == G.ny) ? ty : -1;   // the halo row of the last y-tile is row 0 of the mesh
    const int zh = (z0 + bw == G.nz) ? 0 : z0 + bw;     // halo column: the first cell of the next tile along z
    const uint2* tab = table + ((int64_t)pencil * G.nband + band) * G.nx_ext;
    for (int i = lane; i < ps; i += 32) reinterpret_cast<float4*>(ring)[i] = zero4;
    __syncwarp();

    // ---- flush one ring plane of this warp to the mesh, clearing it on the way
    auto flush = [&](int pl) {
      const int slot = (pl - (xa - 1)) & 3;
      float* src = ring + slot * ps;
      bool valid = true;
      int gpl = pl;
      if (G.periodic) {
        if (gpl < 0) gpl += G.nx;
        else if (gpl >= G.nx) gpl -= G.nx;
      } else {
        valid = pl >= 0 && pl < G.nx_ext;
      }
      float* mpl = mesh + (int64_t)gpl * plane_elems;
      const int ntot = (ty + 1) * ngr;
      if (DET) {
        // second visit of a wrap-around plane (same warp, same lanes): accumulate instead of overwrite
        const bool again = pl >= G.nx - 1;
        const int64_t tp = ((int64_t)pencil * G.nband + band) * G.nx + gpl;       // (tile, plane)
        float* hy = halo + tp * bw;                                               // halo row   [bw]
        float* hz = halo + (int64_t)nitems * G.nx * bw + tp * (ty + 1);           // halo column [ty] + corner
#pragma unroll
        for (int i0 = 0; i0 < ntot; i0 += 32) {
          const int i = i0 + lane;
          if (i < ntot) {
            const int row = i / ngr, gq = i - row * ngr;
            float4* sp = reinterpret_cast<float4*>(src + row * rs + 4 * gq);
            float4 v = *sp;
            *sp = zero4;
            float4* dst = reinterpret_cast<float4*>(row == ty ? hy + 4 * gq : mpl + (unsigned)((y0 + row) * G.nz + z0 + 4 * gq));
            if (again) {
              const float4 o = *dst;
              v.x = __fadd_rn(o.x, v.x); v.y = __fadd_rn(o.y, v.y); v.z = __fadd_rn(o.z, v.z); v.w = __fadd_rn(o.w, v.w);
            }
            *dst = v;
          }
        }
        for (int row = lane; row <= ty; row += 32) {
          float v = src[row * rs + bw];
          src[row * rs + bw] = 0.f;
          if (again) v = __fadd_rn(hz[row], v);
          hz[row] = v;
        }
      } else {
      const bool all_red = (pl <= xa + 1) || (pl >= xb - 1);
#pragma unroll
      for (int i0 = 0; i0 < ntot; i0 += 32) {
        const int i = i0 + lane;
        if (i < ntot) {
          const int row = i / ngr, gq = i - row * ngr;
          float4* sp = reinterpret_cast<float4*>(src + row * rs + 4 * gq);
          const float4 v = *sp;
          *sp = zero4;
          if (valid) {
            const int gy = row == wrap_row ? 0 : y0 + row;
            float* dst = mpl + (unsigned)(gy * G.nz + z0 + 4 * gq);
            if (all_red || row == 0 || row == ty || gq == 0) {
              if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add4(dst, v);
            } else {
              __stcs(reinterpret_cast<float4*>(dst), v);
            }
          }
        }
      }
      for (int row = lane; row <= ty; row += 32) {
        const float v = src[row * rs + bw];
        src[row * rs + bw] = 0.f;
        if (valid && v != 0.f) {
          const int gy = row == wrap_row ? 0 : y0 + row;
          atomicAdd(mpl + (unsigned)(gy * G.nz + zh), v);
        }
      }
      }
      __syncwarp();
    };

    // ---- chunk iterator over the item's (plane, particle range) list; uniform across the warp.
    // The table entry of the next plane is always in flight while the current plane is worked on.
    int xs = xa;
    unsigned b0 = 0, bend = 0;
    uint2 ahead;
    { const uint2 s = __ldg(tab + xa); b0 = s.x; bend = s.y; }
    ahead = __ldg(tab + min(xa + 1, xb - 1));
    auto seek = [&]() {      // make (xs, b0, bend) point at a non-empty chunk or xs == xb
      while (xs < xb) {
        if (b0 < bend) return;
        ++xs;
        if (xs < xb) {
          b0 = ahead.x; bend = ahead.y;
          ahead = __ldg(tab + min(xs + 1, xb - 1));
        }
      }
    };
    seek();
    auto load = [&](int lxs, unsigned lb0, unsigned lbend, PtclRegs& R) -> bool {
      const unsigned p = lb0 + lane;
      const bool act = lxs < xb && p < lbend;
      if (act) {
        const short* pm = pmid + 3 * (int64_t)p;
        const float* dp = disp + 3 * (int64_t)p;
        R.pm[0] = pm[0]; R.pm[1] = pm[1]; R.pm[2] = pm[2];
        R.dp[0] = dp[0]; R.dp[1] = dp[1]; R.dp[2] = dp[2];
        R.v = val ? val[(int64_t)vstride * p] : vscalar;
      }
      return act;
    };
    // two chunks in flight behind the one being worked on (the loads' latency is several chunk times)
    struct Chunk { PtclRegs R; int xs; unsigned b0; bool act; };
    // Take the chunk the iterator points at, advance the iterator, THEN issue the chunk's loads: seek()
    // consumes the prefetched table entry, and a scoreboard wait right behind freshly issued particle
    // loads would wait for those as well (11 % of all stall samples before this order)
    auto fetch = [&](Chunk& c) {
      c.xs = xs; c.b0 = b0;
      const unsigned cbend = bend;
      if (xs < xb) { b0 += 32; seek(); }
      c.act = load(c.xs, c.b0, cbend, c.R);
    };
    Chunk c0, c1, c2;
    fetch(c0);
    fetch(c1);
    int next_flush = xa - 1;

    while (c0.xs < xb) {
      fetch(c2);
      const int cxs = c0.xs;
      const unsigned cb0 = c0.b0;
      const bool cur_act = c0.act;
      const PtclRegs& cur = c0.R;

      while (next_flush <= cxs - 2) { flush(next_flush); ++next_flush; }   // planes out of reach

      // ---- stencil of this lane's particle
      const unsigned p = cb0 + lane;
      int g[3];
      float d0[3];
      const int nn[3] = {G.nx, G.ny, G.nz};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        // idle lanes: a numerator the division's fast path accepts (a zero sends the whole warp down its slow path)
        const float t = cell_units(cur_act ? cur.dp[a] : G.cell, G);
        const float fl = floorf(t);
        d0[a] = __fsub_rn(t, fl);
        g[a] = wrap_fast((cur_act ? (int)cur.pm[a] : 0) + (int)fl, nn[a]);
      }
      int lp = g[0] - G.xoff;
      if (lp < 0) lp += G.nx;
      int dxs = lp - cxs;
      if (G.periodic) {
        if (dxs > (G.nx >> 1)) dxs -= G.nx;
        else if (dxs < -(G.nx >> 1)) dxs += G.nx;
      }
      const int dy = g[1] - y0, dz = g[2] - z0;
      const bool inwin = cur_act && dxs >= -1 && dxs <= 1 && dy >= 0 && dy < ty && dz >= 0 && dz < bw;
      // lanes with the same base cell (issued early: its latency hides behind the weights)
      const unsigned key = inwin ? (unsigned)(((dxs + 1) * ty + dy) * bw + dz) : (0x80000000u | (unsigned)lane);
      unsigned peers = __match_any_sync(full, key);
      // ---- contributions: independent work placed behind the match (~150 cycles of latency); the empty
      // volatile asm keeps the compiler from hoisting it above
      asm volatile("" : "+f"(d0[0]), "+f"(d0[1]), "+f"(d0[2]));
      float wx[2], wy[2], wz[2];
      wx[0] = __fsub_rn(1.f, fabsf(d0[0])); wx[1] = __fsub_rn(1.f, fabsf(__fsub_rn(d0[0], 1.f)));
      wy[0] = __fsub_rn(1.f, fabsf(d0[1])); wy[1] = __fsub_rn(1.f, fabsf(__fsub_rn(d0[1], 1.f)));
      wz[0] = __fsub_rn(1.f, fabsf(d0[2])); wz[1] = __fsub_rn(1.f, fabsf(__fsub_rn(d0[2], 1.f)));
      const float v = inwin ? cur.v : 0.f;
      float c8[8];
#pragma unroll
      for (int bx = 0; bx < 2; ++bx)
#pragma unroll
        for (int by = 0; by < 2; ++by) {
          const float wxy = __fmul_rn(wx[bx], wy[by]);
#pragma unroll
          for (int bz = 0; bz < 2; ++bz)
            c8[(bx * 2 + by) * 2 + bz] = __fmul_rn(v, __fmul_rn(wxy, wz[bz]));
        }
      if (record_strag) {
        // stragglers are staged in a per-warp buffer and appended to the global list 32 at a time, into
        // slots reserved SW_STRAG_BLOCK at a time (one global atomic per chunk on a single counter
        // stalled half of all issue slots; one per 32 stragglers still cost 5 % of the kernel)
        const unsigned sm = __ballot_sync(full, cur_act && !inwin);
        if (sm) {
          if (cur_act && !inwin) sbuf[scount + __popc(sm & ((1u << lane) - 1u))] = p;
          scount += __popc(sm);
          __syncwarp();
          if (scount >= 32) {
            if (sres_left == 0) {
              unsigned base = 0;
              if (lane == 0) base = atomicAdd(&counters[1], (unsigned)SW_STRAG_BLOCK);
              sres_pos = __shfl_sync(full, base, 0);
              sres_left = SW_STRAG_BLOCK;
            }
            strag[sres_pos + lane] = sbuf[lane];
            sres_pos += 32; sres_left -= 32; stotal += 32;
            __syncwarp();
            if (lane < scount - 32) sbuf[lane] = sbuf[32 + lane];
            scount -= 32;
            __syncwarp();
          }
        }
      }
      // ---- merged over lanes with the same base cell
      const bool first = sweep_reduce_peers<8>(peers, c8);
      if (__any_sync(full, inwin)) {
        const bool push = inwin && first;
        const int slot = (cxs + dxs - (xa - 1)) & 3;
        const int off = dy * rs + dz;
        float* p0 = ring + slot * ps + off;
        float* p1 = ring + ((slot + 1) & 3) * ps + off;
        // Lane A's upper x-plane can be lane B's lower one (different base planes, hence different keys, same
        // (y, z) cell), so in general the eight neighbours go one per phase.  If all base planes of the chunk
        // have the same parity that cannot happen, and the two x-planes of a stencil -- different ring planes
        // -- go together: four phases of two independent LDS/FADD/STS chains.
        const bool mixed = __any_sync(full, inwin && (dxs & 1)) && __any_sync(full, inwin && !(dxs & 1));
        if (!mixed) {
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const int o = ((n & 2) ? rs : 0) + (n & 1);
            if (push) {
              const float a0 = p0[o], a1 = p1[o];
              p0[o] = a0 + c8[n];
              p1[o] = a1 + c8[4 + n];
            }
            __syncwarp();
          }
        } else {
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            float* cellp = ((n & 4) ? p1 : p0) + ((n & 2) ? rs : 0) + (n & 1);
            if (push) *cellp = *cellp + c8[n];
            __syncwarp();
          }
        }
      }
      c0 = c1;
      c1 = c2;
    }
    // ---- the rest of the ring: planes up to xb + 1
    while (next_flush <= xb + 1) { flush(next_flush); ++next_flush; }
  }
  if (record_strag) {
    // the staged rest, then the unused slots of this warp's last block are marked empty
    if (scount > 0 && sres_left == 0) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(&counters[1], (unsigned)SW_STRAG_BLOCK);
      sres_pos = __shfl_sync(full, base, 0);
      sres_left = SW_STRAG_BLOCK;
    }
    for (int i = lane; i < sres_left; i += 32) strag[sres_pos + i] = i < scount ? sbuf[i] : SW_STRAG_EMPTY;
    stotal += scount;
    if (lane == 0 && stotal) {
      atomicAdd(&counters[2], stotal);
      if (DET) atomicAdd(&counters[3], stotal);      // sticky: a deterministic deposit that was not (memset never clears it)
    }
  }
}

// Deterministic deposit, second half: add the per-tile halo arrays to the mesh.  One thread per (plane, cell of
// the first row of a y-tile) and one per (plane, cell of the first column of a z-tile); a cell that is both gets
// its terms in the fixed order own + halo row (tile above) + halo column (tile before) + corner (diagonal tile).
template <bool COLS>
__global__ void __launch_bounds__(256)
sweep_det_fix_kernel(SweepGeom G, float* __restrict__ mesh, const float* __restrict__ halo) {
  const int64_t ntile = (int64_t)G.npencil * G.nband;
  const float* hyb = halo;
  const float* hzb = halo + ntile * G.nx * G.bw;
  if (!COLS) {
    const int64_t total = (int64_t)G.nx * G.npencil * G.nz;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int z = (int)(i % G.nz);
      const int64_t r = i / G.nz;
      const int pencil = (int)(r % G.npencil), pl = (int)(r / G.npencil);
      const int band = z / G.bw, k = z - band * G.bw;
      const int up = pencil == 0 ? G.npencil - 1 : pencil - 1;                  // the tile whose halo row this is
      const int64_t tp = ((int64_t)up * G.nband + band) * G.nx + pl;
      float* c = mesh + ((int64_t)pl * G.ny + (int64_t)pencil * G.ty) * G.nz + z;
      *c = __fadd_rn(*c, hyb[tp * G.bw + k]);
    }
  } else {
    const int64_t total = (int64_t)G.nx * G.ny * G.nband;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int band = (int)(i % G.nband);
      const int64_t r = i / G.nband;
      const int y = (int)(r % G.ny), pl = (int)(r / G.ny);
      const int pencil = y / G.ty, row = y - pencil * G.ty;
      const int prev = band == 0 ? G.nband - 1 : band - 1;                      // the tile whose halo column this is
      const int64_t tp = ((int64_t)pencil * G.nband + prev) * G.nx + pl;
      float* c = mesh + ((int64_t)pl * G.ny + y) * G.nz + (int64_t)band * G.bw;
      float v = __fadd_rn(*c, hzb[tp * (G.ty + 1) + row]);
      if (row == 0) {
        const int up = pencil == 0 ? G.npencil - 1 : pencil - 1;
        const int64_t td = ((int64_t)up * G.nband + prev) * G.nx + pl;
        v = __fadd_rn(v, hzb[td * (G.ty + 1) + G.ty]);                          // the diagonal tile's corner
      }
      *c = v;
    }
  }
}

// Cells that receive REDs must start at zero: the first row of every y-tile (it is also the halo row
// of the tile below) and whole planes b-1, b, b+1 at every segment boundary b -> full rows; the first
// float4 of every z-tile (it holds the halo cell of the tile before) -> 16 bytes per tile.  One warp per row.
__global__ void __launch_bounds__(256)
sweep_zero_kernel(SweepGeom G, float* __restrict__ mesh) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows = (int64_t)G.nx_ext * G.ny;
  const int nz4 = G.nz >> 2, bw4 = G.bw >> 2;
  for (int64_t r = warp; r < rows; r += nwarp) {
    const int pl = (int)(r / G.ny), y = (int)(r - (int64_t)pl * G.ny);
    const int m = pl % G.lx;
    const bool whole = (y % G.ty == 0) || m == 0 || m == 1 || m == G.lx - 1 || pl == G.nx_ext - 1;
    float4* row = reinterpret_cast<float4*>(mesh + r * G.nz);
    if (whole) {
      for (int i = lane; i < nz4; i += 32) row[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      // whole 32-byte sectors (two float4) where the tile is wide enough: a half-sector write makes L2 fetch
      // the other half first; the second float4 is overwritten by the sweep's plain store afterwards
      const int b2 = 2 * G.nband;
      if (bw4 >= 2 && (bw4 & 1) == 0) {
        for (int i = lane; i < b2; i += 32) row[(i >> 1) * bw4 + (i & 1)] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        for (int b = lane; b < G.nband; b += 32) row[b * bw4] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

// Stragglers: per-particle global REDs (drop rule of enmesh for slabs), after the sweep.
__global__ void __launch_bounds__(256)
sweep_straggler_kernel(SweepGeom G, const short* __restrict__ pmid, const float* __restrict__ disp,
                       const float* __restrict__ val, int vstride, float vscalar, float* __restrict__ mesh,
                       const unsigned* __restrict__ counters, const uint32_t* __restrict__ strag) {
  const unsigned n = counters[1];
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (strag[i] == SW_STRAG_EMPTY) continue;
    const int64_t p = strag[i];
    int ix[2], iy[2], iz[2];
    float wx[2], wy[2], wz[2];
    const int nn[3] = {G.nx, G.ny, G.nz};
    int* idx[3] = {ix, iy, iz};
    float* w[3] = {wx, wy, wz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float t = __fdiv_rn(disp[3 * p + a], G.cell);
      const int i0 = (int)floorf(t);
      const float d = __fsub_rn(t, (float)i0);
      w[a][0] = __fsub_rn(1.f, fabsf(d));
      w[a][1] = __fsub_rn(1.f, fabsf(__fsub_rn(d, 1.f)));
      const int i = wrap_index((int)pmid[3 * p + a] + i0, nn[a]);
      idx[a][0] = i;
      idx[a][1] = (i + 1 == nn[a]) ? 0 : i + 1;
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      int l = ix[b] - G.xoff;
      if (l < 0) l += G.nx;
      ix[b] = l < G.nx_ext ? l : -1;
    }
    const float v = val ? val[(int64_t)vstride * p] : vscalar;
#pragma unroll
    for (int bx = 0; bx < 2; ++bx) {
      if (ix[bx] < 0) continue;
#pragma unroll
      for (int by = 0; by < 2; ++by)
#pragma unroll
        for (int bz = 0; bz < 2; ++bz)
          atomicAdd(mesh + ((int64_t)ix[bx] * G.ny + iy[by]) * G.nz + iz[bz],
                    __fmul_rn(v, __fmul_rn(__fmul_rn(wx[bx], wy[by]), wz[bz])));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (tile, plane) segment table from the keys of the last cell sort (reorder.cu, sweep key layout):
// id(i) = key / (ty * bw) = (y-tile * nband + z-tile) * nx_ext + plane of particle slot i.
__global__ void __launch_bounds__(256)
sweep_table_kernel(int64_t n, const uint32_t* __restrict__ keys, uint32_t key_div, uint2* __restrict__ table) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t id = keys[i] / key_div;
    const uint32_t idp = i > 0 ? keys[i - 1] / key_div : 0xffffffffu;
    if (i == 0 || idp != id) {
      table[id].x = (uint32_t)i;
      if (i > 0) table[idp].y = (uint32_t)i;
    }
    if (i == n - 1) table[id].y = (uint32_t)n;
  }
}

// every particle slot must lie inside the range of its own id (together with the total length this
// makes the table an exact partition): counts violations
__global__ void __launch_bounds__(256)
sweep_table_check_kernel(int64_t n, const uint32_t* __restrict__ keys, uint32_t key_div,
                         const uint2* __restrict__ table, unsigned* __restrict__ bad) {
  unsigned local = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint2 s = table[keys[i] / key_div];
    if (!((uint32_t)i >= s.x && (uint32_t)i < s.y)) ++local;
  }
  if (local) atomicAdd(bad, local);
}

// total length of the ranges (must equal n)
__global__ void __launch_bounds__(256)
sweep_table_cover_kernel(int64_t nid, const uint2* __restrict__ table, unsigned long long* __restrict__ total) {
  unsigned long long local = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nid; i += (int64_t)gridDim.x * blockDim.x) {
    const uint2 s = table[i];
    if (s.y > s.x) local += s.y - s.x;
    else if (s.y < s.x) local += 1ull << 40;       // inverted range: poison
  }
  if (local) atomicAdd(total, local);
}

// ---------------------------------------------------------------------------------------------
// per warp: the ring + 64 words of straggler staging
static size_t sweep_ring_bytes(int ty, int bw) { return (size_t)4 * (ty + 1) * (bw + 4) * sizeof(float) + 256; }

// Fills G for (d, ty, bw, lx); returns false if the geometry is not supported by the sweep kernels.
static bool sweep_geom(const pmwd_cic_desc* d, int ty, int bw, int lx, SweepGeom* G) {
  if (!cic_is_fast(d)) return false;
  G->n = d->ptcl_num;
  G->nx = d->wrap_shape[0]; G->ny = d->wrap_shape[1]; G->nz = d->wrap_shape[2];
  G->nx_ext = d->mesh_shape[0];
  G->xoff = slab_xoff(d);
  G->periodic = (G->nx_ext == G->nx && G->xoff == 0) ? 1 : 0;
  if (!G->periodic && G->nx_ext > G->nx) return false;      // halos wider than the box: RED kernel
  G->cell = (float)d->cell_size;
  {
    int e = 0;
    const float m = frexpf(G->cell, &e);
    G->inv_cell = (m == 0.5f && e > -100 && e < 100) ? 1.f / G->cell : 0.f;
    const char* env = getenv("PMWD_SWEEP_DIV");
    if (env && atoi(env) > 0) G->inv_cell = 0.f;
  }
  if (ty < 2 || ty > 64 || G->ny % ty != 0) return false;
  if (bw < 4 || bw > 256 || (bw & 3) != 0 || G->nz % bw != 0) return false;
  if (lx < 1 || G->nx_ext < 4) return false;
  G->ty = ty; G->npencil = G->ny / ty;
  G->bw = bw; G->nband = G->nz / bw;
  G->lx = lx < G->nx_ext ? lx : G->nx_ext;
  G->nseg = (G->nx_ext + G->lx - 1) / G->lx;
  G->rs = bw + 4;
  G->ps = (ty + 1) * G->rs;
  if (d->ptcl_num >= ((int64_t)1 << 32)) return false;
  if ((int64_t)G->nx_ext * G->ny * G->nz > ((int64_t)1 << 32)) return false;     // 32-bit sort keys
  if ((int64_t)G->npencil * G->nband * G->nseg >= ((int64_t)1 << 31)) return false;
  const size_t avail = 226 * 1024;
  int nw = (int)(avail / sweep_ring_bytes(ty, bw));
  if (nw > SW_MAX_WARPS) nw = SW_MAX_WARPS;
  if (nw < 4) return false;
  G->nwarps = nw;
  return true;
}

}  // namespace pmwd

using namespace pmwd;

// Tile shape for this mesh: ty = the largest divisor of ny up to 8, bw = a multiple of 4 dividing nz
// (64 preferred).  Returns 1 and fills *ty, *bw, or 0 if the sweep kernels do not support the mesh.
extern "C" int pmwd_sweep_pick(const pmwd_cic_desc* d, int* ty_out, int* bw_out) {
  if (!d || d->dim != 3) return 0;
  const int ny = d->wrap_shape[1], nz = d->wrap_shape[2];
  int ty = 0, bw = 0;
  // measured at 512^3 / 1024^3 over the whole 63-step run (profiles/r02_sweep_tiles.txt): 8 x 64 tiles
  // 33.85 ms/step, 16 x 32: 34.42; 8 x 32 and 16 x 64 are slower (more RED rows / fewer warps per SM)
  for (int t = 8; t >= 2; --t) if (ny % t == 0) { ty = t; break; }
  for (int b = 64; b >= 4; b -= 4) if (nz % b == 0) { bw = b; break; }
  if (bw < 16) for (int b = 68; b <= 128; b += 4) if (nz % b == 0) { bw = b; break; }
  { const char* e = getenv("PMWD_SWEEP_TY"); if (e && atoi(e) > 0 && ny % atoi(e) == 0) ty = atoi(e); }
  { const char* e = getenv("PMWD_SWEEP_BW"); if (e && atoi(e) > 0 && nz % atoi(e) == 0 && atoi(e) % 4 == 0) bw = atoi(e); }
  if (!ty || !bw) return 0;
  SweepGeom G;
  if (!sweep_geom(d, ty, bw, 64, &G)) return 0;
  if (ty_out) *ty_out = ty;
  if (bw_out) *bw_out = bw;
  return 1;
}

extern "C" size_t pmwd_sweep_table_bytes(const pmwd_cic_desc* d, int ty, int bw) {
  if (!d || ty <= 0 || bw <= 0 || d->wrap_shape[1] % ty || d->wrap_shape[2] % bw) return 0;
  return (size_t)(d->wrap_shape[1] / ty) * (d->wrap_shape[2] / bw) * d->mesh_shape[0] * sizeof(uint2);
}

// counters (64 bytes: [0] work items, [1] length of the straggler list, [2] stragglers) + straggler list
// (4 bytes per particle + one open block of reserved slots per warp)
static size_t sweep_scratch_bytes(int64_t n) { return 64 + (size_t)(n + SW_STRAG_SLACK) * sizeof(uint32_t); }

extern "C" size_t pmwd_sweep_scratch_bytes(const pmwd_cic_desc* d) {
  if (!d) return 0;
  return sweep_scratch_bytes(d->ptcl_num);
}

// Build the (tile, plane) table from the sorted keys of pmwd_cell_sort_perm(..., ty, bw).
// `status` (device, 16 bytes, zeroed here): [0] = particles outside their segment (unsigned),
// [8] = total length of all ranges (unsigned long long); the table is valid iff status[0] == 0 and
// total == ptcl_num (true by construction for sorted keys; tests read it back).  Enqueue-only.
extern "C" int pmwd_sweep_table(void* stream, const pmwd_cic_desc* d, int ty, int bw, const uint32_t* keys,
                                uint32_t* table, void* status) {
  PMWD_REQUIRE(d && table && status && keys, "null buffer");
  SweepGeom G;
  PMWD_REQUIRE(sweep_geom(d, ty, bw, 64, &G), "geometry not supported by the sweep scatter");
  cudaStream_t st = as_stream(stream);
  const int64_t nid = (int64_t)G.npencil * G.nband * G.nx_ext;
  PMWD_CUDA_TRY(cudaMemsetAsync(table, 0, (size_t)nid * sizeof(uint2), st));
  PMWD_CUDA_TRY(cudaMemsetAsync(status, 0, 16, st));
  if (G.n == 0) return PMWD_OK;
  const uint32_t key_div = (uint32_t)(ty * bw);
  const int grid = grid_for(G.n, 256, 8);
  sweep_table_kernel<<<grid, 256, 0, st>>>(G.n, keys, key_div, (uint2*)table);
  PMWD_LAUNCH_CHECK();
  sweep_table_check_kernel<<<grid, 256, 0, st>>>(G.n, keys, key_div, (const uint2*)table, (unsigned*)status);
  PMWD_LAUNCH_CHECK();
  sweep_table_cover_kernel<<<grid_for(nid, 256, 8), 256, 0, st>>>(nid, (const uint2*)table,
                                                                  (unsigned long long*)((char*)status + 8));
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

namespace pmwd {

bool sweep_usable(const pmwd_cic_desc* d, const pmwd_sweep* sw) {
  if (!sw || !sw->table || !sw->scratch) return false;
  SweepGeom G;
  if (!sweep_geom(d, sw->ty, sw->bw, sw->lx, &G)) return false;
  if (sw->nx_ext != G.nx_ext || sw->xoff != G.xoff) return false;      // table built for another slab
  return sw->scratch_bytes >= sweep_scratch_bytes(d->ptcl_num);
}

}  // namespace pmwd

// 1 if `sweep` (table + scratch) matches the mesh descriptor, i.e. pmwd_scatter_sweep / the sweep argument of
// pmwd_force* would use the tiled kernels; 0: the per-particle RED kernel is the one to call.
extern "C" int pmwd_sweep_usable(const pmwd_cic_desc* d, const pmwd_sweep* sweep) {
  return (d && sweep && pmwd::sweep_usable(d, sweep)) ? 1 : 0;
}

namespace pmwd {

// One channel: mesh (NOT pre-zeroed by the caller) <- deposit of val[p * vstride] (or vscalar).
// reuse_stragglers: the list recorded by an earlier call with the same particles is used again.
template <bool DET>
static int sweep_launch(cudaStream_t st, const SweepGeom& G, int grid, size_t smem, const short* pmid, const float* disp,
                        const float* val, int vstride, float vscalar, float* mesh, const uint2* table,
                        unsigned* counters, uint32_t* strag, int record, float* halo) {
  auto kernel = scatter_sweep_kernel<0, 0, DET>;
  if (G.ty == 8 && G.bw == 64) kernel = scatter_sweep_kernel<8, 64, DET>;
  else if (G.ty == 8 && G.bw == 32) kernel = scatter_sweep_kernel<8, 32, DET>;
  else if (G.ty == 16 && G.bw == 32) kernel = scatter_sweep_kernel<16, 32, DET>;
  static int smem_set = 0;
  if (!smem_set) {
    PMWD_CUDA_TRY(cudaFuncSetAttribute(scatter_sweep_kernel<0, 0, DET>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    PMWD_CUDA_TRY(cudaFuncSetAttribute(scatter_sweep_kernel<8, 64, DET>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    PMWD_CUDA_TRY(cudaFuncSetAttribute(scatter_sweep_kernel<8, 32, DET>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    PMWD_CUDA_TRY(cudaFuncSetAttribute(scatter_sweep_kernel<16, 32, DET>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    smem_set = 1;
  }
  kernel<<<grid, G.nwarps * 32, smem, st>>>(G, pmid, disp, val, vstride, vscalar, mesh, table, counters, strag, record, halo);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

// bytes of the per-tile halo arrays of the deterministic deposit: halo row [bw] + halo column and corner [ty + 1]
// per (tile, plane)
static size_t sweep_det_halo_bytes(const SweepGeom& G) {
  return (size_t)G.npencil * G.nband * G.nx * (G.bw + G.ty + 1) * sizeof(float);
}

bool sweep_det_usable(const pmwd_cic_desc* d, const pmwd_sweep* sw) {
  if (!sweep_usable(d, sw) || !sw->det_halo) return false;
  SweepGeom G;
  if (!sweep_geom(d, sw->ty, sw->bw, d->mesh_shape[0], &G)) return false;
  return G.periodic && G.nseg == 1 && sw->det_halo_bytes >= sweep_det_halo_bytes(G);
}

int scatter_sweep(cudaStream_t st, const pmwd_cic_desc* d, const pmwd_sweep* sw, const void* pmid,
                  const float* disp, const float* val, int vstride, float vscalar, float* mesh,
                  bool reuse_stragglers, bool deterministic) {
  SweepGeom G;
  // deterministic: one segment over all planes (nothing is shared between work items along x)
  PMWD_REQUIRE(sweep_geom(d, sw->ty, sw->bw, deterministic ? d->mesh_shape[0] : sw->lx, &G),
               "geometry not supported by the sweep scatter");
  PMWD_REQUIRE(!deterministic || sweep_det_usable(d, sw), "deterministic sweep needs the whole periodic mesh and its halo arrays");
  unsigned* counters = (unsigned*)sw->scratch;
  uint32_t* strag = (uint32_t*)((char*)sw->scratch + 64);
  const size_t smem = (size_t)G.nwarps * sweep_ring_bytes(G.ty, G.bw);
  PMWD_CUDA_TRY(cudaMemsetAsync(counters, 0, reuse_stragglers ? 4 : 12, st));
  if (!deterministic) {
    const int64_t rows = (int64_t)G.nx_ext * G.ny;
    sweep_zero_kernel<<<grid_for(rows * 32, 256, 8), 256, 0, st>>>(G, mesh);
    PMWD_LAUNCH_CHECK();
  }
  const int64_t nitems = (int64_t)G.npencil * G.nband * G.nseg;
  const int64_t want = (nitems + G.nwarps - 1) / G.nwarps;
  const int grid = (int)(want < sm_count() ? want : sm_count());
  PMWD_REQUIRE((int64_t)grid * G.nwarps * SW_STRAG_BLOCK <= SW_STRAG_SLACK, "straggler list slack too small for this grid");
  int rc;
  if (deterministic)
    rc = sweep_launch<true>(st, G, grid, smem, (const short*)pmid, disp, val, vstride, vscalar, mesh,
                            (const uint2*)sw->table, counters, strag, reuse_stragglers ? 0 : 1, (float*)sw->det_halo);
  else
    rc = sweep_launch<false>(st, G, grid, smem, (const short*)pmid, disp, val, vstride, vscalar, mesh,
                             (const uint2*)sw->table, counters, strag, reuse_stragglers ? 0 : 1, nullptr);
  if (rc) return rc;
  if (deterministic) {
    sweep_det_fix_kernel<false><<<grid_for((int64_t)G.nx * G.npencil * G.nz, 256, 8), 256, 0, st>>>(G, mesh, (const float*)sw->det_halo);
    PMWD_LAUNCH_CHECK();
    sweep_det_fix_kernel<true><<<grid_for((int64_t)G.nx * G.ny * G.nband, 256, 8), 256, 0, st>>>(G, mesh, (const float*)sw->det_halo);
    PMWD_LAUNCH_CHECK();
  }
  // (deterministic: the list is empty for freshly sorted storage; anything in it is deposited all the same and
  // counted in counters[3], which pmwd_sweep_det_violations reports)
  // 32 registers: 8 CTAs of 256 threads per SM; each thread has one scattered particle read in flight, so the
  // number of resident threads is the memory-level parallelism of this latency-bound pass
  sweep_straggler_kernel<<<sm_count() * 8, 256, 0, st>>>(G, (const short*)pmid, disp, val, vstride, vscalar, mesh,
                                                         counters, strag);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

}  // namespace pmwd

// Standalone entry (tests, slab composition): nch = 1 (m0) or 3 (SoA meshes m0, m1, m2 with
// val = float[N][3]).  The meshes are OVERWRITTEN (no memset needed).
extern "C" int pmwd_scatter_sweep(void* stream, const pmwd_cic_desc* d, const pmwd_sweep* sweep, const void* pmid,
                                  const float* disp, const float* val, float val_scalar, int nch, float* m0,
                                  float* m1, float* m2) {
  PMWD_REQUIRE(d && sweep && m0 && (d->ptcl_num == 0 || (pmid && disp)), "null buffer");
  PMWD_REQUIRE(nch == 1 || (nch == 3 && m1 && m2 && val), "nch must be 1 or 3 (three meshes, per-particle values)");
  PMWD_REQUIRE(sweep_usable(d, sweep), "sweep descriptor does not match the mesh descriptor");
  cudaStream_t st = as_stream(stream);
  StageTimer t(nch == 1 ? ST_SCATTER : ST_SCATTER3, st);
  if (nch == 1) return scatter_sweep(st, d, sweep, pmid, disp, val, 1, val_scalar, m0, false, false);
  float* m[3] = {m0, m1, m2};
  for (int c = 0; c < 3; ++c) {
    int rc = scatter_sweep(st, d, sweep, pmid, disp, val + c, 3, 0.f, m[c], c > 0, false);
    if (rc) return rc;
  }
  return PMWD_OK;
}

// The same through the deterministic variant (bitwise reproducible for freshly sorted storage; the sweep
// descriptor must carry the halo arrays, pmwd_sweep_det_halo_bytes).
extern "C" int pmwd_scatter_sweep_det(void* stream, const pmwd_cic_desc* d, const pmwd_sweep* sweep, const void* pmid,
                                      const float* disp, const float* val, float val_scalar, int nch, float* m0,
                                      float* m1, float* m2) {
  PMWD_REQUIRE(d && sweep && m0 && (d->ptcl_num == 0 || (pmid && disp)), "null buffer");
  PMWD_REQUIRE(nch == 1 || (nch == 3 && m1 && m2 && val), "nch must be 1 or 3 (three meshes, per-particle values)");
  PMWD_REQUIRE(sweep_det_usable(d, sweep), "sweep descriptor does not support the deterministic deposit on this mesh");
  cudaStream_t st = as_stream(stream);
  StageTimer t(nch == 1 ? ST_SCATTER : ST_SCATTER3, st);
  if (nch == 1) return scatter_sweep(st, d, sweep, pmid, disp, val, 1, val_scalar, m0, false, true);
  float* m[3] = {m0, m1, m2};
  for (int c = 0; c < 3; ++c) {
    int rc = scatter_sweep(st, d, sweep, pmid, disp, val + c, 3, 0.f, m[c], c > 0, true);
    if (rc) return rc;
  }
  return PMWD_OK;
}

// Size of the halo arrays the deterministic deposit needs for this mesh and tile shape (0: not supported).
extern "C" size_t pmwd_sweep_det_halo_bytes(const pmwd_cic_desc* d, int ty, int bw) {
  if (!d) return 0;
  SweepGeom G;
  if (!sweep_geom(d, ty, bw, d->mesh_shape[0], &G) || !G.periodic) return 0;
  return sweep_det_halo_bytes(G);
}

// Particles that a deterministic deposit had to hand to the (order dependent) straggler path since the scratch
// area was last cleared by pmwd_sweep_det_reset: 0 for freshly sorted storage.  Synchronises.
extern "C" long long pmwd_sweep_det_violations(void* stream, const pmwd_sweep* sweep) {
  if (!sweep || !sweep->scratch) return -1;
  unsigned c[4] = {0, 0, 0, 0};
  if (cudaStreamSynchronize(as_stream(stream)) != cudaSuccess) return -1;
  if (cudaMemcpy(c, sweep->scratch, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (long long)c[3];
}

extern "C" int pmwd_sweep_det_reset(void* stream, const pmwd_sweep* sweep) {
  PMWD_REQUIRE(sweep && sweep->scratch, "null buffer");
  PMWD_CUDA_TRY(cudaMemsetAsync((char*)sweep->scratch + 12, 0, 4, as_stream(stream)));
  return PMWD_OK;
}

// The number of stragglers of the last sweep that recorded them (reads device memory: synchronises).
extern "C" long long pmwd_sweep_last_stragglers(void* stream, const pmwd_sweep* sweep) {
  if (!sweep || !sweep->scratch) return -1;
  unsigned c[3] = {0, 0, 0};
  if (cudaStreamSynchronize(as_stream(stream)) != cudaSuccess) return -1;
  if (cudaMemcpy(c, sweep->scratch, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (long long)c[2];
}
