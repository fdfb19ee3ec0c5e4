// Tiled CIC deposit through shared-memory mesh tiles ("pencil sweep"): pmwd/scatter.py:60-83 for
// the gravity fast path (3-D, int16 pmid, offset = whole planes, cell_size=None).
//
// The per-particle global-RED kernel (cic_fast.cu) is bound by the LSU's atomic issue rate
// (~1.3 cycles per lane: 31 cycles per RED warp instruction measured, profiles/r01_final_ncu_full.txt)
// and moves 2.3x its algorithmic DRAM bytes (memset + read-modify-write of the whole mesh).  Shared
// memory atomics are no way out on sm_100a (ATOMS: 2 cycles per lane, float add = CAS loop).  This
// kernel instead accumulates with PLAIN shared-memory read-modify-writes made conflict free by
// construction, and writes every mesh cell that only one CTA can touch with coalesced float4
// STORES (no memset, no read):
//
//  * the integrator keeps its particle storage sorted by (y-pencil, x-plane, y-pair, z)
//    (reorder.cu; Lagrangian order has the same segment structure), and a table gives the
//    particle range of every (pencil, plane);
//  * a work item = one pencil (TY rows, all of z) x one segment of LX planes.  The CTA sweeps the
//    segment plane by plane with a RING of four planes [(TY+1) rows][nz] in shared memory: a particle
//    sorted at plane xs now sits at xs-1 .. xs+1 (it moved since the last re-sort) and touches
//    planes xs-1 .. xs+2.  When plane xs is done, plane xs-1 is complete and is flushed;
//  * each batch of 512 particles is binned by z-band (16 bands, one per warp): warp w alone updates
//    band w of the ring, eight sequential phases (one per neighbour) of plain LDS/FADD/STS after the
//    lanes that share a base cell have been merged (reduce_peers); only the first cell of every band,
//    which the previous band's last cell spills into, takes shared atomics;
//  * flush: rows 1 .. TY-1 of the planes only this item can reach -> float4 stores; the pencil's
//    first row and halo row (shared with the neighbouring pencils) and the three planes at either
//    end of the segment (shared with the neighbouring segments) -> red.global.add.v4.f32 onto cells
//    that sweep_zero_kernel cleared beforehand (1/TY of the rows + 3 planes per LX);
//  * particles outside their window (moved more than one plane, or out of the pencil in y) are
//    "stragglers": appended to a list and deposited by a second, per-particle RED kernel that runs
//    after this one (so that it can never race with the plain stores).
//
// Correctness never depends on how fresh the sort is; only the straggler fraction does.
#include <cooperative_groups.h>

#include "cic.cuh"

namespace pmwd {

int slab_xoff(const pmwd_cic_desc* d);
bool cic_is_fast(const pmwd_cic_desc* d);

constexpr int SW_THREADS = 512;
constexpr int SW_WARPS = SW_THREADS / 32;

struct SweepGeom {
  int64_t n;
  int nx, ny, nz;     // periodic wrap shape (the GLOBAL mesh on a slab)
  int nx_ext, xoff;   // planes held by the mesh array, global index of plane 0
  int periodic;       // the array spans the whole periodic x axis
  float cell;
  int ty, npencil;    // rows per pencil, ny / ty
  int lx, nseg;       // planes per x segment, ceil(nx_ext / lx)
  int bw_shift;       // z band width = 1 << bw_shift cells (<= SW_WARPS bands)
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

__global__ void __launch_bounds__(SW_THREADS, 1)
scatter_sweep_kernel(SweepGeom G, const short* __restrict__ pmid, const float* __restrict__ disp,
                     const float* __restrict__ val, int vstride, float vscalar, float* __restrict__ mesh,
                     const uint2* __restrict__ table, unsigned* __restrict__ counters,
                     uint32_t* __restrict__ strag, int record_strag) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int plane_sz = (G.ty + 1) * G.nz;               // floats per ring plane (multiple of 4)
  float* ring = reinterpret_cast<float*>(smraw);        // [4][ty+1][nz]
  float4* rec = reinterpret_cast<float4*>(ring + 4 * (size_t)plane_sz);   // [SW_THREADS]
  float* recv = reinterpret_cast<float*>(rec + SW_THREADS);               // [SW_THREADS]
  int* cnt = reinterpret_cast<int*>(recv + SW_THREADS);                   // [SW_WARPS]
  uint2* segs = reinterpret_cast<uint2*>(cnt + SW_WARPS);                 // [lx]
  __shared__ int s_item;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nitems = G.npencil * G.nseg;
  const int bw_mask = (1 << G.bw_shift) - 1;
  if (tid < SW_WARPS) cnt[tid] = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = (int)atomicAdd(&counters[0], 1u);
    __syncthreads();
    const int item = s_item;
    if (item >= nitems) break;
    const int pencil = item % G.npencil, sgi = item / G.npencil;
    const int xa = sgi * G.lx, xb = min(xa + G.lx, G.nx_ext);
    const int y0 = pencil * G.ty;
    for (int i = tid; i < xb - xa; i += SW_THREADS) segs[i] = table[(int64_t)pencil * G.nx_ext + xa + i];
    {
      float4* r4 = reinterpret_cast<float4*>(ring);
      for (int i = tid; i < plane_sz; i += SW_THREADS) r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // 4 planes / 4
    }
    __syncthreads();

    // ---- flush one ring plane to the mesh and clear it
    auto flush = [&](int pl) {
      const int slot = (pl - (xa - 1)) & 3;
      float4* src = reinterpret_cast<float4*>(ring + (size_t)slot * plane_sz);
      bool valid = true;
      int gpl = pl;
      if (G.periodic) {
        if (gpl < 0) gpl += G.nx;
        else if (gpl >= G.nx) gpl -= G.nx;
      } else {
        valid = pl >= 0 && pl < G.nx_ext;
      }
      const bool all_red = (pl <= xa + 1) || (pl >= xb - 1);
      const int nz4 = G.nz >> 2;
      for (int i = tid; i < (plane_sz >> 2); i += SW_THREADS) {
        const float4 v = src[i];
        src[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!valid) continue;
        const int row = i / nz4, z4 = i - row * nz4;
        int gy = y0 + row;
        if (gy == G.ny) gy = 0;
        float* dst = mesh + ((int64_t)gpl * G.ny + gy) * G.nz + 4 * z4;
        if (all_red || row == 0 || row == G.ty) {
          if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add4(dst, v);
        } else {
          __stcs(reinterpret_cast<float4*>(dst), v);
        }
      }
    };

    // ---- batch iterator over the item's (plane, particle range) list; uniform across the CTA
    int xs = xa;
    unsigned b0 = 0, bend = 0;
    auto seek = [&]() {      // make (xs, b0, bend) point at a non-empty batch or xs == xb
      while (xs < xb) {
        if (b0 < bend) return;
        ++xs;
        if (xs < xb) { const uint2 s = segs[xs - xa]; b0 = s.x; bend = s.y; }
      }
    };
    { const uint2 s = segs[0]; b0 = s.x; bend = s.y; }
    seek();
    auto load = [&](int lxs, unsigned lb0, unsigned lbend, PtclRegs& R) -> bool {
      const unsigned p = lb0 + tid;
      const bool act = lxs < xb && p < lbend;
      if (act) {
        R.pm[0] = pmid[3 * (int64_t)p + 0]; R.pm[1] = pmid[3 * (int64_t)p + 1]; R.pm[2] = pmid[3 * (int64_t)p + 2];
        R.dp[0] = disp[3 * (int64_t)p + 0]; R.dp[1] = disp[3 * (int64_t)p + 1]; R.dp[2] = disp[3 * (int64_t)p + 2];
        R.v = val ? val[(int64_t)vstride * p] : vscalar;
      }
      return act;
    };
    PtclRegs cur, nxt;
    bool cur_act = load(xs, b0, bend, cur);
    int next_flush = xa - 1;

    while (xs < xb) {
      const int cxs = xs;
      const unsigned cb0 = b0;
      // advance the iterator and prefetch the next batch before working on this one
      b0 += SW_THREADS;
      seek();
      const bool nxt_act = load(xs, b0, bend, nxt);

      if (next_flush <= cxs - 2) {          // planes that can no longer be reached
        while (next_flush <= cxs - 2) { flush(next_flush); ++next_flush; }
        __syncthreads();
      }

      // ---- stencil of this thread's particle
      const unsigned p = cb0 + tid;
      int g[3];
      float d0[3];
      const int nn[3] = {G.nx, G.ny, G.nz};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float t = __fdiv_rn(cur_act ? cur.dp[a] : 0.f, G.cell);
        const int i0 = (int)floorf(t);
        d0[a] = __fsub_rn(t, (float)i0);
        g[a] = wrap_index((cur_act ? (int)cur.pm[a] : 0) + i0, nn[a]);
      }
      int lp = g[0] - G.xoff;
      if (lp < 0) lp += G.nx;
      int dxs = lp - cxs;
      if (G.periodic) {
        if (dxs > (G.nx >> 1)) dxs -= G.nx;
        else if (dxs < -(G.nx >> 1)) dxs += G.nx;
      }
      const int dy = g[1] - y0;
      const bool inwin = cur_act && dxs >= -1 && dxs <= 1 && dy >= 0 && dy < G.ty;
      if (record_strag) {
        const unsigned sm = __ballot_sync(0xffffffffu, cur_act && !inwin);
        if (sm) {
          const int leader = __ffs(sm) - 1;
          unsigned base = 0;
          if (lane == leader) base = atomicAdd(&counters[1], (unsigned)__popc(sm));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (cur_act && !inwin) strag[base + __popc(sm & ((1u << lane) - 1u))] = p;
        }
      }
      // ---- bin by z band: slot inside the band's queue
      const int band = g[2] >> G.bw_shift;
      int myoff = 0;
      {
        unsigned todo = __ballot_sync(0xffffffffu, inwin);
        while (todo) {
          const int leader = __ffs(todo) - 1;
          const int b = __shfl_sync(0xffffffffu, band, leader);
          const unsigned m = __ballot_sync(0xffffffffu, inwin && band == b);
          int base = 0;
          if (lane == leader) base = atomicAdd(&cnt[b], __popc(m));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (inwin && band == b) myoff = base + __popc(m & ((1u << lane) - 1u));
          todo &= ~m;
        }
      }
      __syncthreads();                                                     // S1: counts final
      int qbeg = 0, qlen = 0, mybase = 0;
#pragma unroll
      for (int b = 0; b < SW_WARPS; ++b) {
        const int c = cnt[b];
        if (b < warp) qbeg += c;
        if (b == warp) qlen = c;
        if (b < band) mybase += c;
      }
      if (inwin) {
        const int slot = (cxs + dxs - (xa - 1)) & 3;
        const int packed = (slot << 28) | (dy << 20) | g[2];
        rec[mybase + myoff] = make_float4(d0[0], d0[1], d0[2], __int_as_float(packed));
        recv[mybase + myoff] = cur.v;
      }
      __syncthreads();                                                     // S2: records visible
      if (tid < SW_WARPS) cnt[tid] = 0;

      // ---- warp `warp` deposits the particles of its band
      for (int c = 0; c < qlen; c += 32) {
        const bool act = c + lane < qlen;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        float v = 0.f;
        if (act) { r = rec[qbeg + c + lane]; v = recv[qbeg + c + lane]; }
        const int packed = __float_as_int(r.w);
        const unsigned key = act ? (unsigned)packed : (0xc0000000u | (unsigned)lane);   // bits 30-31 set: never a real cell
        float wx[2], wy[2], wz[2];
        wx[0] = __fsub_rn(1.f, fabsf(r.x)); wx[1] = __fsub_rn(1.f, fabsf(__fsub_rn(r.x, 1.f)));
        wy[0] = __fsub_rn(1.f, fabsf(r.y)); wy[1] = __fsub_rn(1.f, fabsf(__fsub_rn(r.y, 1.f)));
        wz[0] = __fsub_rn(1.f, fabsf(r.z)); wz[1] = __fsub_rn(1.f, fabsf(__fsub_rn(r.z, 1.f)));
        float c8[8];
#pragma unroll
        for (int bx = 0; bx < 2; ++bx)
#pragma unroll
          for (int by = 0; by < 2; ++by) {
            const float wxy = __fmul_rn(wx[bx], wy[by]);
#pragma unroll
            for (int bz = 0; bz < 2; ++bz)
              c8[(bx * 2 + by) * 2 + bz] = act ? __fmul_rn(v, __fmul_rn(wxy, wz[bz])) : 0.f;
          }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const bool first = sweep_reduce_peers<8>(peers, c8);
        const bool push = act && first;
        const int slot = (packed >> 28) & 3, row = (packed >> 20) & 0xff, z = packed & 0xfffff;
        const int z1 = (z + 1 == G.nz) ? 0 : z + 1;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const int bx = n >> 2, by = (n >> 1) & 1, bz = n & 1;
          if (push) {
            const int zt = bz ? z1 : z;
            float* cellp = ring + (size_t)((slot + bx) & 3) * plane_sz + (row + by) * G.nz + zt;
            if ((zt & bw_mask) == 0) atomicAdd(cellp, c8[n]);     // a band's first cell: shared with the band below
            else *cellp = *cellp + c8[n];
          }
          __syncwarp();
        }
      }
      __syncthreads();                                                     // S3: ring quiescent

      cur = nxt;
      cur_act = nxt_act;
    }
    // ---- the rest of the ring: planes up to xb + 1
    while (next_flush <= xb + 1) { flush(next_flush); ++next_flush; }
  }
}

// Cells that receive vector REDs must start at zero: the first row of every pencil (it is also the
// halo row of the pencil below) on every plane, and whole planes b-1, b, b+1 at every segment
// boundary b.  One warp per mesh row.
__global__ void __launch_bounds__(256)
sweep_zero_kernel(SweepGeom G, float* __restrict__ mesh) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t rows = (int64_t)G.nx_ext * G.ny;
  const int nz4 = G.nz >> 2;
  for (int64_t r = warp; r < rows; r += nwarp) {
    const int pl = (int)(r / G.ny), y = (int)(r - (int64_t)pl * G.ny);
    const int m = pl % G.lx;
    const bool zero = (y % G.ty == 0) || m == 0 || m == 1 || m == G.lx - 1 || pl == G.nx_ext - 1;
    if (!zero) continue;
    float4* row = reinterpret_cast<float4*>(mesh + r * G.nz);
    for (int i = lane; i < nz4; i += 32) row[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Stragglers: per-particle global REDs (drop rule of enmesh for slabs), after the sweep.
__global__ void __launch_bounds__(256)
sweep_straggler_kernel(SweepGeom G, const short* __restrict__ pmid, const float* __restrict__ disp,
                       const float* __restrict__ val, int vstride, float vscalar, float* __restrict__ mesh,
                       const unsigned* __restrict__ counters, const uint32_t* __restrict__ strag) {
  const unsigned n = counters[1];
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
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
// (pencil, plane) segment table.  id(i) = pencil * nx_ext + plane of particle slot i, either from the
// keys of the last cell sort (reorder.cu, sweep key layout) or from pmid alone (Lagrangian order:
// the particle grid's C order has one contiguous run per (plane, pencil)).
__global__ void __launch_bounds__(256)
sweep_table_kernel(int64_t n, const uint32_t* __restrict__ keys, uint32_t key_div,
                   const short* __restrict__ pmid, SweepGeom G, uint2* __restrict__ table) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    auto id_of = [&](int64_t j) -> uint32_t {
      if (keys) return keys[j] / key_div;
      int lp = wrap_index((int)pmid[3 * j + 0], G.nx) - G.xoff;
      if (lp < 0) lp += G.nx;
      if (lp >= G.nx_ext) lp = G.nx_ext - 1;
      const int gy = wrap_index((int)pmid[3 * j + 1], G.ny);
      return (uint32_t)((gy / G.ty) * G.nx_ext + lp);
    };
    const uint32_t id = id_of(i);
    if (i == 0 || id_of(i - 1) != id) {
      table[id].x = (uint32_t)i;
      if (i > 0) table[id_of(i - 1)].y = (uint32_t)i;
    }
    if (i == n - 1) table[id].y = (uint32_t)n;
  }
}

// every particle slot must lie inside the range of its own id (together with one run per id this
// makes the table an exact partition): counts violations
__global__ void __launch_bounds__(256)
sweep_table_check_kernel(int64_t n, const uint32_t* __restrict__ keys, uint32_t key_div,
                         const short* __restrict__ pmid, SweepGeom G, const uint2* __restrict__ table,
                         unsigned* __restrict__ bad) {
  unsigned local = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t id;
    if (keys) {
      id = keys[i] / key_div;
    } else {
      int lp = wrap_index((int)pmid[3 * i + 0], G.nx) - G.xoff;
      if (lp < 0) lp += G.nx;
      if (lp >= G.nx_ext) lp = G.nx_ext - 1;
      id = (uint32_t)((wrap_index((int)pmid[3 * i + 1], G.ny) / G.ty) * G.nx_ext + lp);
    }
    const uint2 s = table[id];
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
static int ilog2_ceil_i(int x) { int b = 0; while ((1 << b) < x) ++b; return b; }

static size_t sweep_smem_bytes(const SweepGeom& G) {
  return (size_t)4 * (G.ty + 1) * G.nz * sizeof(float) + SW_THREADS * (sizeof(float4) + sizeof(float)) +
         SW_WARPS * sizeof(int) + (size_t)G.lx * sizeof(uint2) + 64;
}

// Fills G for (d, ty, lx); returns false if the geometry is not supported by the sweep kernels.
static bool sweep_geom(const pmwd_cic_desc* d, int ty, int lx, SweepGeom* G) {
  if (!cic_is_fast(d)) return false;
  G->n = d->ptcl_num;
  G->nx = d->wrap_shape[0]; G->ny = d->wrap_shape[1]; G->nz = d->wrap_shape[2];
  G->nx_ext = d->mesh_shape[0];
  G->xoff = slab_xoff(d);
  G->periodic = (G->nx_ext == G->nx && G->xoff == 0) ? 1 : 0;
  if (!G->periodic && G->nx_ext > G->nx) return false;      // halos wider than the box: RED kernel
  G->cell = (float)d->cell_size;
  if (ty < 2 || ty > 64 || G->ny % ty != 0 || (G->nz & 3) != 0 || G->nz > (1 << 20) || G->nz < 8) return false;
  if (lx < 1 || G->nx_ext < 4) return false;
  G->ty = ty;
  G->npencil = G->ny / ty;
  G->lx = lx < G->nx_ext ? lx : G->nx_ext;
  G->nseg = (G->nx_ext + G->lx - 1) / G->lx;
  G->bw_shift = ilog2_ceil_i((G->nz + SW_WARPS - 1) / SW_WARPS);
  if (G->bw_shift < 1) G->bw_shift = 1;
  if (d->ptcl_num >= ((int64_t)1 << 32)) return false;
  return sweep_smem_bytes(*G) <= 226 * 1024;      // 227 KB per CTA minus the kernel's static shared memory
}

}  // namespace pmwd

using namespace pmwd;

// Largest supported pencil height for this mesh (0: the sweep scatter cannot be used).
extern "C" int pmwd_sweep_pick_ty(const pmwd_cic_desc* d) {
  if (!d) return 0;
  SweepGeom G;
  for (int ty = 8; ty >= 2; ty >>= 1)
    if (sweep_geom(d, ty, 64, &G)) return ty;
  return 0;
}

extern "C" size_t pmwd_sweep_table_bytes(const pmwd_cic_desc* d, int ty) {
  if (!d || ty <= 0 || d->mesh_shape[1] % ty) return 0;
  return (size_t)(d->mesh_shape[1] / ty) * d->mesh_shape[0] * sizeof(uint2);
}

// counters (64 bytes) + straggler list (4 bytes per particle)
extern "C" size_t pmwd_sweep_scratch_bytes(const pmwd_cic_desc* d) {
  if (!d) return 0;
  return 64 + (size_t)d->ptcl_num * sizeof(uint32_t);
}

// Build the (pencil, plane) table.  `keys`: the sorted keys of pmwd_cell_sort_perm in the sweep key
// layout (ty = the sort's ty), or NULL to derive the segments from pmid (Lagrangian order).
// `status` (device, 16 bytes, zeroed here): [0] = particles outside their segment (unsigned),
// [8] = total length of all ranges (unsigned long long); the table is valid iff status[0] == 0 and
// total == ptcl_num.  Enqueue-only; the caller reads `status` after synchronising.
extern "C" int pmwd_sweep_table(void* stream, const pmwd_cic_desc* d, int ty, const uint32_t* keys,
                                const void* pmid, uint32_t* table, void* status) {
  PMWD_REQUIRE(d && table && status && (keys || pmid), "null buffer");
  SweepGeom G;
  PMWD_REQUIRE(sweep_geom(d, ty, 64, &G), "geometry not supported by the sweep scatter");
  cudaStream_t st = as_stream(stream);
  const int64_t nid = (int64_t)G.npencil * G.nx_ext;
  PMWD_CUDA_TRY(cudaMemsetAsync(table, 0, (size_t)nid * sizeof(uint2), st));
  PMWD_CUDA_TRY(cudaMemsetAsync(status, 0, 16, st));
  if (G.n == 0) return PMWD_OK;
  const uint32_t key_div = (uint32_t)((ty / 2) * G.nz);
  const int grid = grid_for(G.n, 256, 8);
  sweep_table_kernel<<<grid, 256, 0, st>>>(G.n, keys, key_div, (const short*)pmid, G, (uint2*)table);
  PMWD_LAUNCH_CHECK();
  sweep_table_check_kernel<<<grid, 256, 0, st>>>(G.n, keys, key_div, (const short*)pmid, G, (const uint2*)table,
                                                 (unsigned*)status);
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
  if (!sweep_geom(d, sw->ty, sw->lx, &G)) return false;
  if (sw->nx_ext != G.nx_ext || sw->xoff != G.xoff) return false;      // table built for another slab
  return sw->scratch_bytes >= 64 + (size_t)d->ptcl_num * sizeof(uint32_t);
}

// One channel: mesh (NOT pre-zeroed by the caller) <- deposit of val[p * vstride] (or vscalar).
// reuse_stragglers: the list recorded by an earlier call with the same particles is used again.
int scatter_sweep(cudaStream_t st, const pmwd_cic_desc* d, const pmwd_sweep* sw, const void* pmid,
                  const float* disp, const float* val, int vstride, float vscalar, float* mesh,
                  bool reuse_stragglers) {
  SweepGeom G;
  PMWD_REQUIRE(sweep_geom(d, sw->ty, sw->lx, &G), "geometry not supported by the sweep scatter");
  unsigned* counters = (unsigned*)sw->scratch;
  uint32_t* strag = (uint32_t*)((char*)sw->scratch + 64);
  static int smem_set = 0;
  const size_t smem = sweep_smem_bytes(G);
  if (!smem_set) {
    PMWD_CUDA_TRY(cudaFuncSetAttribute(scatter_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       226 * 1024));
    smem_set = 1;
  }
  PMWD_CUDA_TRY(cudaMemsetAsync(counters, 0, reuse_stragglers ? 4 : 8, st));
  const int64_t rows = (int64_t)G.nx_ext * G.ny;
  sweep_zero_kernel<<<grid_for(rows * 32, 256, 8), 256, 0, st>>>(G, mesh);
  PMWD_LAUNCH_CHECK();
  const int nitems = G.npencil * G.nseg;
  const int grid = nitems < sm_count() ? nitems : sm_count();
  scatter_sweep_kernel<<<grid, SW_THREADS, smem, st>>>(G, (const short*)pmid, disp, val, vstride, vscalar, mesh,
                                                       (const uint2*)sw->table, counters, strag,
                                                       reuse_stragglers ? 0 : 1);
  PMWD_LAUNCH_CHECK();
  sweep_straggler_kernel<<<sm_count() * 4, 256, 0, st>>>(G, (const short*)pmid, disp, val, vstride, vscalar, mesh,
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
  if (nch == 1) return scatter_sweep(st, d, sweep, pmid, disp, val, 1, val_scalar, m0, false);
  float* m[3] = {m0, m1, m2};
  for (int c = 0; c < 3; ++c) {
    int rc = scatter_sweep(st, d, sweep, pmid, disp, val + c, 3, 0.f, m[c], c > 0);
    if (rc) return rc;
  }
  return PMWD_OK;
}

// The number of stragglers of the last sweep that recorded them (reads device memory: synchronises).
extern "C" long long pmwd_sweep_last_stragglers(void* stream, const pmwd_sweep* sweep) {
  if (!sweep || !sweep->scratch) return -1;
  unsigned c[2] = {0, 0};
  if (cudaStreamSynchronize(as_stream(stream)) != cudaSuccess) return -1;
  if (cudaMemcpy(c, sweep->scratch, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (long long)c[1];
}
