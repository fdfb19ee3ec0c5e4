// Eulerian re-ordering of the particle storage.
//
// The reference keeps particles in Lagrangian order forever (pmwd/particles.py:135-139).  Once
// particles have moved several cells, a warp of 32 consecutive particles touches 32 unrelated
// mesh rows: ncu shows the 3-mesh gather moving 6x its algorithmic bytes and the scatter 4x
// (profiles/r01_ncu_full_late_step46.txt).  The integrator therefore periodically re-sorts its
// private copy of the particle arrays by mesh cell (stable LSD radix sort on the linear cell
// index, cub::DeviceRadixSort) and carries the Lagrangian index along so that outputs are
// returned in the reference's order.  Per-particle arithmetic is unchanged; only the order of
// the (already unordered) float32 scatter additions differs.
//
//   pmwd_cell_sort_perm : perm[i] = storage index of the particle that goes to sorted slot i
//   pmwd_permute_rows   : dst[i] = src[perm[i]]  (gather)   or  dst[perm[i]] = src[i]  (scatter)
#include <cub/device/device_radix_sort.cuh>

#include <stdlib.h>

#include "cic.cuh"

namespace pmwd {

int slab_xoff(const pmwd_cic_desc* d);
bool cic_is_fast(const pmwd_cic_desc* d);

struct SortParams {
  int64_t n;
  int nx, ny, nz;   // periodic wrap shape
  int nx_ext, xoff; // slab: planes held locally and global index of the first one
  int shift;        // xy footprint of a storage "column" is (1<<shift)^2 cells
  int nyc;          // ceil(ny >> shift)
  int ty, bw;       // > 0: sweep layout (y / ty, z / bw, x, y % ty, z % bw) -- scatter_sweep.cu
  float cell;
  // two sources (slab runs, pmwd_cell_sort_perm2): rows [0, nA) come from (pmid, disp), rows [nA, n) from
  // (pmidB, dispB); a row of A whose owner is another rank gets the largest key and sorts behind everything
  int64_t nA;
  const uint8_t* ownerA;
  int rank;
  // keys from PREDICTED positions disp + vel * pred (vel == NULL: the positions as they are): the storage order is
  // then centred on the force evaluations until the next re-sort instead of being exact now and 2 steps stale last
  const float* vel;
  const float* velB;
  float pred;
};

constexpr uint32_t SORT_KEY_GONE = 0xffffffffu;

// TWO = false: one source, nobody left (the single-device re-sort): the plain loads the compiler vectorises
// (the two-source pointer selects cost this memory-bound kernel 50 %: 0.89 -> 1.36 ms at 512^3)
template <bool TWO>
__global__ void __launch_bounds__(256)
sort_keys_kernel(SortParams P, const short* __restrict__ pmid, const float* __restrict__ disp,
                 const short* __restrict__ pmidB, const float* __restrict__ dispB,
                 uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P.n;
       p += (int64_t)gridDim.x * blockDim.x) {
    int c[3];
    const int nn[3] = {P.nx, P.ny, P.nz};
    if (TWO) {
      vals[p] = (uint32_t)p;
      const bool fromA = p < P.nA;
      if (fromA && P.ownerA && P.ownerA[p] != (uint8_t)P.rank) { keys[p] = SORT_KEY_GONE; continue; }
      const short* pm = fromA ? pmid + 3 * p : pmidB + 3 * (p - P.nA);
      const float* dp = fromA ? disp + 3 * p : dispB + 3 * (p - P.nA);
      const float* vp = P.vel ? (fromA ? P.vel + 3 * p : P.velB + 3 * (p - P.nA)) : nullptr;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float x = vp ? __fadd_rn(dp[a], __fmul_rn(vp[a], P.pred)) : dp[a];
        float t = __fdiv_rn(x, P.cell);
        c[a] = wrap_index((int)pm[a] + (int)floorf(t), nn[a]);
      }
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float x = P.vel ? __fadd_rn(disp[3 * p + a], __fmul_rn(P.vel[3 * p + a], P.pred)) : disp[3 * p + a];
        float t = __fdiv_rn(x, P.cell);
        c[a] = wrap_index((int)pmid[3 * p + a] + (int)floorf(t), nn[a]);
      }
    }
    int lx = c[0] - P.xoff;           // slab-local plane keeps the key below 2^32 on big meshes
    if (lx < 0) lx += P.nx;
    if (lx >= P.nx_ext) lx = P.nx_ext - 1;
    // columns of (1<<shift)^2 cells in (x, y), ordered by z inside: consecutive particles are
    // as dense along the contiguous z axis as the Lagrangian lattice was, which is what makes
    // a warp's 32 stencils share 32-byte sectors
    if (P.ty > 0) {
      const int pencil = c[1] / P.ty, yin = c[1] - pencil * P.ty;
      const int band = c[2] / P.bw, zin = c[2] - band * P.bw;
      const int64_t id = ((int64_t)pencil * (P.nz / P.bw) + band) * P.nx_ext + lx;
      keys[p] = (uint32_t)((id * P.ty + yin) * P.bw + zin);
    } else {
      keys[p] = (uint32_t)(((int64_t)(lx >> P.shift) * P.nyc + (c[1] >> P.shift)) * P.nz + c[2]);
    }
    if (!TWO) vals[p] = (uint32_t)p;
  }
}

struct RowArgs {
  int narr;
  const void* src[8];
  const void* srcB[8];   // second source (gather only): rows nA .. of the virtual concatenation; may be null
  int64_t nA;
  void* dst[8];
  int words[8];     // row size in 2-byte words (3 for int16[3], 6 for float[3], 2 for uint32)
};

template <bool INVERSE, bool TWO>
__global__ void __launch_bounds__(256)
permute_rows_kernel(int64_t n, const uint32_t* __restrict__ perm, RowArgs A) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = perm[i];
    int64_t from = INVERSE ? i : j;
    const int64_t to = INVERSE ? j : i;
    const bool second = TWO && !INVERSE && from >= A.nA;
    if (second) from -= A.nA;
    for (int a = 0; a < A.narr; ++a) {
      const int w = A.words[a];
      const void* base = (TWO && second) ? A.srcB[a] : A.src[a];
      if ((w & 1) == 0) {   // rows of 4-byte words
        const uint32_t* s = reinterpret_cast<const uint32_t*>(base) + from * (w >> 1);
        uint32_t* d = reinterpret_cast<uint32_t*>(A.dst[a]) + to * (w >> 1);
        for (int k = 0; k < (w >> 1); ++k) d[k] = __ldg(s + k);
      } else {
        const uint16_t* s = reinterpret_cast<const uint16_t*>(base) + from * w;
        uint16_t* d = reinterpret_cast<uint16_t*>(A.dst[a]) + to * w;
        for (int k = 0; k < w; ++k) d[k] = __ldg(s + k);
      }
    }
  }
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static int ilog2_ceil(int64_t x) { int b = 0; while (((int64_t)1 << b) < x) ++b; return b; }

struct SortLayout { size_t keys_in, vals_in, keys_out, cub, total, cub_bytes; };

static int sort_layout(int64_t n, int end_bit, SortLayout* L) {
  size_t cub_bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr,
                                                  (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                                  (uint32_t*)nullptr, n, 0, end_bit);
  if (e != cudaSuccess) { set_error("cub temp query failed: %s", cudaGetErrorString(e)); return (int)e; }
  size_t off = 0;
  L->keys_in = off;  off += align_up((size_t)n * 4);
  L->vals_in = off;  off += align_up((size_t)n * 4);
  L->keys_out = off; off += align_up((size_t)n * 4);
  L->cub = off;      off += align_up(cub_bytes);
  L->cub_bytes = cub_bytes;
  L->total = off;
  return PMWD_OK;
}

}  // namespace pmwd

using namespace pmwd;

extern "C" size_t pmwd_cell_sort_scratch_bytes(const pmwd_cic_desc* d) {
  if (!d || d->dim != 3) return 0;
  int64_t ncell = (int64_t)d->mesh_shape[0] * d->mesh_shape[1] * d->mesh_shape[2];
  SortLayout L;
  (void)ncell;
  if (sort_layout(d->ptcl_num, 32, &L)) return 0;      // sized for full-width keys (pmwd_cell_sort_perm2's mask)
  return L.total;
}

extern "C" const uint32_t* pmwd_cell_sort_sorted_keys(const pmwd_cic_desc* d, const void* scratch) {
  if (!d || !scratch) return nullptr;
  return (const uint32_t*)((const char*)scratch + 2 * align_up((size_t)d->ptcl_num * 4));
}

// Two sources (slab runs with Eulerian ownership): the rows to sort are the virtual concatenation of
// A = (pmid, disp)[0, nA) -- of which the rows with ownerA[i] != rank have left for another rank -- and the
// arrivals B = (pmidB, dispB)[0, nB).  d->ptcl_num must be nA + nB (it sizes the scratch layout).  perm has
// nA + nB entries; the rows that left sort to the end, i.e. the first nA + nB - (number gone) entries are the
// new storage order.  ownerA == NULL: nobody left; nB == 0: no arrivals.
extern "C" int pmwd_cell_sort_perm2(void* stream, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                                    int64_t nA, const uint8_t* ownerA, int rank, const void* pmidB,
                                    const float* dispB, uint32_t* perm, void* scratch, size_t scratch_bytes,
                                    int ty, int bw, const float* vel, const float* velB, float pred) {
  PMWD_REQUIRE(d && d->dim == 3 && d->pmid_bytes == 2 && !d->general,
               "cell sort supports the 3-D int16 fast path");
  PMWD_REQUIRE(perm && scratch, "null buffer");
  SortParams P;
  P.n = d->ptcl_num;
  PMWD_REQUIRE(nA >= 0 && nA <= P.n, "nA must lie in [0, ptcl_num]");
  PMWD_REQUIRE(nA == P.n || (pmidB && dispB), "null second source");
  P.nA = nA;
  P.ownerA = ownerA;
  P.rank = rank;
  const bool predict = vel != nullptr && pred != 0.f && (nA == d->ptcl_num || velB != nullptr);
  P.vel = predict ? vel : nullptr;
  P.velB = predict ? velB : nullptr;
  P.pred = predict ? pred : 0.f;
  P.nx = d->wrap_shape[0]; P.ny = d->wrap_shape[1]; P.nz = d->wrap_shape[2];
  P.nx_ext = d->mesh_shape[0];
  P.xoff = slab_xoff(d);
  P.cell = (float)d->cell_size;
  PMWD_REQUIRE(ty >= 0 && (ty == 0 || (bw > 0 && d->wrap_shape[1] % ty == 0 && d->wrap_shape[2] % bw == 0)),
               "sweep key layout needs tile sizes that divide the y and z extents");
  P.ty = ty;
  P.bw = bw;
  {
    const char* e = getenv("PMWD_SORT_SHIFT");
    P.shift = e ? atoi(e) : 1;
    if (P.shift < 0 || P.shift > 4) P.shift = 1;
    P.nyc = (P.ny + (1 << P.shift) - 1) >> P.shift;
  }
  PMWD_REQUIRE(cic_is_fast(d), "cell sort needs a fast-path (slab) descriptor");
  int64_t ncell = (int64_t)P.nx_ext * P.ny * P.nz;
  PMWD_REQUIRE(ncell <= ((int64_t)1 << 32) && P.n < ((int64_t)1 << 32),
               "cell sort needs mesh_size <= 2^32 and ptcl_num < 2^32 per device");
  PMWD_REQUIRE(!ownerA || ncell < ((int64_t)1 << 32), "the ownership mask needs mesh_size < 2^32 (one spare key)");
  if (P.n == 0) return PMWD_OK;
  PMWD_REQUIRE(nA == 0 || (pmid && disp), "null buffer");
  int end_bit = ownerA ? 32 : ilog2_ceil(ncell);
  if (end_bit < 1) end_bit = 1;
  SortLayout L;
  int rc = sort_layout(P.n, 32, &L);     // as pmwd_cell_sort_scratch_bytes
  if (rc) return rc;
  if (scratch_bytes < L.total) {
    set_error("cell sort needs %zu bytes of scratch, got %zu", L.total, scratch_bytes);
    return PMWD_ENOMEM;
  }
  cudaStream_t st = as_stream(stream);
  StageTimer timer(ST_OTHER, st);
  char* base = (char*)scratch;
  uint32_t* keys_in = (uint32_t*)(base + L.keys_in);
  uint32_t* vals_in = (uint32_t*)(base + L.vals_in);
  uint32_t* keys_out = (uint32_t*)(base + L.keys_out);
  if (nA == P.n && !ownerA)
    sort_keys_kernel<false><<<grid_for(P.n, 256, 8), 256, 0, st>>>(P, (const short*)pmid, disp, nullptr, nullptr,
                                                                   keys_in, vals_in);
  else
    sort_keys_kernel<true><<<grid_for(P.n, 256, 8), 256, 0, st>>>(P, (const short*)pmid, disp, (const short*)pmidB, dispB,
                                                                  keys_in, vals_in);
  PMWD_LAUNCH_CHECK();
  size_t cub_bytes = L.cub_bytes;
  // sweep layout with a power-of-two tile width: the z position inside a tile row (the low log2(bw) bits) is
  // left out of the sort -- neither the tiled deposit nor the gathers' sector counts depend on the order of
  // the ~8 particles of one 256-byte row, and 24 instead of 30 key bits is one radix pass less
  int begin_bit = 0;
  if (ty > 0 && (bw & (bw - 1)) == 0 && !getenv("PMWD_SORT_FULL")) begin_bit = ilog2_ceil(bw);
  if (begin_bit >= end_bit) begin_bit = 0;
  PMWD_CUDA_TRY(cub::DeviceRadixSort::SortPairs(base + L.cub, cub_bytes, keys_in, keys_out, vals_in,
                                                perm, P.n, begin_bit, end_bit, st));
  return PMWD_OK;
}

extern "C" int pmwd_cell_sort_perm(void* stream, const pmwd_cic_desc* d, const void* pmid,
                                   const float* disp, uint32_t* perm, void* scratch,
                                   size_t scratch_bytes, int ty, int bw) {
  PMWD_REQUIRE(d != nullptr, "null descriptor");
  return pmwd_cell_sort_perm2(stream, d, pmid, disp, d->ptcl_num, nullptr, 0, nullptr, nullptr, perm, scratch,
                              scratch_bytes, ty, bw, nullptr, nullptr, 0.f);
}

// dst[i] = (srcA ++ srcB)[perm[i]] for i < n: gather from the virtual concatenation of two sources (rows
// [0, nA) of srcA, then srcB) -- the companion of pmwd_cell_sort_perm2.
extern "C" int pmwd_permute_rows2(void* stream, int64_t n, const uint32_t* perm, int narr, const void* const* srcA,
                                  int64_t nA, const void* const* srcB, void* const* dst, const int32_t* row_bytes) {
  PMWD_REQUIRE(n >= 0 && narr >= 1 && narr <= 8 && nA >= 0, "bad sizes");
  PMWD_REQUIRE(perm && srcA && dst && row_bytes, "null buffer");
  if (n == 0) return PMWD_OK;
  RowArgs A;
  A.narr = narr;
  A.nA = nA;
  for (int a = 0; a < narr; ++a) {
    PMWD_REQUIRE((nA == 0 || srcA[a]) && dst[a] && srcA[a] != dst[a], "permute needs distinct non-null buffers");
    PMWD_REQUIRE(row_bytes[a] > 0 && row_bytes[a] % 2 == 0 && row_bytes[a] <= 64, "bad row size");
    A.src[a] = srcA[a];
    A.srcB[a] = srcB ? srcB[a] : nullptr;
    A.dst[a] = dst[a];
    A.words[a] = row_bytes[a] / 2;
  }
  cudaStream_t st = as_stream(stream);
  StageTimer timer(ST_OTHER, st);
  if (srcB && nA < ((int64_t)1 << 40))
    permute_rows_kernel<false, true><<<grid_for(n, 256, 8), 256, 0, st>>>(n, perm, A);
  else
    permute_rows_kernel<false, false><<<grid_for(n, 256, 8), 256, 0, st>>>(n, perm, A);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

extern "C" int pmwd_permute_rows(void* stream, int64_t n, const uint32_t* perm, int narr,
                                 const void* const* src, void* const* dst, const int32_t* row_bytes,
                                 int inverse) {
  PMWD_REQUIRE(n >= 0 && narr >= 1 && narr <= 8, "bad sizes");
  PMWD_REQUIRE(perm && src && dst && row_bytes, "null buffer");
  if (n == 0) return PMWD_OK;
  RowArgs A;
  A.narr = narr;
  A.nA = (int64_t)1 << 62;      // single source
  for (int a = 0; a < narr; ++a) {
    PMWD_REQUIRE(src[a] && dst[a] && src[a] != dst[a], "permute needs distinct non-null buffers");
    PMWD_REQUIRE(row_bytes[a] > 0 && row_bytes[a] % 2 == 0 && row_bytes[a] <= 64, "bad row size");
    A.src[a] = src[a];
    A.srcB[a] = nullptr;
    A.dst[a] = dst[a];
    A.words[a] = row_bytes[a] / 2;
  }
  cudaStream_t st = as_stream(stream);
  StageTimer timer(ST_OTHER, st);
  int grid = grid_for(n, 256, 8);
  if (inverse) permute_rows_kernel<true, false><<<grid, 256, 0, st>>>(n, perm, A);
  else permute_rows_kernel<false, false><<<grid, 256, 0, st>>>(n, perm, A);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}

// ---------------------------------------------------------------------------------------
// Slab-FFT transposes as ONE kernel over NVLink peer memory (no pack/unpack copies, no NCCL
// staging): every rank stores the rows of its local array straight into the peers' receive
// buffers (symmetric memory, peer-mapped pointers) at their final position.
//   mode 0 (forward):  src[mx][My][nzc]  ->  peer q = y / my :  dst_q[(rank*mx + ix)][iy][nzc]
//                      (x-slab after the local 2-D R2C  ->  y-slab [Mx][my][nzc] for the x-pass)
//   mode 1 (inverse):  src[Mx][my][nzc]  ->  peer p = x / mx :  dst_p[ix][rank*my + iy][nzc]
//                      (y-slab after the x-pass  ->  x-slab [mx][My][nzc], already unpacked for C2R)
// One warp per row of nzc complex64 (contiguous on both sides), 8-byte accesses, 4 in flight.
namespace pmwd {

struct PeerPtrs { float2* p[8]; };

__global__ void __launch_bounds__(256)
transpose_p2p_kernel(int mode, int nranks, int rank, int mx, int my, int nzc,
                     const float2* __restrict__ src, PeerPtrs dst) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int My = my * nranks, Mx = mx * nranks;
  const int64_t rows = mode == 0 ? (int64_t)mx * My : (int64_t)Mx * my;
  for (int64_t row = warp; row < rows; row += nwarp) {
    float2* d;
    if (mode == 0) {
      const int ix = (int)(row / My), y = (int)(row - (int64_t)ix * My);
      const int q = y / my, iy = y - q * my;
      d = dst.p[q] + ((int64_t)(rank * mx + ix) * my + iy) * nzc;
    } else {
      const int x = (int)(row / my), iy = (int)(row - (int64_t)x * my);
      const int p = x / mx, ix = x - p * mx;
      d = dst.p[p] + ((int64_t)ix * My + rank * my + iy) * nzc;
    }
    const float2* s = src + row * nzc;
    // 16-byte remote stores (NVLink packets) where the destination row is 16-byte aligned from
    // element `head`; loads stay 8-byte because source and destination rows are skewed
    const int head = (int)((reinterpret_cast<uintptr_t>(d) >> 3) & 1);     // 1 if d is only 8-byte aligned
    if (lane == 0 && head) d[0] = __ldcs(s);
    const int npair = (nzc - head) >> 1;
    for (int i = lane; i < npair; i += 128) {
      float2 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + 32 * u < npair) {
          a[u] = __ldcs(s + head + 2 * (i + 32 * u));
          b[u] = __ldcs(s + head + 2 * (i + 32 * u) + 1);
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + 32 * u < npair)
          *reinterpret_cast<float4*>(d + head + 2 * (i + 32 * u)) = make_float4(a[u].x, a[u].y, b[u].x, b[u].y);
    }
    const int tail = head + 2 * npair;
    if (lane == 0 && tail < nzc) d[tail] = __ldcs(s + tail);
  }
}

}  // namespace pmwd

extern "C" int pmwd_transpose_p2p(void* stream, int mode, int nranks, int rank, int mx, int my,
                                  int nzc, const void* src, const uint64_t* peer_ptrs) {
  PMWD_REQUIRE(src && peer_ptrs, "null buffer");
  PMWD_REQUIRE(nranks >= 1 && nranks <= 8 && rank >= 0 && rank < nranks, "bad rank / world size");
  PMWD_REQUIRE(mx > 0 && my > 0 && nzc > 0 && (mode == 0 || mode == 1), "bad sizes");
  pmwd::PeerPtrs P;
  for (int i = 0; i < 8; ++i) P.p[i] = i < nranks ? reinterpret_cast<float2*>(peer_ptrs[i]) : nullptr;
  for (int i = 0; i < nranks; ++i) PMWD_REQUIRE(P.p[i] != nullptr, "null peer pointer");
  cudaStream_t st = pmwd::as_stream(stream);
  pmwd::StageTimer timer(pmwd::ST_OTHER, st);
  int64_t rows = mode == 0 ? (int64_t)mx * my * nranks : (int64_t)mx * nranks * my;
  int grid = pmwd::grid_for(rows * 32, 256, 8);
  pmwd::transpose_p2p_kernel<<<grid, 256, 0, st>>>(mode, nranks, rank, mx, my, nzc, (const float2*)src, P);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}


// The same transposes on the COPY ENGINES: per peer one strided 2-D cudaMemcpy (runs of my * nzc
// complex64 = the peer's y-range of one x plane), so that no SM is busy pushing data over NVLink
// while the 2-D C2R of the previous component runs (the P2P-store kernel above doubled that C2R's
// time at N = 8: SCALE_r01 phase table).  The copies are spread over a few internal streams
// (several copy engines) and start with the next rank so that the ranks do not all target the same
// destination at once; `stream` waits for all of them.
namespace pmwd {
struct CeStreams {
  cudaStream_t s[4];
  cudaEvent_t fork, join[4];
  int device = -1;
};
static CeStreams g_ce;
static int ce_init() {
  int dev = 0;
  PMWD_CUDA_TRY(cudaGetDevice(&dev));
  if (g_ce.device == dev) return PMWD_OK;
  for (int i = 0; i < 4; ++i) {
    PMWD_CUDA_TRY(cudaStreamCreateWithFlags(&g_ce.s[i], cudaStreamNonBlocking));
    PMWD_CUDA_TRY(cudaEventCreateWithFlags(&g_ce.join[i], cudaEventDisableTiming));
  }
  PMWD_CUDA_TRY(cudaEventCreateWithFlags(&g_ce.fork, cudaEventDisableTiming));
  g_ce.device = dev;
  return PMWD_OK;
}
}  // namespace pmwd

extern "C" int pmwd_transpose_ce(void* stream, int mode, int nranks, int rank, int mx, int my, int nzc,
                                 const void* src, const uint64_t* peer_ptrs, int nstreams) {
  PMWD_REQUIRE(src && peer_ptrs, "null buffer");
  PMWD_REQUIRE(nranks >= 1 && nranks <= 8 && rank >= 0 && rank < nranks, "bad rank / world size");
  PMWD_REQUIRE(mx > 0 && my > 0 && nzc > 0 && (mode == 0 || mode == 1), "bad sizes");
  if (nstreams < 1) nstreams = 1;
  if (nstreams > 4) nstreams = 4;
  int rc = pmwd::ce_init();
  if (rc) return rc;
  cudaStream_t st = pmwd::as_stream(stream);
  pmwd::StageTimer timer(pmwd::ST_OTHER, st);
  PMWD_CUDA_TRY(cudaEventRecord(pmwd::g_ce.fork, st));
  for (int i = 0; i < nstreams; ++i) PMWD_CUDA_TRY(cudaStreamWaitEvent(pmwd::g_ce.s[i], pmwd::g_ce.fork, 0));
  const size_t run = (size_t)my * nzc * sizeof(float2);          // contiguous bytes per (x plane, peer)
  const size_t My_run = run * nranks;                            // bytes of a full x plane [My][nzc]
  const char* s8 = (const char*)src;
  for (int k = 0; k < nranks; ++k) {
    const int q = (rank + 1 + k) % nranks;                       // own block last
    char* d8 = reinterpret_cast<char*>(peer_ptrs[q]);
    PMWD_REQUIRE(d8 != nullptr, "null peer pointer");
    cudaStream_t cs = pmwd::g_ce.s[k % nstreams];
    if (mode == 0) {
      // src[ix][q*my + iy][:] -> dst_q[rank*mx + ix][iy][:]
      PMWD_CUDA_TRY(cudaMemcpy2DAsync(d8 + (size_t)rank * mx * run, run, s8 + (size_t)q * run, My_run, run,
                                      (size_t)mx, cudaMemcpyDeviceToDevice, cs));
    } else {
      // src[q*mx + ix][iy][:] -> dst_q[ix][rank*my + iy][:]
      PMWD_CUDA_TRY(cudaMemcpy2DAsync(d8 + (size_t)rank * run, My_run, s8 + (size_t)q * mx * run, run, run,
                                      (size_t)mx, cudaMemcpyDeviceToDevice, cs));
    }
  }
  for (int i = 0; i < nstreams; ++i) {
    PMWD_CUDA_TRY(cudaEventRecord(pmwd::g_ce.join[i], pmwd::g_ce.s[i]));
    PMWD_CUDA_TRY(cudaStreamWaitEvent(st, pmwd::g_ce.join[i], 0));
  }
  return PMWD_OK;
}


// Generic form used by the chunk-pipelined slab FFT (pmwd_b200/dist.py): to every rank q one strided
// 2-D copy of `height` rows of `width` bytes,
//   src + q * src_peer_stride (row pitch spitch)  ->  peer_ptrs[q] + dst_off (row pitch dpitch),
// on the copy engines, spread over `nstreams` internal streams forked from / joined into `stream`.
extern "C" int pmwd_peer_copy2d(void* stream, int nranks, int rank, size_t width, size_t height, const void* src,
                                size_t src_peer_stride, size_t spitch, const uint64_t* peer_ptrs, size_t dst_off,
                                size_t dpitch, int nstreams) {
  PMWD_REQUIRE(src && peer_ptrs, "null buffer");
  PMWD_REQUIRE(nranks >= 1 && nranks <= 8 && rank >= 0 && rank < nranks, "bad rank / world size");
  PMWD_REQUIRE(width > 0 && height > 0 && spitch >= width && dpitch >= width, "bad sizes");
  if (nstreams < 1) nstreams = 1;
  if (nstreams > 4) nstreams = 4;
  int rc = pmwd::ce_init();
  if (rc) return rc;
  cudaStream_t st = pmwd::as_stream(stream);
  pmwd::StageTimer timer(pmwd::ST_OTHER, st);
  PMWD_CUDA_TRY(cudaEventRecord(pmwd::g_ce.fork, st));
  for (int i = 0; i < nstreams; ++i) PMWD_CUDA_TRY(cudaStreamWaitEvent(pmwd::g_ce.s[i], pmwd::g_ce.fork, 0));
  // optionally split every peer's rows over several streams (= copy engines); measured at N = 2
  // (profiles/r02_slab_ab_n2.txt): no gain, one engine already reaches the ~600 GB/s the pair sustains
  static const int split_env = [] { const char* e = getenv("PMWD_P2P_CE_SPLIT"); return e ? atoi(e) : 0; }();
  int nsplit = split_env > 0 ? split_env : 1;
  if ((size_t)nsplit > height) nsplit = (int)height;
  if (nsplit < 1) nsplit = 1;
  int job = 0;
  for (int k = 0; k < nranks; ++k) {
    const int q = (rank + 1 + k) % nranks;                       // own block last
    char* d8 = reinterpret_cast<char*>(peer_ptrs[q]);
    PMWD_REQUIRE(d8 != nullptr, "null peer pointer");
    for (int part = 0; part < nsplit; ++part, ++job) {
      const size_t r0 = height * part / nsplit, r1 = height * (part + 1) / nsplit;
      if (r1 == r0) continue;
      PMWD_CUDA_TRY(cudaMemcpy2DAsync(d8 + dst_off + r0 * dpitch, dpitch,
                                      (const char*)src + (size_t)q * src_peer_stride + r0 * spitch, spitch, width,
                                      r1 - r0, cudaMemcpyDeviceToDevice, pmwd::g_ce.s[job % nstreams]));
    }
  }
  for (int i = 0; i < nstreams; ++i) {
    PMWD_CUDA_TRY(cudaEventRecord(pmwd::g_ce.join[i], pmwd::g_ce.s[i]));
    PMWD_CUDA_TRY(cudaStreamWaitEvent(st, pmwd::g_ce.join[i], 0));
  }
  return PMWD_OK;
}


// ---------------------------------------------------------------------------------------------
// Slab bookkeeping in one pass over the particles: the rank that owns each particle's base plane
// (Eulerian ownership, pmwd_b200/migrate.py) and how many halo planes the slab [x0, x0 + mx) needs for
// the particles it holds (max over particles, atomicMax into *need).  Same float32 cell arithmetic as
// the CIC kernels (pm_util.py:129-136).
namespace pmwd {
__global__ void __launch_bounds__(256)
slab_owner_kernel(int64_t n, const short* __restrict__ pmid, const float* __restrict__ disp, float cell, int Mx,
                  int planes_per_rank, int x0, int mx, uint8_t* __restrict__ owner, int* __restrict__ need) {
  int local = 0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const float t = __fdiv_rn(disp[3 * p], cell);
    const int plane = wrap_index((int)pmid[3 * p] + (int)floorf(t), Mx);
    if (owner) owner[p] = (uint8_t)(plane / planes_per_rank);
    int d = plane - x0;
    if (d < 0) d += Mx;
    const int right = d + 2 - mx;
    const int nd = d < mx ? (right > 0 ? right : 0) : (right < Mx - d ? right : Mx - d);
    local = nd > local ? nd : local;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const int t = __shfl_xor_sync(0xffffffffu, local, o);
    local = t > local ? t : local;
  }
  if ((threadIdx.x & 31) == 0 && local > 0) atomicMax(need, local);
}
}  // namespace pmwd

extern "C" int pmwd_slab_owner(void* stream, int64_t n, const void* pmid, const float* disp, double cell_size,
                               int Mx, int nranks, int x0, int mx, uint8_t* owner, int32_t* need) {
  PMWD_REQUIRE(need && (n == 0 || (pmid && disp)), "null buffer");
  PMWD_REQUIRE(Mx > 0 && nranks >= 1 && nranks <= 255 && Mx % nranks == 0 && mx > 0, "bad slab geometry");
  cudaStream_t st = pmwd::as_stream(stream);
  PMWD_CUDA_TRY(cudaMemsetAsync(need, 0, sizeof(int32_t), st));
  if (n == 0) return PMWD_OK;
  pmwd::slab_owner_kernel<<<pmwd::grid_for(n, 256, 8), 256, 0, st>>>(n, (const short*)pmid, disp, (float)cell_size, Mx,
                                                                    Mx / nranks, x0, mx, owner, need);
  PMWD_LAUNCH_CHECK();
  return PMWD_OK;
}
