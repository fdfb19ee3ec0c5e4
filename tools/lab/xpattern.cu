// Lab: memory-access-pattern ceiling of the fused x-pass (no FFT, no algebra).
// Each CTA walks tiles of T adjacent kz columns for one y; thread (j, c) moves the points
// x = j + J e, e = 0..15 of column c: reads one array, writes NOUT arrays, all with the x stride
// of a [nx][ny][nzc] complex64 mesh (4.2 MB at 1024^3).  Prints GB/s per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lab/xpattern tools/lab/xpattern.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <typename V, int ROWS_PER_THREAD, int NOUT, bool BLOCKED>
__global__ void __launch_bounds__(1024) pattern(const V* __restrict__ in, V* o0, V* o1, V* o2, int nx, int ny, int rowv /* V per (x,y) row */,
                                               int tv /* V per tile row */) {
  const int lanes = tv;                       // threads along a tile row
  const int c = threadIdx.x % lanes, j = threadIdx.x / lanes;
  const int J = blockDim.x / lanes;           // rows covered per pass
  const int ztiles = rowv / tv;               // (ignore ragged tail)
  const int64_t ntiles = (int64_t)ny * ztiles;
  const int64_t plane = (int64_t)ny * rowv;
  const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
  const int64_t t0 = BLOCKED ? blockIdx.x * per : blockIdx.x;
  const int64_t t1 = BLOCKED ? (t0 + per < ntiles ? t0 + per : ntiles) : ntiles;
  const int64_t ts = BLOCKED ? 1 : gridDim.x;
  for (int64_t tile = t0; tile < t1; tile += ts) {
    const int iy = (int)(tile / ztiles);
    const int z0 = (int)(tile - (int64_t)iy * ztiles) * tv;
    const int64_t g0 = (int64_t)j * plane + (int64_t)iy * rowv + z0 + c;
    V v[ROWS_PER_THREAD];
#pragma unroll
    for (int e = 0; e < ROWS_PER_THREAD; ++e) v[e] = __ldcs(in + g0 + (int64_t)(J * e) * plane);
#pragma unroll
    for (int e = 0; e < ROWS_PER_THREAD; ++e) {
      const int64_t g = g0 + (int64_t)(J * e) * plane;
      __stcs(o0 + g, v[e]);
      if (NOUT > 1) __stcs(o1 + g, v[e]);
      if (NOUT > 2) __stcs(o2 + g, v[e]);
    }
  }
}

// y-pass pattern: for a fixed x plane, a tile is `tv` V-words of one kz run times all ny rows (row stride
// rowv V-words, 4 KB); in place (read + write the same array), like a strided C2C pass of a 2-D FFT.
template <typename V, int ROWS_PER_THREAD>
__global__ void __launch_bounds__(1024) ypattern(V* data, int nx, int ny, int rowv, int tv, int second /* offset of a 2nd word, 0 = none */) {
  const int lanes = tv;
  const int c = threadIdx.x % lanes, j = threadIdx.x / lanes;
  const int J = blockDim.x / lanes;
  const int ztiles = rowv / (second ? 2 * tv : tv);
  const int64_t ntiles = (int64_t)nx * ztiles;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int ix = (int)(tile / ztiles);
    const int z0 = (int)(tile - (int64_t)ix * ztiles) * (second ? 2 * tv : tv);
    const int64_t g0 = ((int64_t)ix * ny + j) * rowv + z0 + c;
    V v[ROWS_PER_THREAD], w[ROWS_PER_THREAD];
#pragma unroll
    for (int e = 0; e < ROWS_PER_THREAD; ++e) {
      v[e] = __ldcs(data + g0 + (int64_t)(J * e) * rowv);
      if (second) w[e] = __ldcs(data + g0 + (int64_t)(J * e) * rowv + second);
    }
#pragma unroll
    for (int e = 0; e < ROWS_PER_THREAD; ++e) {
      __stcs(data + g0 + (int64_t)(J * e) * rowv, v[e]);
      if (second) __stcs(data + g0 + (int64_t)(J * e) * rowv + second, w[e]);
    }
  }
}

template <typename V, int RPT>
static void runy(const char* name, int tile_bytes, int second_bytes, int nx, int ny, int nzc_row, void* data) {
  const int rowv = nzc_row * 8 / sizeof(V);
  const int tv = tile_bytes / sizeof(V);
  const int threads = ny / RPT * tv;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e9;
  for (int it = 0; it < 4; ++it) {
    CK(cudaEventRecord(a));
    ypattern<V, RPT><<<148, threads>>>((V*)data, nx, ny, rowv, tv, second_bytes / (int)sizeof(V));
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (it > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const int cols = (second_bytes ? 2 : 1) * tile_bytes / 8;
  const double bytes = 2.0 * nx * ny * (double)(nzc_row / cols * cols) * 8;
  printf("%-52s threads %4d  %7.3f ms  %7.1f GB/s\n", name, threads, best, bytes / best * 1e-6);
}

template <typename V, int RPT, int NOUT, bool BLOCKED>
static void run(const char* name, int tile_bytes, int nx, int ny, int nzc_pad, void* in, void* o0, void* o1, void* o2, int ctas_per_sm,
                bool inplace = false) {
  const int rowv = nzc_pad * 8 / sizeof(V);
  const int tv = tile_bytes / sizeof(V);
  const int threads = nx / RPT * tv;
  if (threads > 512 && threads != 1024) { printf("%s: bad threads %d\n", name, threads); return; }
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const int grid = 148 * ctas_per_sm;
  float best = 1e9;
  for (int it = 0; it < 4; ++it) {
    CK(cudaEventRecord(a));
    pattern<V, RPT, NOUT, BLOCKED><<<grid, threads>>>((const V*)in, (V*)(inplace ? in : o0), (V*)o1, (V*)o2, nx, ny, rowv, tv);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (it > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double bytes = (double)nx * ny * nzc_pad * 8 * (1 + NOUT);
  printf("%-44s threads %4d x %d CTA/SM  %7.3f ms  %7.1f GB/s\n", name, threads, ctas_per_sm, best, bytes / best * 1e-6);
}

int main() {
  const int nx = 1024, ny = 1024, nzc = 512;   // 512 (not 513): keeps every tile full; pattern is what matters
  const size_t bytes = (size_t)nx * ny * nzc * 8;
  const size_t bytes_pad = (size_t)nx * ny * 528 * 8;   // the padded-row y-pass variants
  void *in, *o0, *o1, *o2;
  CK(cudaMalloc(&in, bytes_pad)); CK(cudaMalloc(&o0, bytes)); CK(cudaMalloc(&o1, bytes)); CK(cudaMalloc(&o2, bytes));
  CK(cudaMemset(in, 1, bytes));
  // reference: plain streaming copy
  {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaMemcpy(o0, in, bytes, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(a)); CK(cudaMemcpyAsync(o0, in, bytes, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    printf("%-44s %7.3f ms  %7.1f GB/s\n", "cudaMemcpy D2D 4.3 GB", ms, 2.0 * bytes / ms * 1e-6);
  }
  run<float2, 16, 3, false>("64B rows, 16 rows/thr, 3 out, strided tiles", 64, nx, ny, nzc, in, o0, o1, o2, 1);
  run<float2, 16, 3, false>("same, 2 CTA/SM", 64, nx, ny, nzc, in, o0, o1, o2, 2);
  run<float2, 16, 3, false>("same, 4 CTA/SM", 64, nx, ny, nzc, in, o0, o1, o2, 4);
  run<float2, 16, 1, false>("64B rows, 1 out", 64, nx, ny, nzc, in, o0, o1, o2, 1);
  run<float2, 16, 1, false>("64B rows, 1 out, in place", 64, nx, ny, nzc, in, o0, o1, o2, 1, true);
  run<float2, 16, 1, false>("64B rows, 1 out, 4 CTA/SM", 64, nx, ny, nzc, in, o0, o1, o2, 4);
  run<float2, 16, 3, true>("64B rows, 3 out, blocked tiles", 64, nx, ny, nzc, in, o0, o1, o2, 1);
  run<float2, 32, 3, false>("128B rows, 32 rows/thr, 3 out", 128, nx, ny, nzc, in, o0, o1, o2, 1);
  run<float2, 32, 3, false>("128B rows, 32 rows/thr, 3 out, 2 CTA/SM", 128, nx, ny, nzc, in, o0, o1, o2, 2);
  run<float4, 16, 3, false>("128B rows as float4 (8 lanes), 3 out", 128, nx, ny, nzc, in, o0, o1, o2, 1);
  run<float4, 16, 3, false>("128B rows as float4, 3 out, 2 CTA/SM", 128, nx, ny, nzc, in, o0, o1, o2, 2);
  run<float4, 16, 3, false>("64B rows as float4 (4 lanes), 3 out", 64, nx, ny, nzc, in, o0, o1, o2, 1);
  run<float4, 16, 3, false>("64B rows as float4 (4 lanes), 3 out, 4 CTA/SM", 64, nx, ny, nzc, in, o0, o1, o2, 4);
  run<float4, 16, 3, false>("256B rows as float4 (16 lanes), 3 out", 256, nx, ny, nzc, in, o0, o1, o2, 1);
  run<float4, 16, 1, false>("128B rows as float4, 1 out", 128, nx, ny, nzc, in, o0, o1, o2, 1);
  // ---- y-pass (row stride 4 KB within a plane), in place; cuFFT's strided pass does 8.6 GB in 2.19 ms = 3.9 TB/s
  printf("y-pass patterns (in place, 1 read + 1 write):\n");
  runy<float2, 16>("row 512 words, 64B runs (8 x float2)", 64, 0, nx, ny, 512, in);
  runy<float2, 16>("row 512 words, 128B runs as float2 c and c+8", 64, 64, nx, ny, 512, in);
  runy<float4, 16>("row 512 words, 128B runs as float4 (aligned)", 128, 0, nx, ny, 512, in);
  runy<float2, 16>("row 513 words (as in the product), 64B runs", 64, 0, nx, ny, 513, in);
  runy<float2, 16>("row 513 words, 128B runs as float2 c and c+8", 64, 64, nx, ny, 513, in);
  runy<float4, 16>("row 528 words (padded), 128B runs as float4", 128, 0, nx, ny, 528, in);
  return 0;
}
