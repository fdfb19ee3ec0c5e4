// Lab: throughput of the accumulation primitives a tiled CIC scatter could use, under the access
// pattern of CIC (8 neighbours per particle, particles of a warp mostly in neighbouring cells):
//   shared float atomicAdd (ATOMS.CAST.SPIN loop), shared int32 / uint64 atomicAdd (native ATOMS.ADD),
//   global float RED.ADD.F32, global RED.ADD.F32x2.
// Prints G updates/s per variant for a dense (clustered) and a sparse (early-time) particle layout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lab/smem_atomics tools/lab/smem_atomics.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int TX = 8, TY = 8, TZ = 64;                 // shared tile (cells), + 1 halo on the upper sides
constexpr int SX = TX + 1, SY = TY + 1, SZ = TZ + 1;
constexpr int TILE = SX * SY * SZ;

__device__ __forceinline__ uint32_t hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// mode 0: shared float, 1: shared int32 fixed point, 2: shared uint64 fixed point, 3: global float RED,
// 4: global float2 RED on even z.  ppc = particles per cell (density of the layout).
template <int MODE>
__global__ void __launch_bounds__(256) deposit(float* gmesh, int ntile_iters, float ppc, unsigned long long* sink) {
  __shared__ unsigned long long sm64[MODE == 2 ? TILE : 1];
  __shared__ float smf[MODE == 0 ? TILE : 1];
  __shared__ int smi[MODE == 1 ? TILE : 1];
  for (int i = threadIdx.x; i < TILE; i += blockDim.x) {
    if (MODE == 0) smf[i] = 0.f;
    if (MODE == 1) smi[i] = 0;
    if (MODE == 2) sm64[i] = 0ull;
  }
  __syncthreads();
  const int npart = (int)(TX * TY * TZ * ppc);
  float* gt = gmesh + (size_t)blockIdx.x * TILE;       // this CTA's patch of the global mesh
  for (int it = 0; it < ntile_iters; ++it) {
    for (int p = threadIdx.x; p < npart; p += blockDim.x) {
      // particles ordered along z inside 2x2 columns, like the product's storage
      const int col = p / (npart / 16 + 1), inz = p % (npart / 16 + 1);
      const uint32_t h = hash(p * 2654435761u + it);
      const int x = (col & 3) * 2 + (h & 1), y = (col >> 2) * 2 + ((h >> 1) & 1);
      const int z = (int)((long long)inz * TZ / (npart / 16 + 1));
      const float fx = (h >> 8 & 255) / 256.f, fy = (h >> 16 & 255) / 256.f, fz = (h >> 24) / 256.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int bx = n & 1, by = (n >> 1) & 1, bz = n >> 2;
        const float w = (bx ? fx : 1 - fx) * (by ? fy : 1 - fy) * (bz ? fz : 1 - fz) * 8.f;
        const int idx = ((x + bx) * SY + (y + by)) * SZ + z + bz;
        if (MODE == 0) atomicAdd(&smf[idx], w);
        if (MODE == 1) atomicAdd(&smi[idx], (int)(w * 65536.f));
        if (MODE == 2) atomicAdd(&sm64[idx], (unsigned long long)(w * 4294967296.f));
        if (MODE == 3) atomicAdd(gt + idx, w);
        if (MODE == 4) {
          if (bz == 0) {
            const float w1 = (bx ? fx : 1 - fx) * (by ? fy : 1 - fy) * fz * 8.f;
            if ((idx & 1) == 0) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gt + idx), "f"(w), "f"(w1) : "memory");
            else { atomicAdd(gt + idx, w); atomicAdd(gt + idx + 1, w1); }
          }
        }
      }
    }
    __syncthreads();
  }
  unsigned long long acc = 0;
  for (int i = threadIdx.x; i < TILE; i += blockDim.x) {
    if (MODE == 0) acc += (unsigned long long)smf[i];
    if (MODE == 1) acc += smi[i];
    if (MODE == 2) acc += sm64[i] >> 32;
  }
  if (acc == 0xdeadbeefdeadbeefull) *sink = acc;
}

template <int MODE>
static void run(const char* name, float ppc, float* gmesh, unsigned long long* sink) {
  const int grid = 148 * 4, iters = 64;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  deposit<MODE><<<grid, 256>>>(gmesh, 2, ppc, sink);
  CK(cudaEventRecord(a));
  deposit<MODE><<<grid, 256>>>(gmesh, iters, ppc, sink);
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  const double upd = (double)grid * iters * (int)(TX * TY * TZ * ppc) * 8;
  printf("%-34s ppc %5.3f  %8.3f ms  %7.2f G updates/s\n", name, ppc, ms, upd / ms * 1e-6);
}

int main() {
  float* gmesh; unsigned long long* sink;
  CK(cudaMalloc(&gmesh, (size_t)148 * 4 * TILE * sizeof(float)));
  CK(cudaMemset(gmesh, 0, (size_t)148 * 4 * TILE * sizeof(float)));
  CK(cudaMalloc(&sink, 8));
  for (float ppc : {0.125f, 1.0f, 8.0f}) {
    run<0>("shared float atomicAdd", ppc, gmesh, sink);
    run<1>("shared int32 fixed point", ppc, gmesh, sink);
    run<2>("shared uint64 fixed point", ppc, gmesh, sink);
    run<3>("global float RED (L2 resident)", ppc, gmesh, sink);
    run<4>("global float2 RED on even z", ppc, gmesh, sink);
  }
  // reference point: the product's scatter does 134e6 particles x 8 updates in 5.55 ms = 193 G updates/s
  return 0;
}
