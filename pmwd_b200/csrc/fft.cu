// Context + cuFFT plan cache: rfftn / irfftn of pmwd/pm_util.py:236-344 (fftfwd / fftinv).
//
// Plans are created once per (rank, shape) by pmwd_ctx_reserve -- the only place that
// allocates -- and share one work area sized for the largest plan; execution only enqueues
// on the caller's stream (cufftSetStream per call).  The 1/N of numpy's 'backward' norm is a
// separate fused scale pass here only when a caller asks for it outside the force pipeline
// (the force pipeline folds it into the k-space kernel's `scale`).
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace pmwd {

struct PlanPair {
  cufftHandle r2c = 0, c2r = 0;
  size_t work = 0;
  int chunk = 0;     // 2-D plans only: planes per cuFFT call (0 = the whole batch in one call)
  int pad = 0;       // 2-D plans only: row length of the half-spectrum (>= nz/2+1)
};

}  // namespace pmwd

struct pmwd_ctx {
  int device = 0;
  std::mutex mu;
  std::map<std::tuple<int, int, int, int>, pmwd::PlanPair> plans;
  std::map<std::tuple<int, long long, long long>, cufftHandle> xplans;   // strided 1-D C2C
  void* work = nullptr;
  size_t work_bytes = 0;
  int fft2d_chunk = -1;    // planes per call of the (y, z) transforms; -1 = PMWD_FFT2D_CHUNK or the default
  int fft2d_pad = 0;       // lab: row length (complex elements) of the half-spectrum of the 2-D plans; 0 = nz/2+1
};

namespace pmwd {

static std::tuple<int, int, int, int> key_of(int rank, const int32_t* shape) {
  return std::make_tuple(rank, shape[0], rank > 1 ? shape[1] : 1, rank > 2 ? shape[2] : 1);
}

__global__ void __launch_bounds__(256) scale_kernel(float* x, int64_t n, float s) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // float4 body, scalar tail
  int64_t n4 = n >> 2;
  float4* x4 = reinterpret_cast<float4*>(x);
  for (int64_t j = i; j < n4; j += stride) {
    float4 v = x4[j];
    v.x = __fmul_rn(v.x, s); v.y = __fmul_rn(v.y, s); v.z = __fmul_rn(v.z, s); v.w = __fmul_rn(v.w, s);
    x4[j] = v;
  }
  for (int64_t j = (n4 << 2) + i; j < n; j += stride) x[j] = __fmul_rn(x[j], s);
}

int find_plans(pmwd_ctx* ctx, int rank, const int32_t* shape, PlanPair* out) {
  PMWD_REQUIRE(ctx != nullptr, "null context");
  PMWD_REQUIRE(rank >= 1 && rank <= 3 && shape, "bad rank/shape");
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->plans.find(key_of(rank, shape));
  if (it == ctx->plans.end()) {
    set_error("no cuFFT plan for this shape: call pmwd_ctx_reserve first");
    return PMWD_ESTATE;
  }
  *out = it->second;
  return PMWD_OK;
}

int fft_r2c(pmwd_ctx* ctx, cudaStream_t st, int rank, const int32_t* shape, const float* in,
            void* out) {
  PlanPair pp;
  int rc = find_plans(ctx, rank, shape, &pp);
  if (rc) return rc;
  PMWD_CUFFT_TRY(cufftSetStream(pp.r2c, st));
  PMWD_CUFFT_TRY(cufftExecR2C(pp.r2c, const_cast<float*>(in), (cufftComplex*)out));
  return PMWD_OK;
}

int fft_c2r(pmwd_ctx* ctx, cudaStream_t st, int rank, const int32_t* shape, void* in, float* out) {
  PlanPair pp;
  int rc = find_plans(ctx, rank, shape, &pp);
  if (rc) return rc;
  PMWD_CUFFT_TRY(cufftSetStream(pp.c2r, st));
  PMWD_CUFFT_TRY(cufftExecC2R(pp.c2r, (cufftComplex*)in, out));
  return PMWD_OK;
}

static int find_plans2d(pmwd_ctx* ctx, const int32_t* shape, PlanPair* out) {
  PMWD_REQUIRE(ctx != nullptr && shape, "null context/shape");
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->plans.find(std::make_tuple(23, shape[0], shape[1], shape[2]));
  if (it == ctx->plans.end()) {
    set_error("no 2-D cuFFT plans for this shape: call pmwd_ctx_reserve first");
    return PMWD_ESTATE;
  }
  *out = it->second;
  return PMWD_OK;
}

// in[nx][ny][nz] real -> out[nx][ny][nz/2+1]: R2C over (y, z) for every x plane.
// cuFFT runs a 2-D transform as two passes over the whole batch, i.e. the intermediate array (8.6 GB at
// 1024^3) goes out to DRAM and comes back.  The plans can be issued `chunk` planes at a time so that the second
// pass could find the first one's output in the 126 MB L2 -- measured: no gain (profiles/r02_fft2d_chunks.txt),
// so chunk = 0 (one call) unless pmwd_ctx_set_fft2d_chunk / PMWD_FFT2D_CHUNK says otherwise.
int fft2d_r2c(pmwd_ctx* ctx, cudaStream_t st, const int32_t* shape, const float* in, void* out) {
  PlanPair pp;
  int rc = find_plans2d(ctx, shape, &pp);
  if (rc) return rc;
  PMWD_CUFFT_TRY(cufftSetStream(pp.r2c, st));
  const int64_t pin = (int64_t)shape[1] * shape[2], pout = (int64_t)shape[1] * (pp.pad ? pp.pad : shape[2] / 2 + 1);
  const int step = pp.chunk > 0 ? pp.chunk : shape[0];
  for (int x = 0; x < shape[0]; x += step)
    PMWD_CUFFT_TRY(cufftExecR2C(pp.r2c, const_cast<float*>(in) + x * pin, (cufftComplex*)out + x * pout));
  return PMWD_OK;
}

int fft2d_c2r(pmwd_ctx* ctx, cudaStream_t st, const int32_t* shape, void* in, float* out) {
  PlanPair pp;
  int rc = find_plans2d(ctx, shape, &pp);
  if (rc) return rc;
  PMWD_CUFFT_TRY(cufftSetStream(pp.c2r, st));
  const int64_t pin = (int64_t)shape[1] * (pp.pad ? pp.pad : shape[2] / 2 + 1), pout = (int64_t)shape[1] * shape[2];
  const int step = pp.chunk > 0 ? pp.chunk : shape[0];
  for (int x = 0; x < shape[0]; x += step)
    PMWD_CUFFT_TRY(cufftExecC2R(pp.c2r, (cufftComplex*)in + x * pin, out + x * pout));
  return PMWD_OK;
}

}  // namespace pmwd

using namespace pmwd;

extern "C" int pmwd_ctx_create(pmwd_ctx** out, int device) {
  PMWD_REQUIRE(out != nullptr, "null ctx pointer");
  int ndev = 0;
  PMWD_CUDA_TRY(cudaGetDeviceCount(&ndev));
  PMWD_REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
  pmwd_ctx* c = new pmwd_ctx();
  c->device = device;
  *out = c;
  return PMWD_OK;
}

extern "C" int pmwd_ctx_destroy(pmwd_ctx* ctx) {
  if (!ctx) return PMWD_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(ctx->device);
  for (auto& kv : ctx->plans) {
    if (kv.second.r2c) cufftDestroy(kv.second.r2c);
    if (kv.second.c2r) cufftDestroy(kv.second.c2r);
  }
  for (auto& kv : ctx->xplans) cufftDestroy(kv.second);
  if (ctx->work) cudaFree(ctx->work);
  cudaSetDevice(prev);
  delete ctx;
  return PMWD_OK;
}

// Planes per cuFFT call of the (y, z) transforms for plans reserved AFTER this call: 0 = whole batch in one
// call, -1 = PMWD_FFT2D_CHUNK or the L2-sized default.
extern "C" int pmwd_ctx_set_fft2d_chunk(pmwd_ctx* ctx, int planes) {
  PMWD_REQUIRE(ctx != nullptr, "null context");
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->fft2d_chunk = planes;
  return PMWD_OK;
}

// Lab only (tools/lab/fft2d_pad.py): row length, in complex elements, of the half-spectrum the 2-D plans reserved
// AFTER this call read / write (>= nz/2+1; 0 = unpadded).  The force pipeline does not use padded spectra.
extern "C" int pmwd_ctx_set_fft2d_pad(pmwd_ctx* ctx, int row_elems) {
  PMWD_REQUIRE(ctx != nullptr && row_elems >= 0, "bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->fft2d_pad = row_elems;
  return PMWD_OK;
}

extern "C" int pmwd_ctx_reserve(pmwd_ctx* ctx, int rank, const int32_t* shape) {
  PMWD_REQUIRE(ctx != nullptr, "null context");
  PMWD_REQUIRE(rank >= 1 && rank <= 3 && shape, "bad rank/shape");
  for (int a = 0; a < rank; ++a) PMWD_REQUIRE(shape[a] > 0, "non-positive shape");
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto key = key_of(rank, shape);
  if (ctx->plans.count(key)) return PMWD_OK;
  int prev = 0;
  PMWD_CUDA_TRY(cudaGetDevice(&prev));
  PMWD_CUDA_TRY(cudaSetDevice(ctx->device));
  long long n[3];
  for (int a = 0; a < rank; ++a) n[a] = shape[a];
  PlanPair pp;
  size_t w1 = 0, w2 = 0;
  PMWD_CUFFT_TRY(cufftCreate(&pp.r2c));
  PMWD_CUFFT_TRY(cufftSetAutoAllocation(pp.r2c, 0));
  PMWD_CUFFT_TRY(cufftMakePlanMany64(pp.r2c, rank, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, 1, &w1));
  PMWD_CUFFT_TRY(cufftCreate(&pp.c2r));
  PMWD_CUFFT_TRY(cufftSetAutoAllocation(pp.c2r, 0));
  PMWD_CUFFT_TRY(cufftMakePlanMany64(pp.c2r, rank, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, 1, &w2));
  pp.work = w1 > w2 ? w1 : w2;
  if (pp.work > ctx->work_bytes) {
    // grow the shared work area and re-point every existing plan at it
    PMWD_CUDA_TRY(cudaDeviceSynchronize());
    if (ctx->work) PMWD_CUDA_TRY(cudaFree(ctx->work));
    ctx->work = nullptr;
    PMWD_CUDA_TRY(cudaMalloc(&ctx->work, pp.work));
    ctx->work_bytes = pp.work;
    for (auto& kv : ctx->plans) {
      PMWD_CUFFT_TRY(cufftSetWorkArea(kv.second.r2c, ctx->work));
      PMWD_CUFFT_TRY(cufftSetWorkArea(kv.second.c2r, ctx->work));
    }
  }
  if (ctx->work) {
    PMWD_CUFFT_TRY(cufftSetWorkArea(pp.r2c, ctx->work));
    PMWD_CUFFT_TRY(cufftSetWorkArea(pp.c2r, ctx->work));
  }
  ctx->plans[key] = pp;
  if (rank == 3) {
    // (y, z) 2-D transforms batched over the x planes: the companions of the fused x-pass
    // (xpass.cu).  Key: rank tag 23.
    PlanPair p2;
    long long n2[2] = {shape[1], shape[2]};
    long long nzc = shape[2] / 2 + 1;
    size_t v1 = 0, v2 = 0;
    // planes per call: the largest divisor of the batch that keeps one chunk's intermediate array
    // (chunk * ny * nzc complex64) within the L2-resident budget
    int want = ctx->fft2d_chunk;
    if (want < 0) {
      const char* e = getenv("PMWD_FFT2D_CHUNK");
      want = e ? atoi(e) : -1;
    }
    // measured at 1024^3 (profiles/r02_fft2d_chunks.txt): every chunk size is SLOWER than one call for the
    // whole batch (3.74 ms; 64 planes per call 3.95, 8 planes 4.96): cuFFT's small-batch launches do not fill
    // the GPU and nothing is gained from L2.  Default = one call; the knob stays for other shapes / parts.
    if (want < 0) want = 0;
    int chunk = 0;
    if (want > 0 && want < shape[0])
      for (int c = want; c >= 1; --c) if (shape[0] % c == 0) { chunk = c; break; }
    p2.chunk = chunk;
    const long long batch = chunk > 0 ? chunk : shape[0];
    long long rembed[2] = {shape[1], shape[2]};
    long long cembed[2] = {shape[1], ctx->fft2d_pad >= nzc ? ctx->fft2d_pad : nzc};
    PMWD_CUFFT_TRY(cufftCreate(&p2.r2c));
    PMWD_CUFFT_TRY(cufftSetAutoAllocation(p2.r2c, 0));
    PMWD_CUFFT_TRY(cufftMakePlanMany64(p2.r2c, 2, n2, rembed, 1, rembed[0] * rembed[1], cembed, 1,
                                       cembed[0] * cembed[1], CUFFT_R2C, batch, &v1));
    PMWD_CUFFT_TRY(cufftCreate(&p2.c2r));
    PMWD_CUFFT_TRY(cufftSetAutoAllocation(p2.c2r, 0));
    PMWD_CUFFT_TRY(cufftMakePlanMany64(p2.c2r, 2, n2, cembed, 1, cembed[0] * cembed[1], rembed, 1,
                                       rembed[0] * rembed[1], CUFFT_C2R, batch, &v2));
    p2.pad = (int)cembed[1];
    p2.work = v1 > v2 ? v1 : v2;
    if (p2.work > ctx->work_bytes) {
      PMWD_CUDA_TRY(cudaDeviceSynchronize());
      if (ctx->work) PMWD_CUDA_TRY(cudaFree(ctx->work));
      ctx->work = nullptr;
      PMWD_CUDA_TRY(cudaMalloc(&ctx->work, p2.work));
      ctx->work_bytes = p2.work;
      for (auto& kv : ctx->plans) {
        PMWD_CUFFT_TRY(cufftSetWorkArea(kv.second.r2c, ctx->work));
        PMWD_CUFFT_TRY(cufftSetWorkArea(kv.second.c2r, ctx->work));
      }
    }
    if (ctx->work) {
      PMWD_CUFFT_TRY(cufftSetWorkArea(p2.r2c, ctx->work));
      PMWD_CUFFT_TRY(cufftSetWorkArea(p2.c2r, ctx->work));
    }
    ctx->plans[std::make_tuple(23, shape[0], shape[1], shape[2])] = p2;
  }
  PMWD_CUDA_TRY(cudaSetDevice(prev));
  return PMWD_OK;
}

extern "C" int pmwd_fft_r2c(pmwd_ctx* ctx, void* stream, int rank, const int32_t* shape,
                            const float* in, void* out) {
  PMWD_REQUIRE(in && out, "null buffer");
  return fft_r2c(ctx, as_stream(stream), rank, shape, in, out);
}

extern "C" int pmwd_fft_c2r(pmwd_ctx* ctx, void* stream, int rank, const int32_t* shape,
                            void* in, float* out, float scale) {
  PMWD_REQUIRE(in && out, "null buffer");
  int rc = fft_c2r(ctx, as_stream(stream), rank, shape, in, out);
  if (rc) return rc;
  if (scale != 1.f) {
    int64_t n = 1;
    for (int a = 0; a < rank; ++a) n *= shape[a];
    int grid = grid_for((n + 3) / 4, 256, 8);
    scale_kernel<<<grid, 256, 0, as_stream(stream)>>>(out, n, scale);
    PMWD_LAUNCH_CHECK();
  }
  return PMWD_OK;
}

// 1-D complex transforms along the LEADING axis of a C-order array data[n][inner] (stride =
// inner, `inner` transforms), in place: the x-pass of the slab-decomposed FFT after the
// all-to-all (dist.py).  The plan (with its own work area) is created on first use for a given
// (n, inner) -- that first call allocates; later calls only enqueue.
extern "C" int pmwd_fft_c2c_lead(pmwd_ctx* ctx, void* stream, int n, long long inner, void* data,
                                 int inverse) {
  PMWD_REQUIRE(ctx && data, "null buffer");
  PMWD_REQUIRE(n > 0 && inner > 0, "bad sizes");
  cufftHandle plan;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto key = std::make_tuple(n, inner, 0LL);
    auto it = ctx->xplans.find(key);
    if (it == ctx->xplans.end()) {
      int prev = 0;
      PMWD_CUDA_TRY(cudaGetDevice(&prev));
      PMWD_CUDA_TRY(cudaSetDevice(ctx->device));
      long long nn[1] = {n};
      long long embed[1] = {n};
      size_t work = 0;
      PMWD_CUFFT_TRY(cufftCreate(&plan));
      PMWD_CUFFT_TRY(cufftMakePlanMany64(plan, 1, nn, embed, inner, 1, embed, inner, 1, CUFFT_C2C,
                                         inner, &work));
      PMWD_CUDA_TRY(cudaSetDevice(prev));
      ctx->xplans[key] = plan;
    } else {
      plan = it->second;
    }
  }
  PMWD_CUFFT_TRY(cufftSetStream(plan, as_stream(stream)));
  PMWD_CUFFT_TRY(cufftExecC2C(plan, (cufftComplex*)data, (cufftComplex*)data,
                              inverse ? CUFFT_INVERSE : CUFFT_FORWARD));
  return PMWD_OK;
}

// (y, z) 2-D transforms batched over the leading axis of a [n0][n1][n2] real array (plans are
// created by pmwd_ctx_reserve(ctx, 3, shape)): the local part of the slab-decomposed FFT and
// the companions of the fused x-pass.  Unnormalised; C2R clobbers its input.
extern "C" int pmwd_fft2d_r2c(pmwd_ctx* ctx, void* stream, const int32_t* shape, const float* in,
                              void* out) {
  PMWD_REQUIRE(in && out, "null buffer");
  return fft2d_r2c(ctx, as_stream(stream), shape, in, out);
}

extern "C" int pmwd_fft2d_c2r(pmwd_ctx* ctx, void* stream, const int32_t* shape, void* in,
                              float* out) {
  PMWD_REQUIRE(in && out, "null buffer");
  return fft2d_c2r(ctx, as_stream(stream), shape, in, out);
}
