// Shared host/device helpers for libpmwd_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/pmwd_b200.h"

namespace pmwd {

void set_error(const char* fmt, ...);

#define PMWD_CUDA_TRY(expr)                                                         \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::pmwd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,               \
                        cudaGetErrorString(_e));                                    \
      return (int)_e;                                                               \
    }                                                                               \
  } while (0)

#define PMWD_CUFFT_TRY(expr)                                                        \
  do {                                                                              \
    cufftResult _r = (expr);                                                        \
    if (_r != CUFFT_SUCCESS) {                                                      \
      ::pmwd::set_error("%s:%d: %s -> cufft status %d", __FILE__, __LINE__, #expr,  \
                        (int)_r);                                                   \
      return PMWD_CUFFT_BASE + (int)_r;                                             \
    }                                                                               \
  } while (0)

#define PMWD_REQUIRE(cond, msg)                                                     \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ::pmwd::set_error("%s:%d: invalid argument: %s (%s)", __FILE__, __LINE__,     \
                        msg, #cond);                                                \
      return PMWD_EINVAL;                                                           \
    }                                                                               \
  } while (0)

// every hand-written kernel launch goes through this: error check + launch counter
void count_launch();
#define PMWD_LAUNCH_CHECK()                    \
  do {                                         \
    ::pmwd::count_launch();                    \
    PMWD_CUDA_TRY(cudaPeekAtLastError());      \
  } while (0)

// Optional per-stage CUDA-event profiling (off by default; used by bench.py for the roofline).
enum Stage {
  ST_MEMSET = 0, ST_SCATTER, ST_FFT_R2C, ST_KSPACE, ST_FFT_C2R, ST_GATHER, ST_KICK_DRIFT,
  ST_SCATTER3, ST_KSPACE_ADJ, ST_GATHER_ADJ, ST_KICK_DRIFT_ADJ, ST_OTHER, ST_COUNT
};
struct StageTimer {
  StageTimer(int stage, cudaStream_t st);
  ~StageTimer();
  int idx;
  cudaStream_t st;
};

// B200: 148 SMs.  Grid-stride kernels are launched with a multiple of the SM count.
int sm_count();
inline int grid_for(int64_t work_items, int block, int ctas_per_sm) {
  int64_t need = (work_items + block - 1) / block;
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace pmwd
