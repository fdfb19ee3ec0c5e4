#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace pmwd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    cached_dev = dev;
  }
  return cached;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- stage profiler -----------------------------------------------------------------
struct Rec { int stage; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<Rec> g_recs;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

StageTimer::StageTimer(int stage, cudaStream_t s) : idx(-1), st(s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  Rec r{stage, get_event(), get_event()};
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  idx = (int)g_recs.size() - 1;
}
StageTimer::~StageTimer() {
  if (idx < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx < (int)g_recs.size()) cudaEventRecord(g_recs[idx].b, st);
}

}  // namespace pmwd

extern "C" long long pmwd_launch_count(void) { return pmwd::g_launches.load(); }

extern "C" int pmwd_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(pmwd::g_prof_mu);
  pmwd::g_prof_on = on != 0;
  return PMWD_OK;
}

extern "C" int pmwd_profile_stage_count(void) { return pmwd::ST_COUNT; }

extern "C" const char* pmwd_profile_stage_name(int stage) {
  static const char* names[pmwd::ST_COUNT] = {
      "memset", "scatter", "fft_r2c", "kspace_force", "fft_c2r", "gather3", "kick_drift",
      "scatter3", "kspace_force_adj", "force_adj_gather", "kick_drift_adj", "other"};
  return (stage >= 0 && stage < pmwd::ST_COUNT) ? names[stage] : "";
}

// Synchronises the device, then accumulates per-stage elapsed ms and call counts and clears
// the records.  ms / calls must hold pmwd_profile_stage_count() entries.
extern "C" int pmwd_profile_read(double* ms, long long* calls) {
  PMWD_CUDA_TRY(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(pmwd::g_prof_mu);
  for (int i = 0; i < pmwd::ST_COUNT; ++i) { ms[i] = 0; calls[i] = 0; }
  for (auto& r : pmwd::g_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.stage] += t; calls[r.stage] += 1; }
    pmwd::g_pool.push_back(r.a);
    pmwd::g_pool.push_back(r.b);
  }
  pmwd::g_recs.clear();
  return PMWD_OK;
}

extern "C" int pmwd_abi_version(void) { return PMWD_B200_ABI_VERSION; }

extern "C" int pmwd_last_error(char* buf, size_t len) {
  size_t n = strlen(pmwd::g_err);
  if (buf && len) {
    size_t m = n < len - 1 ? n : len - 1;
    memcpy(buf, pmwd::g_err, m);
    buf[m] = 0;
  }
  return (int)n;
}
