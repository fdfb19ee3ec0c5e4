// Public CIC entry points and the fused force pipeline.
//
//   pmwd_force     = gravity()            pmwd/gravity.py:47-72
//   pmwd_force_adj = force_adj()          pmwd/nbody.py:108-118  (gravity + its VJP w.r.t. disp)
//
// Forward force, N_m mesh cells, N_c = n0*n1*(n2/2+1) complex:
//   memset rho -> scatter (RED.F32) -> cuFFT R2C -> fused k-space (1 read, 3 writes; folds the
//   1.5*Omega_m factor of gravity.py:54 and the 1/N_m of irfftn; the `dens -= 1` of
//   gravity.py:52 only changes the k=0 mode, which laplace zeroes) -> 3x cuFFT C2R ->
//   one 3-mesh gather (+ optional fused half-kick).
// Adjoint force additionally:
//   3-channel scatter of pi -> 3x R2C -> fused k-space transpose (3 reads, 1 write) -> C2R ->
//   one weight-gradient gather over (F_0, F_1, F_2, rho_cot).
#include <stdlib.h>

#include "cic.cuh"

struct pmwd_ctx;

namespace pmwd {

// cic_generic.cu
template <int MODE>
int cic_generic(cudaStream_t st, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                const float* a_in, float a_scalar, const float* m_in, float* m_out,
                float* p_out, float* p_out2);
// cic_fast.cu
bool cic_is_fast(const pmwd_cic_desc* d);
bool cic_is_full_mesh(const pmwd_cic_desc* d);
int scatter_fast(cudaStream_t st, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                 const float* val, float val_scalar, int nch, float* m0, float* m1, float* m2);
int gather3_fast(cudaStream_t st, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                 const float* f0, const float* f1, const float* f2, float* acc, float* vel, float K,
                 const float* next_kd, float* disp_rw);
int force_adj_gather(cudaStream_t st, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                     const float* f0, const float* f1, const float* f2, const float* rho_cot,
                     const float* pi, float val, float* alpha, float* acc);
// scatter_sweep.cu
bool sweep_usable(const pmwd_cic_desc* d, const pmwd_sweep* sw);
bool sweep_det_usable(const pmwd_cic_desc* d, const pmwd_sweep* sw);
int scatter_sweep(cudaStream_t st, const pmwd_cic_desc* d, const pmwd_sweep* sw, const void* pmid,
                  const float* disp, const float* val, int vstride, float vscalar, float* mesh,
                  bool reuse_stragglers, bool deterministic);
// scatter_det.cu
size_t scatter_det_scratch_bytes(const pmwd_cic_desc* d);
int scatter_det(cudaStream_t st, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                const float* val, float val_scalar, int nch, float* m0, float* m1, float* m2,
                void* scratch, size_t scratch_bytes);
// fft.cu
int fft_r2c(pmwd_ctx* ctx, cudaStream_t st, int rank, const int32_t* shape, const float* in, void* out);
int fft_c2r(pmwd_ctx* ctx, cudaStream_t st, int rank, const int32_t* shape, void* in, float* out);
int fft2d_r2c(pmwd_ctx* ctx, cudaStream_t st, const int32_t* shape, const float* in, void* out);
int fft2d_c2r(pmwd_ctx* ctx, cudaStream_t st, const int32_t* shape, void* in, float* out);
// xpass.cu
bool xpass_supported(int nx);
int xpass_run(cudaStream_t st, const int32_t* shape, int y0, int ny_l, double spacing, float scale,
              const void* const* in, void* const* out, bool adjoint);

// The fused x-pass pipeline (2-D cuFFT over (y,z) + xpass.cu) replaces {3-D cuFFT, k-space
// kernel, 3-D cuFFT} whenever the mesh's x extent is a supported power of two.
// PMWD_XPASS=0 selects the plain cuFFT-3D pipeline (kept for A/B measurements and odd sizes).
static bool use_xpass(const int32_t* shape) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("PMWD_XPASS");
    env = (e && e[0] == '0') ? 0 : 1;
  }
  return env == 1 && xpass_supported(shape[0]);
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct ForceLayout {
  size_t real_bytes, spec_bytes;
  size_t rho_f0;      // rho, later F_0            (real)
  size_t f1, f2;      // F_1, F_2                  (real)
  size_t rho_k;       // rho_k, later rho_cot_k    (spectrum)
  size_t g[3];        // gradient spectra; adjoint: V_i real, then rho_cot real in g[0]
  size_t s[3];        // adjoint only: V_i spectra
  size_t det;         // deterministic-scatter scratch
  size_t total;
};

static void force_layout(const pmwd_cic_desc* d, int adjoint, int mode, ForceLayout* L) {
  int64_t nm = (int64_t)d->mesh_shape[0] * d->mesh_shape[1] * d->mesh_shape[2];
  int64_t nc = (int64_t)d->mesh_shape[0] * d->mesh_shape[1] * (d->mesh_shape[2] / 2 + 1);
  L->real_bytes = align_up((size_t)nm * sizeof(float));
  L->spec_bytes = align_up((size_t)nc * sizeof(float2));
  size_t off = 0;
  L->rho_f0 = off; off += L->real_bytes;
  L->f1 = off;     off += L->real_bytes;
  L->f2 = off;     off += L->real_bytes;
  L->rho_k = off;  off += L->spec_bytes;
  for (int a = 0; a < 3; ++a) { L->g[a] = off; off += L->spec_bytes; }
  for (int a = 0; a < 3; ++a) { L->s[a] = off; if (adjoint) off += L->spec_bytes; }
  L->det = off;
  if (mode == PMWD_SCATTER_DETERMINISTIC) off += align_up(scatter_det_scratch_bytes(d));
  L->total = off;
}

static int check_force_args(const pmwd_cic_desc* d) {
  PMWD_REQUIRE(d != nullptr, "null descriptor");
  PMWD_REQUIRE(cic_is_full_mesh(d), "pmwd_force needs the 3-D int16 fast path (offset 0, cell_size=None, mesh == conf.mesh_shape)");
  PMWD_REQUIRE(d->nchan == 1, "pmwd_force: nchan must be 1");
  return PMWD_OK;
}

static int force_forward(pmwd_ctx* ctx, cudaStream_t st, const pmwd_cic_desc* d, const void* pmid,
                         const float* disp, double Omega_m, int mode, char* ws,
                         const ForceLayout& L, float* val_out, const pmwd_sweep* sweep) {
  const int32_t* shape = d->mesh_shape;
  const int64_t nm = (int64_t)shape[0] * shape[1] * shape[2];
  float* rho = (float*)(ws + L.rho_f0);
  // scatter.py:37-39: val = conf.mesh_size / conf.ptcl_num (python float -> float32)
  const float val = (float)((double)nm / (double)d->ptcl_num);
  if (val_out) *val_out = val;
  // tiled sweep scatter (scatter_sweep.cu): overwrites rho, no memset; in deterministic mode its
  // reproducible variant if the caller's sweep descriptor carries the halo arrays (the integrator, which
  // re-sorts the storage before every deposit), else the cell-sorted scatter of scatter_det.cu
  const bool det = mode == PMWD_SCATTER_DETERMINISTIC;
  const bool sweeping = det ? sweep_det_usable(d, sweep) : sweep_usable(d, sweep);
  if (!sweeping) {
    StageTimer t(ST_MEMSET, st);
    PMWD_CUDA_TRY(cudaMemsetAsync(rho, 0, (size_t)nm * sizeof(float), st));
  }
  int rc;
  {
    StageTimer t(ST_SCATTER, st);
    if (sweeping)
      rc = scatter_sweep(st, d, sweep, pmid, disp, nullptr, 0, val, rho, false, det);
    else if (mode == PMWD_SCATTER_DETERMINISTIC)
      rc = scatter_det(st, d, pmid, disp, nullptr, val, 1, rho, nullptr, nullptr, ws + L.det,
                       L.total - L.det);
    else
      rc = scatter_fast(st, d, pmid, disp, nullptr, val, 1, rho, nullptr, nullptr);
  }
  if (rc) return rc;
  const bool xp = use_xpass(shape);
  {
    StageTimer t(ST_FFT_R2C, st);
    rc = xp ? fft2d_r2c(ctx, st, shape, rho, ws + L.rho_k) : fft_r2c(ctx, st, 3, shape, rho, ws + L.rho_k);
  }
  if (rc) return rc;
  void* g[3] = {ws + L.g[0], ws + L.g[1], ws + L.g[2]};
  const float scale = (float)(1.5 * Omega_m / (double)nm);
  {
    StageTimer t(ST_KSPACE, st);
    const void* in[3] = {ws + L.rho_k, nullptr, nullptr};
    rc = xp ? xpass_run(st, shape, 0, shape[1], d->cell_size, scale, in, g, false)
            : pmwd_kspace_force(st, 3, shape, d->cell_size, scale, ws + L.rho_k, g);
  }
  if (rc) return rc;
  float* F[3] = {(float*)(ws + L.rho_f0), (float*)(ws + L.f1), (float*)(ws + L.f2)};
  for (int a = 0; a < 3; ++a) {
    StageTimer t(ST_FFT_C2R, st);
    rc = xp ? fft2d_c2r(ctx, st, shape, g[a], F[a]) : fft_c2r(ctx, st, 3, shape, g[a], F[a]);
    if (rc) return rc;
  }
  return PMWD_OK;
}

}  // namespace pmwd

using namespace pmwd;

// 1 if the descriptor qualifies for the fused 3-D kernels (int16 pmid, cell_size=None, offset =
// whole planes along x only, full y / z extents); 2 if it additionally spans the whole mesh
extern "C" int pmwd_cic_fast_path(const pmwd_cic_desc* d) {
  if (!cic_is_fast(d)) return 0;
  return cic_is_full_mesh(d) ? 2 : 1;
}

extern "C" size_t pmwd_scatter_scratch_bytes(const pmwd_cic_desc* d, int mode) {
  if (mode != PMWD_SCATTER_DETERMINISTIC) return 0;
  return scatter_det_scratch_bytes(d);
}

extern "C" int pmwd_scatter(void* stream, const pmwd_cic_desc* d, const void* pmid,
                            const float* disp, const float* val, float val_scalar, float* mesh,
                            int mode, void* scratch, size_t scratch_bytes) {
  PMWD_REQUIRE(d && mesh && (d->ptcl_num == 0 || (pmid && disp)), "null buffer");
  cudaStream_t st = as_stream(stream);
  if (mode == PMWD_SCATTER_DETERMINISTIC) {
    PMWD_REQUIRE(d->nchan == 1, "deterministic scatter via pmwd_scatter supports scalar fields");
    return scatter_det(st, d, pmid, disp, val, val_scalar, 1, mesh, nullptr, nullptr, scratch,
                       scratch_bytes);
  }
  PMWD_REQUIRE(mode == PMWD_SCATTER_ATOMIC, "unknown scatter mode");
  if (cic_is_fast(d) && d->nchan == 1)
    return scatter_fast(st, d, pmid, disp, val, val_scalar, 1, mesh, nullptr, nullptr);
  return cic_generic<0>(st, d, pmid, disp, val, val_scalar, nullptr, mesh, nullptr, nullptr);
}

extern "C" int pmwd_gather(void* stream, const pmwd_cic_desc* d, const void* pmid,
                           const float* disp, const float* mesh, const float* val,
                           float val_scalar, float* out) {
  PMWD_REQUIRE(d && mesh && (d->ptcl_num == 0 || (pmid && disp && out)), "null buffer");
  return cic_generic<1>(as_stream(stream), d, pmid, disp, val, val_scalar, mesh, nullptr, out, nullptr);
}

extern "C" int pmwd_scatter_adj(void* stream, const pmwd_cic_desc* d, const void* pmid,
                                const float* disp, const float* mesh_cot, const float* val,
                                float val_scalar, float* disp_cot, float* val_cot) {
  PMWD_REQUIRE(d && mesh_cot && (d->ptcl_num == 0 || (pmid && disp && disp_cot)), "null buffer");
  return cic_generic<2>(as_stream(stream), d, pmid, disp, val, val_scalar, mesh_cot, nullptr,
                        disp_cot, val_cot);
}

extern "C" int pmwd_gather_adj(void* stream, const pmwd_cic_desc* d, const void* pmid,
                               const float* disp, const float* mesh, const float* val_cot,
                               float val_cot_scalar, float* disp_cot, float* mesh_cot) {
  PMWD_REQUIRE(d && mesh && (d->ptcl_num == 0 || (pmid && disp && disp_cot)), "null buffer");
  return cic_generic<3>(as_stream(stream), d, pmid, disp, val_cot, val_cot_scalar, mesh, mesh_cot,
                        disp_cot, nullptr);
}

// ---- SoA building blocks of the force pipeline, exposed for the slab-decomposed (multi-GPU)
// composition in pmwd_b200/dist.py: the mesh arrays may be x-slabs with halos, described by
// d->mesh_shape[0] (planes held) and d->offset[0] (= first plane * cell_size).
extern "C" int pmwd_scatter_soa(void* stream, const pmwd_cic_desc* d, const void* pmid,
                                const float* disp, const float* val, float val_scalar, int nch,
                                float* m0, float* m1, float* m2) {
  PMWD_REQUIRE(d && m0 && (d->ptcl_num == 0 || (pmid && disp)), "null buffer");
  PMWD_REQUIRE(nch == 1 || (nch == 3 && m1 && m2), "nch must be 1 or 3 (with three meshes)");
  StageTimer t(nch == 1 ? ST_SCATTER : ST_SCATTER3, as_stream(stream));
  return scatter_fast(as_stream(stream), d, pmid, disp, val, val_scalar, nch, m0, m1, m2);
}

extern "C" int pmwd_gather3(void* stream, const pmwd_cic_desc* d, const void* pmid,
                            const float* disp, const float* f0, const float* f1, const float* f2,
                            float* acc, float* kick_vel, float kick_factor) {
  PMWD_REQUIRE(d && f0 && f1 && f2 && (d->ptcl_num == 0 || (pmid && disp && acc)), "null buffer");
  StageTimer t(ST_GATHER, as_stream(stream));
  return gather3_fast(as_stream(stream), d, pmid, disp, f0, f1, f2, acc, kick_vel, kick_factor, nullptr, nullptr);
}

// pmwd_gather3 + this step's trailing half-kick + the next step's leading half-kick and drift
// (see pmwd_force_kdk); disp is updated in place.
extern "C" int pmwd_gather3_kdk(void* stream, const pmwd_cic_desc* d, const void* pmid, float* disp,
                                const float* f0, const float* f1, const float* f2, float* acc,
                                float* vel, float K2, float K1_next, float D_next) {
  PMWD_REQUIRE(d && f0 && f1 && f2 && (d->ptcl_num == 0 || (pmid && disp && acc && vel)), "null buffer");
  StageTimer t(ST_GATHER, as_stream(stream));
  const float next_kd[2] = {K1_next, D_next};
  return gather3_fast(as_stream(stream), d, pmid, disp, f0, f1, f2, acc, vel, K2, next_kd, disp);
}

extern "C" int pmwd_force_adj_gather(void* stream, const pmwd_cic_desc* d, const void* pmid,
                                     const float* disp, const float* f0, const float* f1,
                                     const float* f2, const float* rho_cot, const float* pi,
                                     float val, float* alpha) {
  PMWD_REQUIRE(d && f0 && f1 && f2 && rho_cot && (d->ptcl_num == 0 || (pmid && disp && pi && alpha)),
               "null buffer");
  StageTimer t(ST_GATHER_ADJ, as_stream(stream));
  return force_adj_gather(as_stream(stream), d, pmid, disp, f0, f1, f2, rho_cot, pi, val, alpha, nullptr);
}

extern "C" size_t pmwd_force_workspace_bytes(const pmwd_cic_desc* d, int adjoint, int mode) {
  if (!d || d->dim != 3) return 0;
  ForceLayout L;
  force_layout(d, adjoint, mode, &L);
  return L.total;
}

extern "C" int pmwd_force(pmwd_ctx* ctx, void* stream, const pmwd_cic_desc* d, const void* pmid,
                          const float* disp, double Omega_m, float* acc, float* kick_vel,
                          float kick_factor, int mode, void* workspace, size_t workspace_bytes,
                          const pmwd_sweep* sweep) {
  int rc = check_force_args(d);
  if (rc) return rc;
  PMWD_REQUIRE(ctx && pmid && disp && acc && workspace, "null buffer");
  PMWD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  ForceLayout L;
  force_layout(d, 0, mode, &L);
  if (workspace_bytes < L.total) {
    set_error("pmwd_force needs %zu workspace bytes, got %zu", L.total, workspace_bytes);
    return PMWD_ENOMEM;
  }
  cudaStream_t st = as_stream(stream);
  char* ws = (char*)workspace;
  rc = force_forward(ctx, st, d, pmid, disp, Omega_m, mode, ws, L, nullptr, sweep);
  if (rc) return rc;
  StageTimer t(ST_GATHER, st);
  return gather3_fast(st, d, pmid, disp, (float*)(ws + L.rho_f0), (float*)(ws + L.f1),
                      (float*)(ws + L.f2), acc, kick_vel, kick_factor, nullptr, nullptr);
}

// One whole KDK step's particle update behind the force (pmwd/nbody.py:121-140 pipelined):
//   acc = gravity(disp); vel += acc*K2;   [this step's trailing half-kick]
//   vel += acc*K1_next; disp += vel*D_next  [the NEXT step's leading half-kick and drift]
// in the gather pass, so that a step is exactly one pmwd_force_kdk call.  Same float32
// operation sequence as pmwd_force(+kick) followed by pmwd_kick_drift.  `disp` is updated in
// place AFTER the force has been evaluated at its incoming value.
extern "C" int pmwd_force_kdk(pmwd_ctx* ctx, void* stream, const pmwd_cic_desc* d, const void* pmid,
                              float* disp, double Omega_m, float* acc, float* vel, float K2,
                              float K1_next, float D_next, int mode, void* workspace,
                              size_t workspace_bytes, const pmwd_sweep* sweep) {
  int rc = check_force_args(d);
  if (rc) return rc;
  PMWD_REQUIRE(ctx && pmid && disp && acc && vel && workspace, "null buffer");
  PMWD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  ForceLayout L;
  force_layout(d, 0, mode, &L);
  if (workspace_bytes < L.total) {
    set_error("pmwd_force_kdk needs %zu workspace bytes, got %zu", L.total, workspace_bytes);
    return PMWD_ENOMEM;
  }
  cudaStream_t st = as_stream(stream);
  char* ws = (char*)workspace;
  rc = force_forward(ctx, st, d, pmid, disp, Omega_m, mode, ws, L, nullptr, sweep);
  if (rc) return rc;
  const float next_kd[2] = {K1_next, D_next};
  StageTimer t(ST_GATHER, st);
  return gather3_fast(st, d, pmid, disp, (float*)(ws + L.rho_f0), (float*)(ws + L.f1),
                      (float*)(ws + L.f2), acc, vel, K2, next_kd, disp);
}

extern "C" int pmwd_force_adj(pmwd_ctx* ctx, void* stream, const pmwd_cic_desc* d,
                              const void* pmid, const float* disp, double Omega_m,
                              const float* pi, float* acc, float* alpha, int mode,
                              void* workspace, size_t workspace_bytes, const pmwd_sweep* sweep) {
  int rc = check_force_args(d);
  if (rc) return rc;
  PMWD_REQUIRE(ctx && pmid && disp && pi && acc && alpha && workspace, "null buffer");
  PMWD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  ForceLayout L;
  force_layout(d, 1, mode, &L);
  if (workspace_bytes < L.total) {
    set_error("pmwd_force_adj needs %zu workspace bytes, got %zu", L.total, workspace_bytes);
    return PMWD_ENOMEM;
  }
  cudaStream_t st = as_stream(stream);
  char* ws = (char*)workspace;
  const int32_t* shape = d->mesh_shape;
  const int64_t nm = (int64_t)shape[0] * shape[1] * shape[2];
  float val = 0.f;
  rc = force_forward(ctx, st, d, pmid, disp, Omega_m, mode, ws, L, &val, sweep);
  if (rc) return rc;
  // acc = gather3(F) is evaluated in the weight-gradient gather at the end (same force values, same
  // float32 arithmetic): one pass over the particles and the three force meshes instead of two
  float* F[3] = {(float*)(ws + L.rho_f0), (float*)(ws + L.f1), (float*)(ws + L.f2)};

  // V_i = scatter(pi_i): the mesh_cot of _gather_bwd (gather.py:113), SoA, in the (now free)
  // gradient-spectrum buffers
  float* V[3] = {(float*)(ws + L.g[0]), (float*)(ws + L.g[1]), (float*)(ws + L.g[2])};
  const bool det = mode == PMWD_SCATTER_DETERMINISTIC;
  const bool sweeping = det ? sweep_det_usable(d, sweep) : sweep_usable(d, sweep);
  if (!sweeping) {
    StageTimer t(ST_MEMSET, st);
    for (int a = 0; a < 3; ++a) PMWD_CUDA_TRY(cudaMemsetAsync(V[a], 0, (size_t)nm * sizeof(float), st));
  }
  {
    StageTimer t(ST_SCATTER3, st);
    if (sweeping) {
      // one sweep per channel; the straggler list recorded by the density scatter is reused
      for (int a = 0; a < 3 && !rc; ++a) rc = scatter_sweep(st, d, sweep, pmid, disp, pi + a, 3, 0.f, V[a], true, det);
    } else if (mode == PMWD_SCATTER_DETERMINISTIC)
      rc = scatter_det(st, d, pmid, disp, pi, 0.f, 3, V[0], V[1], V[2], ws + L.det, L.total - L.det);
    else
      rc = scatter_fast(st, d, pmid, disp, pi, 0.f, 3, V[0], V[1], V[2]);
  }
  if (rc) return rc;
  const void* S[3] = {ws + L.s[0], ws + L.s[1], ws + L.s[2]};
  const bool xp = use_xpass(shape);
  for (int a = 0; a < 3; ++a) {
    StageTimer t(ST_FFT_R2C, st);
    rc = xp ? fft2d_r2c(ctx, st, shape, V[a], ws + L.s[a]) : fft_r2c(ctx, st, 3, shape, V[a], ws + L.s[a]);
    if (rc) return rc;
  }
  // rho_cot_k = (1.5 Omega_m / N_m) * sum_i (+i k_i)(-V_i,k / k^2)   [A_i^T = -A_i]
  const float scale = (float)(1.5 * Omega_m / (double)nm);
  {
    StageTimer t(ST_KSPACE_ADJ, st);
    void* out[3] = {ws + L.rho_k, nullptr, nullptr};
    rc = xp ? xpass_run(st, shape, 0, shape[1], d->cell_size, scale, S, out, true)
            : pmwd_kspace_force_adj(st, 3, shape, d->cell_size, scale, S, ws + L.rho_k);
  }
  if (rc) return rc;
  float* rho_cot = (float*)(ws + L.g[0]);
  {
    StageTimer t(ST_FFT_C2R, st);
    rc = xp ? fft2d_c2r(ctx, st, shape, ws + L.rho_k, rho_cot)
            : fft_c2r(ctx, st, 3, shape, ws + L.rho_k, rho_cot);
  }
  if (rc) return rc;
  StageTimer t(ST_GATHER_ADJ, st);
  return force_adj_gather(st, d, pmid, disp, F[0], F[1], F[2], rho_cot, pi, val, alpha, acc);
}
