/* pmwd_b200.h -- C ABI of the B200-native particle-mesh hot path.
 *
 * Drop-in boundary for the hot path of eelregit/pmwd.  The reference has no native
 * interface (it is pure JAX); the entry points below are what a `jax.ffi` custom
 * call (or any other host: ctypes, torch, C++) binds in place of the XLA-generated
 * scatter-add / gather / elementwise / FFT ops.  Each entry cites the reference
 * function it replaces (paths relative to the pmwd repository root).
 *
 * Conventions (XLA-FFI shaped):
 *   - every pointer is a DEVICE pointer owned by the caller unless stated otherwise;
 *   - calls only ENQUEUE work on `stream` (a cudaStream_t passed as void*): no
 *     synchronisation, no allocation, no use of the default stream;
 *     the only exceptions are pmwd_ctx_create / pmwd_ctx_reserve / pmwd_ctx_destroy;
 *   - return value 0 = success; negative = bad argument (PMWD_E*); positive =
 *     CUDA / cuFFT status (offset by PMWD_CUFFT_BASE for cuFFT).  A message is
 *     retrievable with pmwd_last_error() (thread-local);
 *   - arrays are C-order.  Particles are AoS rows: pmid int{8,16,32}[N][dim],
 *     disp/vel/acc float[N][dim] (pmwd/particles.py:52-59).  Meshes are
 *     float[n0][n1][n2](+channels last); spectra are float2[n0][n1][n2/2+1];
 *   - there is NO CPU fallback anywhere behind this ABI.
 */
#ifndef PMWD_B200_H_
#define PMWD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMWD_B200_ABI_VERSION 2

enum {
  PMWD_OK = 0,
  PMWD_EINVAL = -1,      /* bad argument / unsupported combination */
  PMWD_ENOMEM = -2,      /* caller workspace too small */
  PMWD_ESTATE = -3,      /* context not prepared for this shape */
  PMWD_CUFFT_BASE = 100000
};

/* Scatter accumulation modes (pmwd/scatter.py:80 is an unordered f32 scatter-add on GPU). */
enum {
  PMWD_SCATTER_ATOMIC = 0,        /* f32 red.global.add (x2-vectorised where aligned); fastest; run-to-run ulp noise */
  PMWD_SCATTER_DETERMINISTIC = 1  /* cell-sorted: stable radix sort by cell + fixed-order segmented sum; bitwise reproducible */
};

/* Geometry of one CIC transfer: mirrors the arguments that scatter/gather hand to
 * enmesh (pmwd/pm_util.py:33-91; pmwd/scatter.py:69-71). */
typedef struct pmwd_cic_desc {
  int32_t dim;            /* 1, 2 or 3 */
  int32_t pmid_bytes;     /* 1, 2 (reference default int16) or 4 */
  int64_t ptcl_num;
  int32_t wrap_shape[3];  /* conf.mesh_shape: periodic wrap of indices (enmesh s1) */
  int32_t mesh_shape[3];  /* spatial shape of the mesh array (enmesh s2); out-of-range neighbours are dropped */
  int32_t nchan;          /* product of the channel dims (1 = scalar field); channels last, interleaved */
  int32_t general;        /* 0: fast float32 branch (cell_size=None, pm_util.py:119-136); 1: float64 branch (pm_util.py:99-118) */
  double cell_size;       /* conf.cell_size (enmesh a1) */
  double cell_size2;      /* user cell_size (enmesh a2); ignored unless general */
  double offset[3];       /* enmesh b12 */
} pmwd_cic_desc;

/* Tiled ("tile sweep") CIC deposit, csrc/scatter_sweep.cu: particle storage sorted by
 * (y-tile of `ty` rows, z-tile of `bw` cells, x-plane, y, z) + a table of the particle range of every
 * (tile, plane).  Passed to pmwd_force / pmwd_force_kdk / pmwd_force_adj as an optional argument
 * (NULL, or a descriptor that does not match the mesh: the per-particle RED kernel is used). */
typedef struct pmwd_sweep {
  const uint32_t* table;  /* device, [ny / ty][nz / bw][planes][2] = (first, one-past-last) particle slot */
  int32_t ty;             /* tile rows (pmwd_sweep_pick) */
  int32_t bw;             /* tile cells along z */
  int32_t lx;             /* planes per x segment of a work item (64 is a good value) */
  int32_t nx_ext;         /* planes of the mesh array the table was built for (== desc mesh_shape[0]) */
  int32_t xoff;           /* global index of its plane 0 (offset[0] / cell) */
  int32_t reserved;
  void* scratch;          /* device, pmwd_sweep_scratch_bytes: work counters + straggler list */
  size_t scratch_bytes;
  void* det_halo;         /* device, pmwd_sweep_det_halo_bytes, or NULL: per-tile halo arrays of the deterministic deposit */
  size_t det_halo_bytes;
} pmwd_sweep;

/* Which kernels a descriptor selects: 0 = general path (csrc/cic_generic.cu); 1 = fused 3-D fast
 * path on an x-slab (int16 pmid, cell_size=None, offset = whole float32-cell planes along x, full
 * y and z extents -- the multi-GPU slabs); 2 = fast path on the whole mesh (what pmwd_force needs). */
int pmwd_cic_fast_path(const pmwd_cic_desc* d);

/* ---- library / context ------------------------------------------------------------ */

int pmwd_abi_version(void);
/* Copies the calling thread's last error message; returns its length. */
int pmwd_last_error(char* buf, size_t len);

typedef struct pmwd_ctx pmwd_ctx;
/* A context owns the cuFFT plans (keyed by shape) and their shared work area on one
 * device.  Plans are created by pmwd_ctx_reserve (which allocates and may synchronise);
 * all other calls taking a ctx only enqueue.  One context must not be used from two
 * streams concurrently. */
int pmwd_ctx_create(pmwd_ctx** ctx, int device);
int pmwd_ctx_destroy(pmwd_ctx* ctx);
/* Create R2C + C2R plans for a real field of `rank` dims `shape` (float). */
int pmwd_ctx_reserve(pmwd_ctx* ctx, int rank, const int32_t* shape);
/* The (y, z) 2-D transforms (pmwd_fft2d_*, and inside pmwd_force*) can be issued `planes` x planes per cuFFT
 * call (so that the second pass of a 2-D transform may find the first pass's output in L2).  Applies to
 * plans reserved after the call; 0 = one call for the whole batch, -1 (default) = the PMWD_FFT2D_CHUNK
 * environment variable, else 0 (measured on B200: chunking is slower, profiles/r02_fft2d_chunks.txt). */
int pmwd_ctx_set_fft2d_chunk(pmwd_ctx* ctx, int planes);
/* Lab knob: row length (complex elements, >= nz/2+1, 0 = unpadded) of the half-spectrum of 2-D plans reserved
 * afterwards (measures what 128-byte aligned spectrum rows would buy; pmwd_force* always use unpadded rows). */
int pmwd_ctx_set_fft2d_pad(pmwd_ctx* ctx, int row_elems);

/* ---- instrumentation (no reference counterpart; used by bench.py) -------------------- */
/* Number of hand-written kernels launched by this process so far. */
long long pmwd_launch_count(void);
/* Per-stage CUDA-event timing of the kernels enqueued by pmwd_force / pmwd_force_adj /
 * pmwd_kick_drift(_adj) / pmwd_scatter: events are recorded on the caller's stream around
 * each stage while enabled.  pmwd_profile_read synchronises the device, returns accumulated
 * milliseconds and call counts per stage (arrays of pmwd_profile_stage_count() entries) and
 * resets the records. */
int pmwd_profile_enable(int on);
int pmwd_profile_stage_count(void);
const char* pmwd_profile_stage_name(int stage);
int pmwd_profile_read(double* ms, long long* calls);

/* ---- FFT: pmwd/pm_util.py:236-344 (fftfwd / fftinv = rfftn / irfftn) -------------- */
/* out[n0][n1][n2/2+1] = rfftn(in), unnormalised (pmwd/pm_util.py:281). */
int pmwd_fft_r2c(pmwd_ctx* ctx, void* stream, int rank, const int32_t* shape,
                 const float* in, void* out_c64);
/* out = irfftn(in) * scale; pass scale = 1/prod(shape) for numpy 'backward' norm
 * (pmwd/pm_util.py:336).  `in` is clobbered (cuFFT C2R semantics). */
int pmwd_fft_c2r(pmwd_ctx* ctx, void* stream, int rank, const int32_t* shape,
                 void* in_c64, float* out, float scale);

/* In-place 1-D complex transforms along the leading axis of data[n][inner] (stride = inner):
 * the x-pass of the slab-decomposed FFT.  Unnormalised in both directions.  The plan for a
 * given (n, inner) is created (allocating) on first use. */
int pmwd_fft_c2c_lead(pmwd_ctx* ctx, void* stream, int n, long long inner, void* data_c64,
                      int inverse);

/* (y, z) 2-D transforms batched over the leading axis of in[n0][n1][n2] (real) <->
 * out[n0][n1][n2/2+1]; plans come from pmwd_ctx_reserve(ctx, 3, shape).  Unnormalised. */
int pmwd_fft2d_r2c(pmwd_ctx* ctx, void* stream, const int32_t* shape, const float* in, void* out_c64);
int pmwd_fft2d_c2r(pmwd_ctx* ctx, void* stream, const int32_t* shape, void* in_c64, float* out);

/* ---- CIC scatter / gather and their VJPs ------------------------------------------ */
/* _scatter (pmwd/scatter.py:33-83): mesh[ind] += val * frac.  `val` is either a device
 * array float[N][nchan] or NULL, in which case `val_scalar` is broadcast (0-D val).
 * `mesh` is accumulated in place (the caller zero-fills or copies the input mesh).
 * Deterministic mode needs scratch (pmwd_scatter_scratch_bytes); atomic mode none. */
size_t pmwd_scatter_scratch_bytes(const pmwd_cic_desc* d, int mode);
int pmwd_scatter(void* stream, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                 const float* val, float val_scalar, float* mesh, int mode,
                 void* scratch, size_t scratch_bytes);

/* _gather (pmwd/gather.py:33-77): out = val + sum_n mesh[ind] * frac. */
int pmwd_gather(void* stream, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                const float* mesh, const float* val, float val_scalar, float* out);

/* _scatter_bwd (pmwd/scatter.py:86-148): disp_cot[N][dim], val_cot[N][nchan] (may be NULL). */
int pmwd_scatter_adj(void* stream, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                     const float* mesh_cot, const float* val, float val_scalar,
                     float* disp_cot, float* val_cot);

/* _gather_bwd (pmwd/gather.py:80-142): disp_cot[N][dim]; mesh_cot accumulated in place
 * (caller zero-fills; may be NULL to skip).  `val_cot` float[N][nchan] or NULL+scalar. */
int pmwd_gather_adj(void* stream, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                    const float* mesh, const float* val_cot, float val_cot_scalar,
                    float* disp_cot, float* mesh_cot);

/* ---- k-space kernels: pmwd/gravity.py:9-44, pmwd/lpt.py:13-37 ---------------------- */
/* laplace (gravity.py:9-16): pot = where(k2 != 0, -src/k2, 0); k from fftfreq
 * (pm_util.py:159-199) with grid `spacing`.  In-place allowed. */
int pmwd_laplace(void* stream, int rank, const int32_t* shape, double spacing,
                 const void* src_c64, void* pot_c64);
/* neg_grad (gravity.py:37-44): out = -i k_axis pot, Nyquist planes zeroed. In-place allowed. */
int pmwd_neg_grad(void* stream, int rank, const int32_t* shape, double spacing, int axis,
                  const void* pot_c64, void* out_c64);
/* Fused forward force spectrum (gravity.py:54-62): for a density spectrum rho_k,
 * g_j = -i k_j * ( -(scale * rho_k) / k^2 ), j = 0..rank-1, in one pass. */
int pmwd_kspace_force(void* stream, int rank, const int32_t* shape, double spacing,
                      float scale, const void* rho_c64, void* const* g_c64);
/* Fused transpose (what jax.vjp(gravity) evaluates between the gather and scatter VJPs,
 * pmwd/nbody.py:111-116): out = scale * sum_j (+i k_j) * ( -v_j / k^2 ). */
int pmwd_kspace_force_adj(void* stream, int rank, const int32_t* shape, double spacing,
                          float scale, const void* const* v_c64, void* out_c64);
/* _strain spectrum (lpt.py:13-32): out = -k_i k_j pot (Nyquist zeroed when i != j). */
int pmwd_strain(void* stream, int rank, const int32_t* shape, double spacing, int i, int j,
                const void* pot_c64, void* out_c64);

/* powspec binning (spec_util.py:50-147): one pass over a half spectrum, float64 sums per np.digitize
 * bin accumulated (+=) into out_f64[4][nedges + 1] = {sum k N, sum Re P N, sum Im P N, sum N}, where
 * P = |f|^2 or f conj(g) (g_c64 may be NULL), optionally times prod_a sinc(k_a)^-deconv, N the Hermitian
 * multiplicity, k in cycles per grid unit; edges_f64: nedges (<= 512) ascending device float64. */
int pmwd_powspec_bin(void* stream, const int32_t* shape, const void* f_c64, const void* g_c64,
                     int has_deconv, double deconv, const double* edges_f64, int nedges, int right,
                     double* out_f64);
/* k-space half of the auto spectrum's VJP w.r.t. the field: out = w_k f_k with
 * w_k = wbin_f64[bin(k)] * prod_a sinc(k_a)^-deconv (wbin_f64: nedges + 1 device float64, indexed like the
 * bins above, = Pbar_b * spacing^3 / N_total / N_b); the field cotangent is twice the unnormalised
 * C2R transform of out.  out_c64 may alias f_c64. */
int pmwd_powspec_weight(void* stream, const int32_t* shape, const void* f_c64, int has_deconv,
                        double deconv, const double* edges_f64, int nedges, int right,
                        const double* wbin_f64, void* out_c64);

/* ---- fused elementwise passes of LPT: pmwd/lpt.py:40-76, 190-208 ---------------------- */
/* 2LPT source (m == n) from the six real strain fields s = {s00, s11, s22, s01, s02, s12}[n]:
 * L = s00 s22 + s00 s11 + s11 s22 - s01^2 - s02^2 - s12^2 in the reference's summation order; and
 * its VJP s_cot[k] = dL/ds_k * L_cot. */
int pmwd_lpt_source2(void* stream, int64_t n, const float* const* s, float* L);
int pmwd_lpt_source2_vjp(void* stream, int64_t n, const float* const* s, const float* L_cot,
                         float* const* s_cot);
/* disp[p][a] = (disp0[p][a] + D1 g1[a][p]) + D2 g2[a][p]; vel likewise with V1, V2 (lpt.py:203-208,
 * order 1 then order 2; g2 == NULL for first-order LPT).  VJP: g1_cot[a] = D1 disp_cot[:, a] +
 * V1 vel_cot[:, a] (g2 likewise) and sums[4] += (sum disp_cot g1, sum vel_cot g1, sum disp_cot g2,
 * sum vel_cot g2) in float64 = the cotangents of D1, V1, D2, V2. */
int pmwd_lpt_displace(void* stream, int64_t n, const float* disp0, const float* vel0,
                      const float* const* g1, const float* const* g2, float D1, float V1, float D2,
                      float V2, float* disp, float* vel);
int pmwd_lpt_displace_vjp(void* stream, int64_t n, const float* disp_cot, const float* vel_cot,
                          const float* const* g1, const float* const* g2, float D1, float V1, float D2,
                          float V2, float* const* g1_cot, float* const* g2_cot, double* sums);

/* ---- slab-decomposed (multi-GPU) building blocks ------------------------------------- */
/* The reference's own (offset, mesh shape) semantics describe a slab: a mesh array holding
 * d->mesh_shape[0] x-planes starting at global plane d->offset[0] / cell_size of the periodic
 * mesh d->wrap_shape; neighbours outside the array are dropped (pm_util.py:119-141).
 * SoA scatter of 1 or 3 channels (rho, or the three V_i = scatter(pi_i) of gather.py:113): */
int pmwd_scatter_soa(void* stream, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                     const float* val, float val_scalar, int nch, float* m0, float* m1, float* m2);
/* The same deposit through the tiled sweep kernels (see pmwd_sweep): the meshes are OVERWRITTEN
 * (no memset by the caller); nch = 3 needs per-particle values float[N][3]. */
int pmwd_scatter_sweep(void* stream, const pmwd_cic_desc* d, const pmwd_sweep* sweep, const void* pmid,
                       const float* disp, const float* val, float val_scalar, int nch, float* m0,
                       float* m1, float* m2);
/* Tile shape for this mesh (returns 0 if the sweep kernels do not support it), size of the
 * (tile, plane) table and of the scratch area (work counters + straggler list). */
int pmwd_sweep_pick(const pmwd_cic_desc* d, int* ty, int* bw);
size_t pmwd_sweep_table_bytes(const pmwd_cic_desc* d, int ty, int bw);
size_t pmwd_sweep_scratch_bytes(const pmwd_cic_desc* d);
/* Build the table from the sorted keys of pmwd_cell_sort_perm(..., ty, bw) (pmwd_cell_sort_sorted_keys).
 * status (device, 16 bytes): after synchronising, uint32 [0] == 0 and uint64 [1] == ptcl_num for a
 * valid table (true by construction for sorted keys). */
int pmwd_sweep_table(void* stream, const pmwd_cic_desc* d, int ty, int bw, const uint32_t* keys,
                     uint32_t* table, void* status);
/* Deterministic variant (PMWD_SCATTER_DETERMINISTIC inside pmwd_force* when sweep->det_halo is set): one warp
 * sweeps one (y, z) tile over all planes, tile-interior cells are stored, halo contributions go to per-tile
 * arrays that a second kernel adds in a fixed order -- no float atomics, bitwise reproducible for storage that
 * was sorted (pmwd_cell_sort_perm + pmwd_sweep_table) from the very pmid / disp it is called with.  A particle
 * outside its tile's window would still be deposited (by REDs) and is counted: pmwd_sweep_det_violations
 * (synchronises) must return 0; pmwd_sweep_det_reset clears the counter.  Whole periodic mesh only. */
int pmwd_scatter_sweep_det(void* stream, const pmwd_cic_desc* d, const pmwd_sweep* sweep, const void* pmid,
                           const float* disp, const float* val, float val_scalar, int nch, float* m0,
                           float* m1, float* m2);
size_t pmwd_sweep_det_halo_bytes(const pmwd_cic_desc* d, int ty, int bw);
long long pmwd_sweep_det_violations(void* stream, const pmwd_sweep* sweep);
int pmwd_sweep_det_reset(void* stream, const pmwd_sweep* sweep);
/* 1 if `sweep` matches the descriptor (same slab planes, scratch large enough): the tiled kernels would run. */
int pmwd_sweep_usable(const pmwd_cic_desc* d, const pmwd_sweep* sweep);
/* Stragglers (particles outside their tile's window) of the last recording sweep; synchronises. */
long long pmwd_sweep_last_stragglers(void* stream, const pmwd_sweep* sweep);
/* acc[N][3] = (gather(f0), gather(f1), gather(f2)) in one pass (gravity.py:61-70), optionally
 * followed by vel += acc * kick_factor (nbody.py:70-77). */
int pmwd_gather3(void* stream, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                 const float* f0, const float* f1, const float* f2, float* acc, float* kick_vel,
                 float kick_factor);
int pmwd_gather3_kdk(void* stream, const pmwd_cic_desc* d, const void* pmid, float* disp,
                     const float* f0, const float* f1, const float* f2, float* acc, float* vel,
                     float K2, float K1_next, float D_next);   /* gather3 + pipelined kick/drift, cf. pmwd_force_kdk */
/* alpha = disp cotangent of gravity from the three force meshes and rho_cot
 * (gather.py:106-110 x3 + scatter.py:112-116). */
int pmwd_force_adj_gather(void* stream, const pmwd_cic_desc* d, const void* pmid,
                          const float* disp, const float* f0, const float* f1, const float* f2,
                          const float* rho_cot, const float* pi, float val, float* alpha);
/* k-space kernels on the transposed slab [n0][ny_local][n2/2+1] that holds global rows
 * y0 .. y0+ny_local-1 of axis 1 (layout after the distributed FFT's all-to-all). */
int pmwd_kspace_force_slab(void* stream, const int32_t* shape, int y0, int ny_local,
                           double spacing, float scale, const void* rho_c64, void* const* g_c64);
int pmwd_kspace_force_adj_slab(void* stream, const int32_t* shape, int y0, int ny_local,
                               double spacing, float scale, const void* const* v_c64,
                               void* out_c64);

/* Fused x-pass (csrc/xpass16.cu, csrc/xpass.cu): on data[nx][ny_local][nz/2+1] already transformed
 * over (y, z), FFT along x on chip + laplace/neg_grad algebra + inverse FFT along x in ONE pass.
 * forward: rho2d -> g2d[0..2] (the three force spectra, still to be inverse-transformed over
 * (y, z));  adjoint: v2d[0..2] -> out2d.  `shape` is the GLOBAL mesh shape, nx in
 * {64..2048} powers of two (pmwd_xpass_supported).  Replaces gravity.py:56-64's
 * rfftn-x / laplace / neg_grad / irfftn-x chain.
 * nx in {256, 512, 1024, 2048} with an even ny_local * (nz/2+1) and 16-byte aligned arrays runs
 * the register-resident kernels (16-byte accesses; 2048: two-CTA clusters); anything else the
 * shared-memory radix-4 ones.  pmwd_xpass_last_variant: 0 = none yet, 1 = radix-4, 2 = register
 * kernel, for the last call of this process (test introspection).
 * The output arrays must not overlap the inputs. */
int pmwd_xpass_supported(int nx);
int pmwd_xpass_last_variant(void);
int pmwd_xpass_force(void* stream, const int32_t* shape, int y0, int ny_local, double spacing,
                     float scale, const void* rho2d_c64, void* const* g2d_c64);
int pmwd_xpass_force_adj(void* stream, const int32_t* shape, int y0, int ny_local, double spacing,
                         float scale, const void* const* v2d_c64, void* out2d_c64);

/* ---- fused force: gravity() (pmwd/gravity.py:47-72) -------------------------------- */
/* Workspace (device bytes) needed by pmwd_force / pmwd_force_adj for this geometry. */
size_t pmwd_force_workspace_bytes(const pmwd_cic_desc* d, int adjoint, int mode);
/* acc[N][3] = gravity(ptcl).  If kick_vel != NULL additionally vel += acc * kick_factor
 * (the second half-kick of pmwd/nbody.py:121-140 fused into the gather).
 * 3-D only; requires pmwd_ctx_reserve(ctx, 3, mesh_shape). */
int pmwd_force(pmwd_ctx* ctx, void* stream, const pmwd_cic_desc* d, const void* pmid,
               const float* disp, double Omega_m, float* acc, float* kick_vel,
               float kick_factor, int mode, void* workspace, size_t workspace_bytes,
               const pmwd_sweep* sweep);
/* One whole KDK step (pmwd/nbody.py:121-140, pipelined across steps) in one call: the force at
 * the incoming disp, this step's trailing half-kick (K2) and the NEXT step's leading half-kick
 * and drift (K1_next, D_next) all applied in the gather pass.  disp/vel are updated in place;
 * float32 operation order identical to pmwd_force(+kick) followed by pmwd_kick_drift. */
int pmwd_force_kdk(pmwd_ctx* ctx, void* stream, const pmwd_cic_desc* d, const void* pmid,
                   float* disp, double Omega_m, float* acc, float* vel, float K2, float K1_next,
                   float D_next, int mode, void* workspace, size_t workspace_bytes,
                   const pmwd_sweep* sweep);
/* force_adj (pmwd/nbody.py:108-118): acc = gravity(ptcl) and alpha = VJP_disp(gravity)(pi).
 * The Omega_m cotangent is sum(pi . acc) / Omega_m, which pmwd_kick_adj reduces anyway. */
int pmwd_force_adj(pmwd_ctx* ctx, void* stream, const pmwd_cic_desc* d, const void* pmid,
                   const float* disp, double Omega_m, const float* pi, float* acc,
                   float* alpha, int mode, void* workspace, size_t workspace_bytes,
                   const pmwd_sweep* sweep);

/* ---- Eulerian re-ordering of the particle storage (no reference counterpart) ---------- */
/* The reference stores particles in Lagrangian order (pmwd/particles.py:135-139); the
 * integrator here periodically re-sorts its private copies by mesh cell so that scatter /
 * gather stay coalesced, and restores the reference order on output.
 * perm[i] = storage index of the particle that moves to sorted slot i (stable by cell).
 * ty == 0: key = (x>>1, y>>1, z) (2x2-cell columns along z); ty, bw > 0: the sweep scatter's layout
 * (y / ty, z / bw, x, y % ty, z % bw).  The sorted keys stay in `scratch` (pmwd_cell_sort_sorted_keys). */
size_t pmwd_cell_sort_scratch_bytes(const pmwd_cic_desc* d);
int pmwd_cell_sort_perm(void* stream, const pmwd_cic_desc* d, const void* pmid, const float* disp,
                        uint32_t* perm, void* scratch, size_t scratch_bytes, int ty, int bw);
const uint32_t* pmwd_cell_sort_sorted_keys(const pmwd_cic_desc* d, const void* scratch);
/* Two-source variant for slab runs with Eulerian ownership: sorts the virtual concatenation of
 * A = (pmid, disp)[0, nA) -- rows with ownerA[i] != rank (pmwd_slab_owner) have left for another rank and sort
 * behind everything -- and the arrivals B = (pmidB, dispB)[0, ptcl_num - nA).  d->ptcl_num = nA + nB.  The
 * first ptcl_num - (rows gone) entries of perm are the new storage order (indices >= nA refer to B).
 * vel / velB / pred (vel may be NULL): the keys are computed from the PREDICTED positions disp + vel * pred, so
 * that the order is centred on the force evaluations until the next re-sort (the deposit and the gathers tolerate
 * any order; only their speed depends on it). */
int pmwd_cell_sort_perm2(void* stream, const pmwd_cic_desc* d, const void* pmid, const float* disp, int64_t nA,
                         const uint8_t* ownerA, int rank, const void* pmidB, const float* dispB, uint32_t* perm,
                         void* scratch, size_t scratch_bytes, int ty, int bw, const float* vel, const float* velB,
                         float pred);
/* For each of `narr` row-major arrays (row_bytes[a] bytes per particle, even):
 * inverse == 0: dst[i] = src[perm[i]];  inverse != 0: dst[perm[i]] = src[i]. */
int pmwd_permute_rows(void* stream, int64_t n, const uint32_t* perm, int narr,
                      const void* const* src, void* const* dst, const int32_t* row_bytes,
                      int inverse);
/* dst[i] = (srcA ++ srcB)[perm[i]], i < n: rows [0, nA) of the concatenation are srcA's, the rest srcB's. */
int pmwd_permute_rows2(void* stream, int64_t n, const uint32_t* perm, int narr, const void* const* srcA,
                       int64_t nA, const void* const* srcB, void* const* dst, const int32_t* row_bytes);

/* Slab bookkeeping in one pass (int16 pmid, 3-D): owner[p] = rank whose x-slab holds particle p's base plane
 * (may be NULL), *need = halo planes the slab [x0, x0 + mx) needs for these particles (device int32, reset here). */
int pmwd_slab_owner(void* stream, int64_t n, const void* pmid, const float* disp, double cell_size,
                    int Mx, int nranks, int x0, int mx, uint8_t* owner, int32_t* need);

/* Slab-FFT transpose as one kernel over NVLink peer memory: every rank stores the rows of its
 * local array straight into the peers' receive buffers (peer-mapped device pointers, one per
 * rank, e.g. from torch symmetric memory) at their final position.
 * mode 0: src[mx][My][nzc] -> peer q = y/my: dst_q[rank*mx+ix][iy][nzc];
 * mode 1: src[Mx][my][nzc] -> peer p = x/mx: dst_p[ix][rank*my+iy][nzc].
 * The caller brackets the call with cross-rank barriers. */
int pmwd_transpose_p2p(void* stream, int mode, int nranks, int rank, int mx, int my, int nzc,
                       const void* src_c64, const uint64_t* peer_ptrs);
/* The same transposes on the copy engines (one strided 2-D peer copy per rank, spread over
 * `nstreams` <= 4 internal streams that `stream` forks and joins): no SM pushes data, so the
 * transposes do not slow down the kernels they overlap with. */
int pmwd_transpose_ce(void* stream, int mode, int nranks, int rank, int mx, int my, int nzc,
                      const void* src_c64, const uint64_t* peer_ptrs, int nstreams);
/* Generic form for chunk-pipelined transposes: to every rank q one strided 2-D copy of `height` rows
 * of `width` bytes, src + q * src_peer_stride (pitch spitch) -> peer_ptrs[q] + dst_off (pitch dpitch). */
int pmwd_peer_copy2d(void* stream, int nranks, int rank, size_t width, size_t height, const void* src,
                     size_t src_peer_stride, size_t spitch, const uint64_t* peer_ptrs, size_t dst_off,
                     size_t dpitch, int nstreams);

/* ---- leapfrog updates: pmwd/nbody.py:39-99 ----------------------------------------- */
/* kick (nbody.py:70-77) then drift (nbody.py:39-46) in one pass over n = N*dim floats:
 * if do_kick: vel += acc * K;  if do_drift: disp += vel * D.  In place. */
int pmwd_kick_drift(void* stream, int64_t n, float* disp, float* vel, const float* acc,
                    float K, float D, int do_kick, int do_drift);
/* kick_adj (nbody.py:80-99): vel += acc*K; xi -= alpha*K; sums[0] += sum(pi * acc).
 * drift_adj (nbody.py:49-67): disp += vel*D; pi -= xi*D;  sums[1] += sum(xi * vel).
 * Fused kick_adj-then-drift_adj in one pass; `sums` is a device double[2] accumulated
 * atomically (caller zero-fills).  xi = ptcl_cot.disp, pi = ptcl_cot.vel, alpha = ptcl_cot.acc. */
int pmwd_kick_drift_adj(void* stream, int64_t n, float* disp, float* vel, const float* acc,
                        float* xi, float* pi, const float* alpha, float K, float D,
                        int do_kick, int do_drift, double* sums);
/* kick_adj(K0), then kick_adj(K) + drift_adj(D) in one pass (the trailing half-kick of one adjoint step
 * and the leading half-kick + drift of the next, between which acc and alpha do not change); float32
 * sequence identical to two pmwd_kick_drift_adj calls.  sums_pre[0] and sums[0] += sum(pi . acc),
 * sums[1] += sum(xi . vel). */
int pmwd_kick_kick_drift_adj(void* stream, int64_t n, float* disp, float* vel, const float* acc,
                             float* xi, float* pi, const float* alpha, float K0, float K, float D,
                             double* sums_pre, double* sums);

#ifdef __cplusplus
}
#endif
#endif  /* PMWD_B200_H_ */
