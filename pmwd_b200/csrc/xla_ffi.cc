// XLA FFI registration of the C ABI (include/pmwd_b200.h) for JAX: the binding a pmwd maintainer
// adds so that pmwd/{scatter,gather,gravity,nbody}.py call the sm_100a kernels through
// jax.ffi.ffi_call inside their existing custom_vjp wrappers.
//
// NOT COMPILED IN THIS IMAGE: jaxlib (xla/ffi/api/ffi.h) is not installed here,
// so this translation unit is excluded from pmwd_b200/build.py (which builds *.cu only) and is
// untested.  Build where JAX is available:
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I/usr/local/cuda/include -Iinclude pmwd_b200/csrc/xla_ffi.cc \
//       -Lpmwd_b200 -lpmwd_b200 -o pmwd_b200/libpmwd_b200_xla.so
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime_api.h>

#include "pmwd_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Check(int rc) {
  if (rc == 0) return ffi::Error::Success();
  char msg[512];
  pmwd_last_error(msg, sizeof msg);
  return ffi::Error(ffi::ErrorCode::kInternal, msg);
}

pmwd_cic_desc FastDesc(int64_t n, ffi::Span<const int32_t> mesh, double cell_size) {
  pmwd_cic_desc d{};
  d.dim = 3;
  d.pmid_bytes = 2;
  d.ptcl_num = n;
  d.nchan = 1;
  d.cell_size = cell_size;
  for (int a = 0; a < 3; ++a) d.wrap_shape[a] = d.mesh_shape[a] = mesh[a];
  return d;
}

// gravity(): pmwd/gravity.py:47-72
ffi::Error ForceImpl(cudaStream_t stream, ffi::Buffer<ffi::S16> pmid, ffi::Buffer<ffi::F32> disp,
                     ffi::Buffer<ffi::U8> workspace, ffi::ResultBuffer<ffi::F32> acc, double Omega_m,
                     double cell_size, int64_t ctx, ffi::Span<const int32_t> mesh) {
  pmwd_cic_desc d = FastDesc(pmid.dimensions()[0], mesh, cell_size);
  return Check(pmwd_force(reinterpret_cast<pmwd_ctx*>(ctx), stream, &d, pmid.typed_data(),
                          disp.typed_data(), Omega_m, acc->typed_data(), nullptr, 0.f,
                          PMWD_SCATTER_ATOMIC, workspace.typed_data(), workspace.size_bytes(),
                          /*sweep=*/nullptr));
}

// force_adj(): pmwd/nbody.py:108-118 (gravity and its VJP w.r.t. disp for the cotangent pi)
ffi::Error ForceAdjImpl(cudaStream_t stream, ffi::Buffer<ffi::S16> pmid, ffi::Buffer<ffi::F32> disp,
                        ffi::Buffer<ffi::F32> pi, ffi::Buffer<ffi::U8> workspace,
                        ffi::ResultBuffer<ffi::F32> acc, ffi::ResultBuffer<ffi::F32> alpha,
                        double Omega_m, double cell_size, int64_t ctx, ffi::Span<const int32_t> mesh) {
  pmwd_cic_desc d = FastDesc(pmid.dimensions()[0], mesh, cell_size);
  return Check(pmwd_force_adj(reinterpret_cast<pmwd_ctx*>(ctx), stream, &d, pmid.typed_data(),
                              disp.typed_data(), Omega_m, pi.typed_data(), acc->typed_data(),
                              alpha->typed_data(), PMWD_SCATTER_ATOMIC, workspace.typed_data(),
                              workspace.size_bytes(), /*sweep=*/nullptr));
}

// kick + drift: pmwd/nbody.py:39-46,70-77 (functional: outputs alias-free copies made by XLA)
ffi::Error KickDriftImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> disp, ffi::Buffer<ffi::F32> vel,
                         ffi::Buffer<ffi::F32> acc, ffi::ResultBuffer<ffi::F32> disp_out,
                         ffi::ResultBuffer<ffi::F32> vel_out, float K, float D, bool do_kick,
                         bool do_drift) {
  const size_t bytes = disp.size_bytes();
  if (cudaMemcpyAsync(disp_out->typed_data(), disp.typed_data(), bytes, cudaMemcpyDeviceToDevice, stream) !=
          cudaSuccess ||
      cudaMemcpyAsync(vel_out->typed_data(), vel.typed_data(), bytes, cudaMemcpyDeviceToDevice, stream) !=
          cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemcpyAsync failed");
  return Check(pmwd_kick_drift(stream, (int64_t)disp.element_count(), disp_out->typed_data(),
                               vel_out->typed_data(), acc.typed_data(), K, D, do_kick, do_drift));
}

// _scatter / _gather fast path: pmwd/scatter.py:33-83, pmwd/gather.py:33-77 (scalar field, default
// geometry; the general entry points pmwd_scatter / pmwd_gather take the full pmwd_cic_desc)
ffi::Error ScatterImpl(cudaStream_t stream, ffi::Buffer<ffi::S16> pmid, ffi::Buffer<ffi::F32> disp,
                       ffi::Buffer<ffi::F32> mesh_in, ffi::ResultBuffer<ffi::F32> mesh, float val,
                       double cell_size, ffi::Span<const int32_t> shape) {
  pmwd_cic_desc d = FastDesc(pmid.dimensions()[0], shape, cell_size);
  if (cudaMemcpyAsync(mesh->typed_data(), mesh_in.typed_data(), mesh_in.size_bytes(),
                      cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "cudaMemcpyAsync failed");
  return Check(pmwd_scatter(stream, &d, pmid.typed_data(), disp.typed_data(), nullptr, val,
                            mesh->typed_data(), PMWD_SCATTER_ATOMIC, nullptr, 0));
}

ffi::Error GatherImpl(cudaStream_t stream, ffi::Buffer<ffi::S16> pmid, ffi::Buffer<ffi::F32> disp,
                      ffi::Buffer<ffi::F32> mesh, ffi::ResultBuffer<ffi::F32> out, float val,
                      double cell_size, ffi::Span<const int32_t> shape) {
  pmwd_cic_desc d = FastDesc(pmid.dimensions()[0], shape, cell_size);
  return Check(pmwd_gather(stream, &d, pmid.typed_data(), disp.typed_data(), mesh.typed_data(), nullptr,
                           val, out->typed_data()));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(PmwdForce, ForceImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<double>("Omega_m")
                                  .Attr<double>("cell_size")
                                  .Attr<int64_t>("ctx")
                                  .Attr<ffi::Span<const int32_t>>("mesh"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(PmwdForceAdj, ForceAdjImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<double>("Omega_m")
                                  .Attr<double>("cell_size")
                                  .Attr<int64_t>("ctx")
                                  .Attr<ffi::Span<const int32_t>>("mesh"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(PmwdKickDrift, KickDriftImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("K")
                                  .Attr<float>("D")
                                  .Attr<bool>("do_kick")
                                  .Attr<bool>("do_drift"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(PmwdScatter, ScatterImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("val")
                                  .Attr<double>("cell_size")
                                  .Attr<ffi::Span<const int32_t>>("shape"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(PmwdGather, GatherImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("val")
                                  .Attr<double>("cell_size")
                                  .Attr<ffi::Span<const int32_t>>("shape"));
#endif  // __has_include("xla/ffi/api/ffi.h")
