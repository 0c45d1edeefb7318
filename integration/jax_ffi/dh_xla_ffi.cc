// XLA FFI custom-call shim over the C ABI of libdeephall_b200.so (include/deephall_b200.h).
//
// STATUS: source only.  This image has no JAX / jaxlib and therefore no `xla/ffi/api/ffi.h`; the file is NOT compiled
// or tested here (the layer that is built and tested is the C ABI + ctypes binding, deephall_b200/_native.py).  It is the
// maintainer-side half of INTEGRATION.md section B: on a machine with jax >= 0.4.35,
//
//   g++ -O2 -shared -fPIC -std=c++17 dh_xla_ffi.cc -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I../../include -L../../deephall_b200 -ldeephall_b200 -lcudart -o libdh_xla_ffi.so
//
// and deephall/b200_ffi.py registers the four handlers with
//   jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA").
//
// Every handler is a thin argument adapter: buffers are device pointers, the stream comes from XLA, the plan handle is
// an int64 attribute (the dh_plan* created once per (device, configuration) by the Python side through ctypes), and the
// workspace is the LAST result buffer (sized with dh_workspace_bytes on the Python side, so XLA owns the memory and
// the library never allocates -- the ownership rule of SURVEY 8b).
//
// Traced values are OPERANDS, not attributes: the proposal width lives in the reference's CheckpointState and changes by
// x1.1 (mcmc.py:181-185), the key is split every step (train.py:127) -- as attributes they would force a recompile per
// value.  The sweep handler therefore takes `width` (1 x f32) and `key` (2 x u64: Philox seed, offset) as device buffers and
// calls dh_mcmc_sweep_dev, whose kernels read them on the device.  Static attributes: plan handle, steps, subsequence0.
// Plans used through this shim keep the library default auto_prepare = 1 (the split / folded weight copies are rebuilt
// from `params` on every call): XLA recycles donated buffer addresses, so an address-keyed weight cache would go stale.
#include <cstdint>
#include <string>

#include <cuda_runtime_api.h>

#include "deephall_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

inline dh_plan* plan_of(int64_t handle) { return reinterpret_cast<dh_plan*>(static_cast<intptr_t>(handle)); }

inline ffi::Error status(int rc, const char* what) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + " failed with code " + std::to_string(rc));
}

// model.apply(params, x) for a batch: x (B, N, 2) f32 -> (B, 2) f32 = (log|psi|, phase)      networks/psiformer.py:72-76
ffi::Error LogpsiImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> x,
                      ffi::ResultBuffer<ffi::F32> out, ffi::ResultBuffer<ffi::U8> ws, int64_t plan) {
  const int64_t B = x.dimensions()[0];
  return status(dh_logpsi(plan_of(plan), params.typed_data(), x.typed_data(), B, out->typed_data(), ws->typed_data(),
                          ws->element_count(), stream),
                "dh_logpsi");
}

// local_energy(f, system)(params, x): E_L, kinetic (B, 2) f32; potential, L_z, L_z^2, L^2 (B) f32    hamiltonian.py:175-212
ffi::Error LocalEnergyImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> x,
                           ffi::ResultBuffer<ffi::F32> el, ffi::ResultBuffer<ffi::F32> kinetic,
                           ffi::ResultBuffer<ffi::F32> potential, ffi::ResultBuffer<ffi::F32> lz,
                           ffi::ResultBuffer<ffi::F32> lz2, ffi::ResultBuffer<ffi::F32> l2, ffi::ResultBuffer<ffi::U8> ws,
                           int64_t plan) {
  const int64_t B = x.dimensions()[0];
  return status(dh_local_energy(plan_of(plan), params.typed_data(), x.typed_data(), B, el->typed_data(),
                                kinetic->typed_data(), potential->typed_data(), lz->typed_data(), lz2->typed_data(),
                                l2->typed_data(), nullptr, ws->typed_data(), ws->element_count(), stream),
                "dh_local_energy");
}

// make_mcmc_step(...)(params, data, key, width): `data` is donated (input_output_aliases={1: 0} on the Python side, so
// x_out aliases x_in, train.py:75); naccept is one int64 on the device (pmove = naccept / (steps * B), mcmc.py:146)
ffi::Error McmcSweepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> x_in,
                         ffi::Buffer<ffi::F32> width, ffi::Buffer<ffi::U64> key, ffi::ResultBuffer<ffi::F32> x_out,
                         ffi::ResultBuffer<ffi::S64> naccept, ffi::ResultBuffer<ffi::U8> ws, int64_t plan, int32_t steps,
                         int64_t subsequence0) {
  const int64_t B = x_in.dimensions()[0];
  if (x_out->typed_data() != x_in.typed_data()) {  // not aliased: keep functional semantics
    cudaError_t e = cudaMemcpyAsync(x_out->typed_data(), x_in.typed_data(), x_in.size_bytes(), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return status(static_cast<int>(e), "cudaMemcpyAsync");
  }
  return status(dh_mcmc_sweep_dev(plan_of(plan), params.typed_data(), x_out->typed_data(), B, steps, width.typed_data(),
                                  key.typed_data(), static_cast<uint64_t>(subsequence0),
                                  reinterpret_cast<long long*>(naccept->typed_data()), nullptr, ws->typed_data(),
                                  ws->element_count(), stream),
                "dh_mcmc_sweep_dev");
}

// VJP of b -> (Re, Im) log psi_b with per-walker cotangents cot (B, 2): grad (P) f32             loss.py:53-64,96-106
ffi::Error LogpsiVjpImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> x,
                         ffi::Buffer<ffi::F32> cot, ffi::ResultBuffer<ffi::F32> grad, ffi::ResultBuffer<ffi::U8> ws,
                         int64_t plan) {
  const int64_t B = x.dimensions()[0];
  return status(dh_logpsi_vjp(plan_of(plan), params.typed_data(), x.typed_data(), B, cot.typed_data(), grad->typed_data(),
                              nullptr, ws->typed_data(), ws->element_count(), stream),
                "dh_logpsi_vjp");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(dh_logpsi_ffi, LogpsiImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params (P)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x (B, N, 2)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // log psi (B, 2)
                                  .Ret<ffi::Buffer<ffi::U8>>()    // workspace
                                  .Attr<int64_t>("plan"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(dh_local_energy_ffi, LocalEnergyImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()   // E_L (B, 2)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // kinetic (B, 2)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // potential (B)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // L_z (B)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // L_z^2 (B)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // L^2 (B)
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("plan"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(dh_mcmc_sweep_ffi, McmcSweepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // params (P)
                                  .Arg<ffi::Buffer<ffi::F32>>()   // walkers (B, N, 2), donated
                                  .Arg<ffi::Buffer<ffi::F32>>()   // width (1): traced
                                  .Arg<ffi::Buffer<ffi::U64>>()   // key (2) = (Philox seed, offset): traced
                                  .Ret<ffi::Buffer<ffi::F32>>()   // walkers (B, N, 2), aliased to the input
                                  .Ret<ffi::Buffer<ffi::S64>>()   // accepted moves (1)
                                  .Ret<ffi::Buffer<ffi::U8>>()    // workspace
                                  .Attr<int64_t>("plan")
                                  .Attr<int32_t>("steps")
                                  .Attr<int64_t>("subsequence0"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(dh_logpsi_vjp_ffi, LogpsiVjpImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()   // cotangents (B, 2)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // gradient (P)
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("plan"));
