// xla_ffi_shim.cc — XLA FFI handlers wrapping the C ABI of include/cc_b200.h so that the reference can call the
// kernels as `jax.ffi.ffi_call` targets.  Needs xla/ffi/api/ffi.h, which ships inside jaxlib;
// jaxlib is not installable in this image, so the whole file is guarded and is NOT part of the default build.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define CC_HAVE_XLA_FFI 1
#endif
#endif

#ifdef CC_HAVE_XLA_FFI
#include <cstring>

#include "../../include/cc_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error err_of(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, cc_last_error());
}
static CcModule module_of(ffi::Span<const uint8_t> bytes) {
  CcModule m;
  std::memcpy(&m, bytes.data(), sizeof(m));
  return m;
}

static ffi::Error RolloutFwd(cudaStream_t stream, ffi::Buffer<ffi::F32> theta, ffi::Buffer<ffi::F32> u,
                             ffi::ResultBuffer<ffi::F32> y, ffi::ResultBuffer<ffi::F32> tape,
                             ffi::Span<const uint8_t> module) {
  const CcModule m = module_of(module);
  const auto d = u.dimensions();
  return err_of(cc_rollout_fwd(&m, theta.typed_data(), u.typed_data(), nullptr, y->typed_data(), nullptr,
                               tape->typed_data(), 1, d[0], (int32_t)d[1], stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(cc_xla_rollout_fwd, RolloutFwd,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<ffi::Span<const uint8_t>>("module"));
// cc_xla_rollout_bwd / cc_xla_closedloop_fwd / cc_xla_closedloop_bwd follow the same pattern: device buffers in,
// result buffers out, the CcModule structs as byte-span attributes, workspace as an extra result buffer.
#endif  // CC_HAVE_XLA_FFI
