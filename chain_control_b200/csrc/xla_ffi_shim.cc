// xla_ffi_shim.cc — XLA FFI handlers wrapping the C ABI of include/cc_b200.h so that the reference can call the kernels as
// `jax.ffi.ffi_call` targets inside a `jax.custom_vjp` (INTEGRATION.md shows the Python side):
//
//   cc_xla_rollout_fwd / _bwd      eqx.filter_vmap(model.unroll)(inputs) and its VJP           cc/train/step_fn.py:124,133-136
//   cc_xla_closedloop_fwd / _bwd   eqx.filter_vmap(controller.unroll(model, merge_x_y, noise))  cc/train/step_fn.py:234-236,249-259
//   cc_xla_mse                     default loss value + cotangent                               cc/utils/utils.py:53-55
//   cc_xla_two_segments_rollout    B two-segment environments x T control steps (data source)  cc/env/collect/collect.py:78-98
//
// Conventions: operands / results are XLA device buffers (F32); the CcModule structs travel as uint8 array attributes
// (numpy.frombuffer of the ctypes struct); scratch and the tape are extra RESULT buffers sized on the host with
// cc_*_workspace_bytes / cc_*_tape_bytes (plain host functions); optional operands (x0, noise) are passed as zero-size
// buffers when absent; `ckpt_every` is an i32 attribute.  A non-zero CcStatus becomes an xla::ffi::Error carrying
// cc_last_error().
//
// Needs xla/ffi/api/ffi.h, which ships inside jaxlib (not installable in this image): the file is not part of the default
// build.  tests/test_ffi_shim.py compiles it against a minimal stand-in (tests/xla_stub) whose Bind().To() checks every
// handler signature against its binding, so it cannot rot.
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
using BufF = ffi::Buffer<ffi::F32>;
using ResF = ffi::ResultBuffer<ffi::F32>;
using Bytes = ffi::Span<const uint8_t>;

static ffi::Error err_of(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, cc_last_error());
}
static bool module_of(Bytes bytes, CcModule* m) {
  if (bytes.size() != sizeof(CcModule)) return false;
  std::memcpy(m, bytes.data(), sizeof(CcModule));
  return true;
}
static ffi::Error bad_module() { return ffi::Error(ffi::ErrorCode::kInvalidArgument, "module attribute is not a CcModule (include/cc_b200.h)"); }
// optional operand: a zero-size buffer means "absent"
static const float* opt(const BufF& b) { return b.element_count() == 0 ? nullptr : b.typed_data(); }
static float* opt(ResF& b) { return b->element_count() == 0 ? nullptr : b->typed_data(); }

// ---- K1: y[B,T+1,O], tape, ws = rollout_fwd(theta[P], u[B,T,I], x0[B,S] | [0]) ----
static ffi::Error RolloutFwd(cudaStream_t stream, BufF theta, BufF u, BufF x0, ResF y, ResF tape, ResF ws, Bytes module, int32_t ckpt_every) {
  CcModule m;
  if (!module_of(module, &m)) return bad_module();
  const auto d = u.dimensions();
  return err_of(cc_rollout_fwd(&m, theta.typed_data(), u.typed_data(), opt(x0), y->typed_data(), nullptr, opt(tape), ckpt_every, d[0],
                               (int32_t)d[1], opt(ws), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(cc_xla_rollout_fwd, RolloutFwd,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<BufF>()   // theta
                                  .Arg<BufF>()   // u
                                  .Arg<BufF>()   // x0 (zero-size: reset())
                                  .Ret<BufF>()   // y
                                  .Ret<BufF>()   // tape (zero-size: inference)
                                  .Ret<BufF>()   // workspace
                                  .Attr<Bytes>("module")
                                  .Attr<int32_t>("ckpt_every"));

// ---- K2: theta_bar[P], ubar[B,T,I] | [0], ws = rollout_bwd(theta, u, x0, tape, ybar[B,T+1,O]) ----
static ffi::Error RolloutBwd(cudaStream_t stream, BufF theta, BufF u, BufF x0, BufF tape, BufF ybar, ResF theta_bar, ResF ubar, ResF ws,
                             Bytes module, int32_t ckpt_every) {
  CcModule m;
  if (!module_of(module, &m)) return bad_module();
  const auto d = u.dimensions();
  return err_of(cc_rollout_bwd(&m, theta.typed_data(), u.typed_data(), opt(x0), tape.typed_data(), ckpt_every, ybar.typed_data(),
                               theta_bar->typed_data(), opt(ubar), d[0], (int32_t)d[1], ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(cc_xla_rollout_bwd, RolloutBwd,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<BufF>()   // theta
                                  .Arg<BufF>()   // u
                                  .Arg<BufF>()   // x0
                                  .Arg<BufF>()   // tape
                                  .Arg<BufF>()   // ybar
                                  .Ret<BufF>()   // theta_bar
                                  .Ret<BufF>()   // ubar (zero-size: not requested)
                                  .Ret<BufF>()   // workspace
                                  .Attr<Bytes>("module")
                                  .Attr<int32_t>("ckpt_every"));

// ---- K3: yhat[B,Tr,O], tape, ws = closedloop_fwd(theta_c, theta_m, refs[B,Tr,R], noise[B,Tr,O] | [0]) ----
static ffi::Error ClosedLoopFwd(cudaStream_t stream, BufF theta_c, BufF theta_m, BufF refs, BufF noise, ResF yhat, ResF tape, ResF ws,
                                Bytes ctrl, Bytes model, int32_t ckpt_every) {
  CcModule c, m;
  if (!module_of(ctrl, &c) || !module_of(model, &m)) return bad_module();
  const auto d = refs.dimensions();
  return err_of(cc_closedloop_fwd(&c, &m, theta_c.typed_data(), theta_m.typed_data(), refs.typed_data(), opt(noise), yhat->typed_data(), nullptr,
                                  opt(tape), ckpt_every, d[0], (int32_t)d[1], opt(ws), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(cc_xla_closedloop_fwd, ClosedLoopFwd,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<BufF>()   // theta_c
                                  .Arg<BufF>()   // theta_m
                                  .Arg<BufF>()   // refs
                                  .Arg<BufF>()   // noise (zero-size: none)
                                  .Ret<BufF>()   // yhat
                                  .Ret<BufF>()   // tape
                                  .Ret<BufF>()   // workspace
                                  .Attr<Bytes>("controller")
                                  .Attr<Bytes>("model")
                                  .Attr<int32_t>("ckpt_every"));

// ---- K4: theta_c_bar[P_c], ws = closedloop_bwd(theta_c, theta_m, refs, noise, tape, ybar[B,Tr,O]) ----
static ffi::Error ClosedLoopBwd(cudaStream_t stream, BufF theta_c, BufF theta_m, BufF refs, BufF noise, BufF tape, BufF ybar, ResF theta_c_bar,
                                ResF ws, Bytes ctrl, Bytes model, int32_t ckpt_every) {
  CcModule c, m;
  if (!module_of(ctrl, &c) || !module_of(model, &m)) return bad_module();
  const auto d = refs.dimensions();
  return err_of(cc_closedloop_bwd(&c, &m, theta_c.typed_data(), theta_m.typed_data(), refs.typed_data(), opt(noise), tape.typed_data(), ckpt_every,
                                  ybar.typed_data(), theta_c_bar->typed_data(), d[0], (int32_t)d[1], ws->typed_data(), ws->size_bytes(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(cc_xla_closedloop_bwd, ClosedLoopBwd,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<BufF>()   // theta_c
                                  .Arg<BufF>()   // theta_m
                                  .Arg<BufF>()   // refs
                                  .Arg<BufF>()   // noise
                                  .Arg<BufF>()   // tape
                                  .Arg<BufF>()   // ybar
                                  .Ret<BufF>()   // theta_c_bar
                                  .Ret<BufF>()   // workspace
                                  .Attr<Bytes>("controller")
                                  .Attr<Bytes>("model")
                                  .Attr<int32_t>("ckpt_every"));

// ---- K6: ybar, loss_partial[1 + 296] = mse(y, yhat); inv_n = 1 / global element count ----
static ffi::Error Mse(cudaStream_t stream, BufF y, BufF yhat, ResF ybar, ResF loss_partial, float inv_n) {
  if (loss_partial->element_count() < 297) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "loss_partial must hold 297 floats");
  return err_of(cc_mse_value_and_cotangent(y.typed_data(), yhat.typed_data(), ybar->typed_data(), loss_partial->typed_data(),
                                           (int64_t)y.element_count(), inv_n, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(cc_xla_mse, Mse,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<BufF>()   // y
                                  .Arg<BufF>()   // yhat
                                  .Ret<BufF>()   // ybar
                                  .Ret<BufF>()   // loss_partial
                                  .Attr<float>("inv_n"));

// ---- data source: obs[B,T+1] = two_segments_rollout(params[n,4] f64, actions[B,T] f32): whole episodes from reset ----
static ffi::Error TwoSegmentsRollout(cudaStream_t stream, ffi::Buffer<ffi::F64> params, BufF actions, ResF obs, int32_t n_sub_steps, float physics_dt) {
  const auto d = actions.dimensions();
  const auto pd = params.dimensions();
  return err_of(cc_env_two_segments_rollout(params.typed_data(), pd[0], d[1] ? actions.typed_data() : nullptr, nullptr, obs->typed_data(), nullptr,
                                            d[0], (int32_t)d[1], n_sub_steps, (double)physics_dt, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(cc_xla_two_segments_rollout, TwoSegmentsRollout,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // params [n, 4]
                                  .Arg<BufF>()                    // actions [B, T]
                                  .Ret<BufF>()                    // obs [B, T + 1]
                                  .Attr<int32_t>("n_sub_steps")
                                  .Attr<float>("physics_dt"));
#endif  // CC_HAVE_XLA_FFI
