/*
 * cc_b200.h — C ABI of the B200-native neural-ODE rollout path of chain_control.
 *
 * The reference (simon-bachhuber/chain_control) is pure Python/JAX and has no FFI; the seam this
 * library replaces is the *batched* unroll that the train step builds with
 *     eqx.filter_vmap(model.unroll)(inputs)                       cc/train/step_fn.py:124,136
 *     eqx.filter_vmap(controller.unroll(model, merge_x_y, noise))(refss, keys)   step_fn.py:234-236,257-259
 * and its reverse-mode gradient taken by eqx.filter_value_and_grad   step_fn.py:133,249-252.
 * Each entry point below is what a `jax.ffi.ffi_call` (forward) / `jax.custom_vjp` (backward) pair
 * would bind; src/xla_ffi_shim.cc wraps exactly these functions.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / XLA types.  All array pointers are DEVICE pointers,
 *     fp32, C-contiguous, owned by the caller.  Kernels never allocate; scratch comes from `ws`.
 *   - asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - return 0 on success, a negative CcStatus otherwise; cc_last_error() gives a thread-local
 *     message.  An unsupported architecture is an ERROR — there is no fallback path.
 *   - deterministic: fixed reduction order, reruns are bit-identical.
 *   - flat parameter vector `theta` is in the reference's flatten_module order
 *     (cc/core/module_utils.py:20-24): f.layers[0].weight[out,in] row-major, f.layers[0].bias,
 *     f.layers[1]..., then g likewise.
 */
#ifndef CC_B200_H_
#define CC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CC_MAX_LAYERS 4 /* MLP depth <= 3 hidden layers */

enum CcAct { CC_ACT_IDENTITY = 0, CC_ACT_RELU = 1, CC_ACT_TANH = 2 };
enum CcUTransform { CC_UT_IDENTITY = 0, CC_UT_ARCTAN = 1, CC_UT_KTANH = 2 /* k*tanh(u/k) */ };
enum CcIntegrator { CC_INT_RK4 = 0, CC_INT_EE = 1, CC_INT_NONE = 2 }; /* nn_lib/integrate.py:16-28 */

enum CcStatus {
  CC_OK = 0,
  CC_ERR_INVALID_ARG = -1,
  CC_ERR_UNSUPPORTED = -2, /* architecture does not fit the kernels (never falls back) */
  CC_ERR_WORKSPACE = -3,   /* ws_bytes too small: see cc_*_workspace_bytes */
  CC_ERR_CUDA = -4
};

/* eqx.nn.Sequential([Linear, Lambda(act), ..., Linear, Lambda(act_final)])  nn_lib/mlp_network.py:5-35 */
typedef struct CcMlp {
  int32_t n_layers;                 /* depth + 1 */
  int32_t sizes[CC_MAX_LAYERS + 1]; /* [in] + depth*[width] + [out] */
  int32_t act;                      /* CcAct after every layer but the last */
  int32_t act_final;                /* CcAct after the last layer */
  int32_t use_bias;
} CcMlp;

/* One neural-ODE module: NeuralOdeModel (cc/examples/neural_ode_model.py:97-177, compact
 * neural_ode_model_compact_example.py:85-124) or NeuralOdeController (neural_ode_controller.py:99-160,
 * compact neural_ode_controller_compact_example.py:88-181). */
typedef struct CcModule {
  int32_t state_dim;  /* S */
  int32_t input_dim;  /* I: model input u / controller input [ref;obs] */
  int32_t output_dim; /* O */
  CcMlp f;            /* [S (+1 if f_time_variant) + I] -> S ; input order [x ; t ; u~] */
  CcMlp g;            /* [S (+1 if g_time_variant) (+I if g_feedthrough_u)] -> O */
  int32_t integrator;        /* CcIntegrator */
  int32_t f_time_variant;    /* neural_ode_model.py:122-125 */
  int32_t g_time_variant;    /* :129-133 (the PRE-step time is fed) */
  int32_t g_feedthrough_u;   /* controllers only, neural_ode_controller.py:129-134 */
  int32_t readout_pre_step;  /* compact variants: y = g(x) with the state BEFORE the update */
  int32_t u_transform;       /* CcUTransform applied to the module input before f (models) */
  float u_scale;             /* k of k*tanh(u/k) */
  float out_scale;           /* compact controller: sigmoid(alpha) (=0.5) with the zero secondary */
  float dt;                  /* control_timestep */
  float t0;                  /* time component of init_state (0 in the reference; lets step() resume at t) */
} CcModule;

/* ---- introspection ------------------------------------------------------------------------- */
int32_t cc_abi_version(void);
const char* cc_last_error(void);
/* number of trainable scalars P of a module (length of theta / theta_bar) */
int64_t cc_param_count(const CcModule* m);
/* 0 if the kernels support this module (with_grad: also the BPTT kernels), else a CcStatus */
int32_t cc_module_supported(const CcModule* m, int32_t with_grad);

/* which kernel family the plan picks for this call (deterministic, from the architecture and the batch size only):
 * 5 = deep-MLP tcgen05 forward kernels (f with 1..3 hidden layers of width <= 64: one tcgen05.mma batch per layer,
 * chain_control_b200/csrc/cc_mlp_kernels.cuh; open-loop forward only),
 * 4 = blocked linear kernels (linear depth-0 model: m integrator steps per tcgen05.mma batch, chain_control_b200/csrc/cc_lin.h),
 * 3 = tcgen05 reverse pass with a tensor-core weight gradient, 2 = latency kernels (depth-0 model, small batch: one warp
 * per trajectory), 1 = tcgen05 kernels (depth-0 model: stage evaluations as tcgen05.mma batches, 3xTF32), 0 = FFMA kernels,
 * < 0 = a CcStatus (unsupported / invalid).  ctrl == NULL: open loop. */
int32_t cc_kernel_family(const CcModule* ctrl, const CcModule* model, int32_t with_grad);
/* same for a given batch size (the latency kernels are picked for small batches only; cc_kernel_family assumes 65,536) */
int32_t cc_kernel_family_at(const CcModule* ctrl, const CcModule* model, int32_t with_grad, int64_t B);

/* ---- K1/K2: open-loop batched unroll and its VJP  (replaces step_fn.py:136 + abstract.py:54-70) ---
 * u      [B,T,I]      inputs
 * x0     [B,S] or NULL (zeros = reset(), neural_ode_model.py:103-104)
 * y      [B,T+1,O]    outputs incl. y0 (include_y0=True)
 * tape   forward->backward residuals (module states every `ckpt_every` steps), opaque layout,
 *        cc_rollout_tape_bytes() bytes; NULL for inference.
 * x_final [B,S] or NULL: state after the last step (lets step() be served by T=1 calls).
 */
size_t cc_rollout_tape_bytes(const CcModule* m, int64_t B, int32_t T, int32_t ckpt_every);
/* scratch of the forward call (0 for most kernel families; ws may then be NULL) / of the reverse call */
size_t cc_rollout_fwd_workspace_bytes(const CcModule* m, int64_t B, int32_t T);
size_t cc_rollout_workspace_bytes(const CcModule* m, int64_t B, int32_t T, int32_t ckpt_every);
int32_t cc_rollout_fwd(const CcModule* m, const float* theta, const float* u, const float* x0,
                       float* y, float* x_final, void* tape, int32_t ckpt_every, int64_t B,
                       int32_t T, void* ws, size_t ws_bytes, void* stream);
/* ybar [B,T+1,O] -> theta_bar [P] (overwritten), ubar [B,T,I] or NULL.  Needs the tape written by
 * cc_rollout_fwd with the same (theta,u,x0,B,T,ckpt_every). */
int32_t cc_rollout_bwd(const CcModule* m, const float* theta, const float* u, const float* x0,
                       const void* tape, int32_t ckpt_every, const float* ybar, float* theta_bar,
                       float* ubar, int64_t B, int32_t T, void* ws, size_t ws_bytes, void* stream);

/* ---- K3/K4: closed-loop batched unroll and its VJP (replaces step_fn.py:257-259 + abstract.py:97-127)
 * refs   [B,Tr,R]     reference series; controller input is [ref ; y_prev] (merge_x_y, step_fn.py:187-192)
 * noise  [B,Tr,O] or NULL: additive observation noise, already scaled (abstract.py:77-79,112-114)
 * yhat   [B,Tr,O]     closed-loop outputs (what unroll_closed_loop returns, abstract.py:125)
 * u_out  [B,Tr,I_m] or NULL: the control series (discarded by the reference, kept for diagnostics)
 * theta_c_bar [P_c]: gradient w.r.t. the controller only; the model is frozen (step_fn.py:249-252).
 * tape   opaque, cc_closedloop_tape_bytes() bytes: cc_closedloop_bwd needs the tape written by cc_closedloop_fwd with the same
 *        (theta_c, theta_m, refs, noise, B, Tr, ckpt_every) -- besides the carried states it may hold operands prepared from
 *        the parameters (blocked linear family: the fp64-prepared block matrices, so that a train step prepares them once).
 */
size_t cc_closedloop_tape_bytes(const CcModule* ctrl, const CcModule* model, int64_t B, int32_t Tr,
                                int32_t ckpt_every);
size_t cc_closedloop_fwd_workspace_bytes(const CcModule* ctrl, const CcModule* model, int64_t B,
                                         int32_t Tr);
size_t cc_closedloop_workspace_bytes(const CcModule* ctrl, const CcModule* model, int64_t B,
                                     int32_t Tr, int32_t ckpt_every);
int32_t cc_closedloop_fwd(const CcModule* ctrl, const CcModule* model, const float* theta_c,
                          const float* theta_m, const float* refs, const float* noise, float* yhat,
                          float* u_out, void* tape, int32_t ckpt_every, int64_t B, int32_t Tr,
                          void* ws, size_t ws_bytes, void* stream);
int32_t cc_closedloop_bwd(const CcModule* ctrl, const CcModule* model, const float* theta_c,
                          const float* theta_m, const float* refs, const float* noise,
                          const void* tape, int32_t ckpt_every, const float* ybar,
                          float* theta_c_bar, int64_t B, int32_t Tr, void* ws, size_t ws_bytes,
                          void* stream);

/* ---- K6: default loss (cc/utils/utils.py:53-55) fused with its cotangent ----------------------
 * loss_partial[0] = sum((yhat-y)^2) * inv_n  ;  ybar = 2*(yhat-y)*inv_n   (inv_n = 1/global count) */
int32_t cc_mse_value_and_cotangent(const float* y, const float* yhat, float* ybar,
                                   float* loss_partial, int64_t n, float inv_n, void* stream);

/* ---- tail of the train step on the device (cc/train/step_fn.py:101-114,166-169,307-308; optax semantics) ----
 * grad' = grad_scale * grad + l1 * sign(theta) + l2 * theta / ||theta||_2   (regularisers on the flat vector, utils.py:38-45)
 * -> optax.clip_by_global_norm(clip_global_norm) (<= 0: off) -> optax.clip(clip_value) (<= 0: off) -> optax.adam(lr, b1, b2, eps)
 * theta, mu, nu [P] are updated in place; count[1] (device int32) is the step counter of the bias correction (incremented);
 * out3 (device, optional) = {l1 norm, l2 norm of theta before the update, ||grad'||_2 before clipping}.
 * One launch, no host round trip: the whole step (rollout, loss, BPTT, update) can be captured in a CUDA graph. */
int32_t cc_opt_step(float* theta, float* mu, float* nu, const float* grad, int32_t* count, float* out3, int64_t P,
                    float lr, float b1, float b2, float eps, float clip_global_norm, float clip_value,
                    float l1_prefactor, float l2_prefactor, float grad_scale, void* stream);

/* ---- data source: B independent two-segment-chain environments -------------------------------------
 * Replaces, for the `two_segments*` ids of cc/env/sample_envs.py:13-60, the per-episode Python loop
 * cc/env/collect/collect.py:78-98 -> control.Environment.step -> MuJoCo mj_step over the model cc/env/envs/two_segments.xml:6-14 +
 * cc/env/envs/two_segments.py:61-104 (cart on a slide joint, two capsule poles on hinges, motor gear 5; RK4 at 1 ms, no
 * gravity / contacts) and the task :119-136,242-263 (ctrl = arctan(action), observation = world x of segment_end, float32:
 * cc/env/make_env.py:43-46).  Pinned by the reference's golden vectors cc/env/test_make_env.py:146-335.
 *   params  [n_params][4] fp64 = {slider damping, slider stiffness, hinge damping, hinge stiffness}; n_params = 1 (shared) or B
 *   actions [B, T] fp32 (arctan evaluated in fp32, as numpy does for float32 actions) or actions_f64 [B, T] (Python-float
 *           actions: arctan in fp64) -- exactly one non-null when T > 0
 *   obs     [B, T + 1] fp32: obs[:, 0] observes the incoming state, obs[:, k + 1] the state after control step k
 *   state   [B, 6] fp64 (qpos[3], qvel[3]) read and written back, or NULL = start from reset (all zero), final state dropped
 *   n_sub_steps = control_timestep / physics_dt (10 for the reference's 0.01 s / 1 ms) */
int32_t cc_env_two_segments_rollout(const double* params, int64_t n_params, const float* actions, const double* actions_f64,
                                    float* obs, double* state, int64_t B, int32_t T, int32_t n_sub_steps, double physics_dt,
                                    void* stream);

/* ---- measurement helper: register-resident FFMA chain, returns nothing; time it with events ---- */
int32_t cc_fp32_peak_probe(float* sink, int32_t iters, void* stream);
/* same, but the SGEMM register-tile pattern acc[i][j] += x[i]*y[j] (three register operands per FFMA):
 * the practical CUDA-core ceiling of a register-tiled kernel; 4*32*4 FMAs per thread per iteration */
int32_t cc_fp32_tile_probe(float* sink, const float* in, int32_t iters, void* stream);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
int64_t cc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CC_B200_H_ */
