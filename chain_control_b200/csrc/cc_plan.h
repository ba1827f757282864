// cc_plan.h — host-built launch plan shared by the forward and backward kernels.
//
// Execution model.  A WARP is an autonomous solver for 32 trajectories for the whole
// time loop; a CTA is just `wpc` such warps sharing one staged copy of the weights.  After the
// prologue there is no CTA-level barrier: a warp only ever reads activations it wrote itself, so
// __syncwarp() is the only synchronisation in the time loop and warps drift freely (one warp's
// epilogue / shuffle phase overlaps another warp's FFMA stream).
//
// Lane mapping: lane = cg * 8 + rg, 4 column groups (cg) x 8 row groups (rg).  A thread owns the
// TM = 4 trajectories 4rg..4rg+3 of its warp and, for a layer with N outputs, the cpt = ceil(N/4)
// consecutive outputs cg*cpt .. cg*cpt+cpt-1  ->  register tile [4][TN], TN >= cpt.
//
// Shared-memory layout.  Every activation buffer of a warp is feature-major, buf[row][32]
// (32 = the warp's trajectories): a thread's 4 trajectories are one float4, a quarter-warp
// (fixed cg, rg = 0..7) reads/writes 128 contiguous bytes -> every LDS.128/STS.128 is conflict free.
// A module ("node") keeps two stage buffers X / Z whose rows are
//     [ x (S) ; t (1, iff the module tracks time) ; v (I, transformed input) ; 1 (bias row) ]
// = the reference's batch_concat order of f's and g's inputs (neural_ode_model.py:122-133,
// neural_ode_controller.py:121-134).  Weights are staged once per CTA as Wt[row][4*GS]
// (contraction-major; column n at (n / cpt) * GS + n % cpt; GS chosen so that the 4 column groups of
// a warp hit disjoint banks) with zero rows where an MLP does not consume t / v and the bias as the
// row that multiplies the constant-1 row.  Each layer evaluation is then a register-tiled SGEMM
//     C[32 x N] = A[K x 32]^T . Wt[K x N]
// RK4 ping-pongs X -> Z -> X -> Z -> X with the state tile and the stage sum in registers.
#pragma once
#include <stdint.h>

#include "../../include/cc_b200.h"

#define CC_TM 4
#define CC_LD 32         /* floats per activation-buffer row = trajectories per warp */
#define CC_MAX_SKINNY 4  /* layers with N <= this use the split-K + shuffle path */
#define CC_TINY_S 8      /* per-lane ("tiny") node path: state_dim <= 8, rows <= 12, outputs <= 4 */
#define CC_TINY_K 12
#define CC_TINY_O 4

enum CcCtrlMode { CC_CTRL_NONE = 0, CC_CTRL_TINY = 1, CC_CTRL_TILE = 2 };

struct CcLayerDev {
  int K;       // contraction rows read from the source buffer (incl. the constant-1 row)
  int N;       // outputs
  int cpt;     // output columns per thread = ceil(N / 4)
  int GS;      // column-group stride inside a Wt row (floats)
  int ldw;     // Wt row stride = 4 * GS
  int Kin;     // input rows whose cotangent goes through the tile path (state rows / hidden units)
  int kcpt;    // columns per thread of the input-gradient tile = cpt of the source buffer
  int GSk;     // group stride / row stride of the row-major copy W[n][...] used by the input gradient
  int ldk;
  int act;     // activation applied to this layer's output
  int in_log;  // logical input width = row stride of the weight in theta
  int w_off;   // offset of weight[N][in_log] in theta
  int b_off;   // offset of bias[N] in theta, or -1
  int first;   // first layer of its MLP: physical->logical row map goes through the node layout
  int use_t;   // (first layers) this MLP consumes the time row
  int use_v;   // (first layers) this MLP consumes the input rows
  int s_wt;    // CTA-shared: Wt[K][ldw]
  int s_w;     // CTA-shared: W[N][ldk] (backward only, else -1)
  int s_wv;    // CTA-shared: WV[I][N] weights of the input rows (backward, first layers), else -1
  int w_gw;    // per-warp: Wbar[N][K] accumulator (backward & trainable, tile path), else -1
  int w_out;   // per-warp: hidden output buffer [(N+1)][32] (hidden layers), else -1
  int g_off;   // offset of this layer's [N][K] block inside the node's gradient record
};

struct CcMlpDev {
  int L;
  CcLayerDev layer[CC_MAX_LAYERS];
};

struct CcNodeDev {
  int S, I, O;
  int has_t, pre, ut, integrator, n_stages;
  float u_scale, out_scale, dt, t0;
  int K0;       // rows of X/Z: S + has_t + I + 1
  int row_t;    // row index of t (valid iff has_t)
  int row_v;    // first input row
  int row_one;  // constant-1 row
  int cptS;     // state columns per thread = ceil(S / 4)
  int tiny;     // evaluated per lane in registers (one trajectory per lane) instead of as tiles
  CcMlpDev f, g;
  int w_X, w_Z;  // per-warp stage buffers [K0][32]
  int w_OUT;     // per-warp [O][32]
  int tape_row;  // first row of this node's state inside a tape record (-1: not on the tape)
  int P;         // parameter count
  // backward only
  int trainable;  // accumulate parameter gradients
  int recompute;  // stage intermediates must be recomputed (trainable or non-linear networks)
  int w_ZS[4];    // stage inputs kept for the reverse pass (ZS[0] aliases X)
  int w_HS[4][CC_MAX_LAYERS];  // hidden activations per stage
  int w_KS[4];    // stage outputs k_i (only if act_final is non-linear), else -1
  int w_XN;       // [K0][32] post-step readout source in keep mode (x_next ; t ; v ; 1), else -1
  int w_GH[CC_MAX_LAYERS];     // g hidden activations
  int w_LAM;      // adjoint of the state [S][32]
  int w_DA, w_DB; // cotangent ping-pong buffers [maxN][32]
  int w_VB;       // input cotangent [I][32]
  int w_OB;       // output cotangent [O][32]
  int w_GP;       // tiny nodes: lane-private gradient accumulators [g_total][32]
  int g_total;    // floats of this node's gradient record (sum of N*K over layers)
  int g_base;     // offset of this node's record inside a per-warp gradient record
};

struct CcKParams {
  CcNodeDev nd[2];  // closed loop: nd[0] = controller, nd[1] = model; open loop: nd[0] = model
  int n_nodes;
  int TN;           // register-tile width of the model kernels (8, 13 or 16)
  int ctrl_mode;    // CcCtrlMode
  int wpc;          // warps per CTA
  int warp_floats;  // per-warp shared memory (floats)
  int cta_floats;   // CTA-shared floats (weights) preceding the per-warp regions
  long long B;
  long long n_wtiles;  // ceil(B / 32)
  int T;  // number of steps: T (open loop) or Tr (closed loop)
  int R;  // closed loop: reference width
  const float* theta[2];
  const float* in;     // u[B,T,I] or refs[B,Tr,R]
  const float* noise;  // closed loop, optional
  const float* x0;     // open loop, optional
  float* out;          // y[B,T+1,O] or yhat[B,Tr,O]
  float* u_out;        // closed loop, optional
  float* x_final;      // open loop, optional
  float* tape;         // [n_ckpt][tape_rows][Bpad] (feature-major); then float ttab[2][T] (step times)
  int ckpt_every, n_ckpt, tape_rows;
  float* ttab;         // step times [2][T] (written by the forward pass, read by the reverse pass and by resumed segments)
  // segments (checkpointed recomputation, ckpt_every > 1): a launch covers the steps [seg_k0, seg_k1)
  int seg_k0, seg_k1;
  const float* carry_in;  // forward: the carry at step seg_k0 as one tape record [tape_rows][Bpad], or nullptr (fresh start)
  float* lam_io;          // reverse: adjoint carry [S_0 (+ S_1) + O][Bpad]: read if seg_k1 < T, written if seg_k0 > 0
  int seg_first;          // reverse: first launch of the pass (the gradient records are initialised, not accumulated)
  int tape_yrow;       // closed loop: first row of the fed-back output inside a tape record
  long long Bpad;      // B rounded up to a multiple of 32
  // backward
  const float* ybar;   // [B,T+1,O] or [B,Tr,O]
  float* ubar;         // open loop, optional [B,T,I]
  float* gpart;        // per-warp-tile gradient records [n_wtiles][g_rec]
  int g_rec;
  float* scratch;      // tcgen05 weight-gradient kernel: per-CTA stage-input slots [3][8*nch][128] (L2 resident)
  int w_yprev, w_yfb, w_araw;  // per-warp closed-loop scratch rows
  // tensor-core path (cc_tc_kernels.cuh): the MODEL node's f runs on tcgen05 (TMEM-resident stage inputs,
  // 3xTF32 split), one trajectory per thread; only the (tiny) controller uses the per-warp regions above
  int tc;          // 0: FFMA kernels, 1: tcgen05 kernels, 2: latency kernels, 3: tcgen05 reverse pass with weight gradient, 4: blocked linear kernels (cc_lin.h)
  int tc_nch;      // state register chunks of 8 per thread (4 or 7): S <= 8 * tc_nch
  int tc_KP;       // A columns (contraction, multiple of 8): fwd [x ; v ; 1], bwd [kbar]
  int tc_NP;       // D columns (multiple of 16): fwd k (S), bwd [zbar (8*nch) ; vbar (I)]
  int tc_off;      // float offset of the tcgen05 region inside dynamic shared memory
  int tc_sms;      // SM count (CTA start stagger: CTAs b and b + sms share an SM in the first wave)
  int tc_stagger;  // cycles the second CTA of an SM waits before its first step
  // blocked linear family (tc == 4; cc_lin.h): the linear model advances lin_m steps per MMA batch
  int lin_m;               // steps per block
  int lin_mt;              // open-loop forward with x_final and T % lin_m != 0: length of the ragged tail (second prep pack), else 0
  int lin_stage;           // coalesced I/O through shared-memory tiles (cc_lin_kernels.cuh)
  int lin_lc;              // closed loop: the controller is linear too -> its integrator step is collapsed (cc_lin_prep.cuh)
  const float* lin_pack;   // output of cc_lin_prep_kernel (block matrices + small tables), global memory
  double* lin_tab;         // fp64 tables for the gradient post-processing (trainable models), or nullptr
  float* lin_q;            // trainable open loop: per-CTA partial sums of Q = Lam (x) X
};
