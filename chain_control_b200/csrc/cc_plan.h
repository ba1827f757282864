// cc_plan.h — host-built launch plan shared by the forward and backward kernels.
//
// Data layout.  A CTA owns a tile of BT trajectories for the whole time loop.  Every
// activation buffer in shared memory is feature-major: buf[row][BTP] with the BT trajectories
// contiguous (BTP = padded row stride), so that a thread's TM=4 trajectories are one float4.
// A module ("node") keeps three stage buffers X / ZA / ZB whose rows are
//     [ x (S) ; t (1, iff the module tracks time) ; v (I, the transformed input) ; 1 (bias row) ]
// which is exactly the reference's batch_concat order of f's and g's inputs
// (neural_ode_model.py:122-133, neural_ode_controller.py:121-134); weights are staged once per CTA
// as Wt[row][Np] (k-major) with zero rows where an MLP does not consume t / v and the bias as the
// row multiplying the constant-1 row.  Each layer evaluation is then a register-tiled SGEMM
//     C[BT x N] = A[K x BT]^T . Wt[K x N],   thread tile TM x TN = 4 trajectories x 8 (or 10) outputs.
#pragma once
#include <stdint.h>

#include "../../include/cc_b200.h"

#define CC_TM 4
#define CC_TC 16        /* time steps staged per I/O chunk */
#define CC_MAX_SKINNY 4 /* layers with N <= this use the split-K path */

struct CcLayerDev {
  int K;       // contraction rows read from the source buffer (incl. the constant-1 row)
  int N;       // outputs
  int CG;      // column groups = ceil(N / TN)
  int Np;      // TN * CG rounded up to a multiple of 4 (row stride of Wt)
  int Kp;      // TN * CGk rounded up to a multiple of 4 (row stride of the row-major copy)
  int CGk;     // column groups over K (input-gradient GEMM)
  int act;     // activation applied to this layer's output
  int in_log;  // logical input width = row stride of the weight in theta
  int w_off;   // offset of weight[N][in_log] in theta
  int b_off;   // offset of bias[N] in theta, or -1
  int first;   // first layer of its MLP: physical->logical row map goes through the node layout
  int use_t;   // (first layers) this MLP consumes the time row
  int use_v;   // (first layers) this MLP consumes the input rows
  int s_wt;    // smem offset (floats) of Wt[K][Np]
  int s_w;     // smem offset of W[N][Kp]      (backward only, else -1)
  int s_gw;    // smem offset of Wbar[N][Kp]   (backward & trainable only, else -1)
  int s_out;   // smem offset of the hidden output buffer [N+1][BTP] (hidden layers), else -1
};

struct CcMlpDev {
  int L;
  CcLayerDev layer[CC_MAX_LAYERS];
};

struct CcNodeDev {
  int S, I, O;
  int has_t, pre, ut, integrator, n_stages;
  float u_scale, out_scale, dt;
  int K0;       // rows of X/ZA/ZB: S + has_t + I + 1
  int row_t;    // row index of t (valid iff has_t)
  int row_v;    // first input row
  int row_one;  // constant-1 row
  int CGs;      // column groups over the state (== f.layer[L-1].CG)
  CcMlpDev f, g;
  int s_X, s_ZA, s_ZB;  // stage buffers [K0][BTP]
  int s_OUT;            // [O][BTP]
  int s_P;              // split-K scratch [CGT][CC_MAX_SKINNY][BTP]
  int tape_row;         // first row of this node's state inside a tape record
  int P;                // parameter count
  // backward only
  int trainable;    // accumulate parameter gradients
  int recompute;    // stage intermediates must be recomputed (trainable or non-linear f)
  int s_Z[4];       // stage inputs kept for the reverse pass (Z[0] aliases X)
  int s_HS[4][CC_MAX_LAYERS];  // hidden activations per stage
  int s_KS[4];      // stage outputs k_i (only if act_final is non-linear), else -1
  int s_GH[CC_MAX_LAYERS];     // g hidden activations
  int s_LAM;        // adjoint of the state [S][BTP]
  int s_DA, s_DB;   // cotangent ping-pong buffers [maxN+?][BTP]
  int s_VB;         // input cotangent [I][BTP]
  int s_OB;         // output cotangent [O][BTP]
  int gw_off;       // offset of this node's gradient block inside a per-CTA partial record
};

struct CcKParams {
  CcNodeDev nd[2];  // closed loop: nd[0] = controller, nd[1] = model; open loop: nd[0] = model
  int n_nodes;
  int TN;  // thread-tile width over outputs: 8 or 10 (chosen to minimise column padding)
  int BT, BTP, RG, CGT, nthreads;
  long long B;
  int T;  // number of steps: T (open loop) or Tr (closed loop)
  int R;  // closed loop: reference width
  const float* theta[2];
  const float* in;     // u[B,T,I] or refs[B,Tr,R]
  const float* noise;  // closed loop, optional
  const float* x0;     // open loop, optional
  float* out;          // y[B,T+1,O] or yhat[B,Tr,O]
  float* u_out;        // closed loop, optional
  float* x_final;      // open loop, optional
  float* tape;         // [n_ckpt][tape_rows][Bpad]  (feature-major, independent of the tile size)
  int ckpt_every, n_ckpt, tape_rows;
  long long Bpad;      // B rounded up to a multiple of 256; after the records: float ttab[2][T] (step times)
  // backward
  const float* ybar;   // [B,T+1,O] or [B,Tr,O]
  float* ubar;         // open loop, optional [B,T,I]
  float* gpart;        // per-CTA gradient partials [n_tiles][gw_total]
  int gw_total;
  // I/O staging (floats)
  int s_in, s_out, s_uo, s_noise, s_yprev, s_yb, s_ub, s_yfb, s_araw;
  int smem_floats;
};
