// cc_lin.h -- layout constants of the BLOCKED LINEAR kernel family ("lin", CcKParams::tc == 4), shared by the host plan
// (cc_abi.cu), the fp64 preparation kernels (cc_lin_prep.cuh) and the tcgen05 rollout kernels (cc_lin_kernels.cuh).
//
// Scope: MODEL nodes whose f and g are single Linear layers with identity final activation, time-invariant (the compact
// model of the reference's notebooks 3/4 and end_to_end_learning, neural_ode_model_compact_example.py:85-124, and the
// full model with f_depth = g_depth = 0, neural_ode_model.py:106-149).  For such a model one integrator step
// (nn_lib/integrate.py:1-28) is the affine map
//     x+ = Phi x + Gam (B u~ + c),   RK4: Phi = sum_{n<=4} (hA)^n/n!,  Gam = h sum_{n<=3} (hA)^n/(n+1)!
// and m consecutive steps collapse into ONE 64 x 64 matrix product per block of m steps (tests/linblock_ref.py states
// the algebra in NumPy and tests/test_linblock_algebra.py checks it against the oracle to machine precision):
//   closed loop ("P" formulation)   [p' - p ; yfree_0..m-1] = Wf [p ; u~_prev(0..m-1) ; e]
//        x_k = p + C u~_prev + c_m e;  y_j = yfree_j + const_j + sum_{i<j+q} H_{j+q-1-i} u~_i  (thread-side, scalar)
//   open loop   ("X" formulation)   [x_{k+m} - x_k ; y_0..m-1] = Wf [x_k ; u~_0..m-1 ; 1]
//   and the transposed recursions for the reverse pass (Wb).
//
// Column convention of the 64-column MMA operands (A) and results (D), the same for every matrix:
//   columns [0, S)       the state-like block (p, x, r, mu)
//   column  62 - t       "slot" t of the per-step block (u~_t, yfree_t, w_t, ufree_t ...; t = step * width + component)
//   column  63           the constant 1 (forward A operands only)
// so that slots sit at COMPILE-TIME columns whatever S is; block length m = min(13, 63 - S) / max(I, O)
// (at most 13 slots: columns 50..62).
#pragma once
#include "cc_plan.h"
#define CC_LIN_OPEN 1

#define CC_LIN_N 64                 // MMA N = K = 64
#define CC_LIN_MAXM 13              // max steps per block
#define CC_LIN_SLOT(t) (62 - (t))   // column of slot t
#define CC_LIN_ONE 63               // column of the constant 1

// ---- prep pack (global memory, written by cc_lin_prep_kernel): float offsets ----
#define CC_LIN_WF_HI 0              // forward matrix, tf32 hi part, MMA B-operand layout ((k>>2)*64 + n)*4 + (k&3)
#define CC_LIN_WF_LO 4096
#define CC_LIN_WB_HI 8192           // reverse matrix
#define CC_LIN_WB_LO 12288
#define CC_LIN_SMALL 16384          // start of the small tables
#define CC_LIN_S_H 0                //   H[n][o][i]   Markov parameters G Phi^n Gam_u, n = 0..2m (stride 16 per n)
#define CC_LIN_S_CONST 448          //   const[j][o]  readout constants d + G sum_{i<j+q} Phi^i gam   (stride 4 per j)
#define CC_LIN_S_G0 512             //   G[o][64]     readout rows (out_scale folded), zero padded
#define CC_LIN_S_D0 768             //   d[o]
#define CC_LIN_S_CM 772             //   c_m[64]
#define CC_LIN_S_C 836              //   C[s][16]     x_k = p + C u~_prev + c_m e      (slot-indexed columns)
#define CC_LIN_S_CHAT 1860          //   Chat[s][16]  mu_k = r + Chat w_blk
#define CC_LIN_SMALL_FLOATS 2884     // end of the model tables
// linear tiny controller (f_c, g_c single Linear layers without activations, time-invariant): its integrator step collapsed
// the same way, x_c+ = Phi_c x_c + (Gam_c B_c) v + Gam_c c_c
#define CC_LIN_S_CPHI 2884          //   (Phi_c - I)[8][8]   (increment form: x_c+ = x_c + (Phi_c - I) x_c + ...)
#define CC_LIN_S_CGB 2948           //   (Gam_c B_c)[8][4]
#define CC_LIN_S_CGAM 2980          //   (Gam_c c_c)[8]
#define CC_LIN_S_CTAB 2992          //   fp64: A_c[8][8], Gam_c[8][8]  (gradient post-processing, cc_lin_cgrad_kernel)
#define CC_LIN_SMALL_ALL 3248
#define CC_LIN_PACK_FLOATS (CC_LIN_SMALL + CC_LIN_SMALL_ALL)  // 19,632 floats (multiple of 16)
// ---- fp64 tables behind the pack (trainable models only: the post-processing of Q needs them) ----
//   PW[n][64][64] (Phi^n, n = 0..m), RG[a][4][64], CU[a][64][4], GS[a][64], plus A, B, c, Gam
#define CC_LIN_TAB_PW 0
#define CC_LIN_TAB_RG(m) (((m) + 1) * 4096)
#define CC_LIN_TAB_CU(m) (CC_LIN_TAB_RG(m) + 28 * 256)
#define CC_LIN_TAB_GS(m) (CC_LIN_TAB_CU(m) + 28 * 256)
#define CC_LIN_TAB_A(m) (CC_LIN_TAB_GS(m) + 28 * 64)
#define CC_LIN_TAB_GAM(m) (CC_LIN_TAB_A(m) + 4096)
#define CC_LIN_TAB_DOUBLES(m) (CC_LIN_TAB_GAM(m) + 4096)

// ---- fp64 post-processing Q -> theta_bar (cc_lin_open.cuh) ----
struct CcLinGradArgs {
  CcNodeDev M;
  const float* qpart;   // [n_cta][128][64] partial sums of Q, then [n_cta][4][4][64] partial sums of ybar_0 (x) [x0 ; 1]
  int n_cta;
  const double* tab;    // tables of cc_lin_prep_kernel
  const float* theta;
  double* work;         // [64*64] Q, then per j: Mx[64*64], Mu[64*4], M1[64], Gb[4*64], db[4]  (CC_LIN_GRAD_WORK_DOUBLES)
  float* theta_bar;
  int m;
};
#define CC_LIN_GRAD_PER_J (4096 + 256 + 64 + 256 + 4)
#define CC_LIN_GRAD_WORK_DOUBLES(m) (4096 + 4 * 64 + (m) * CC_LIN_GRAD_PER_J)

// ---- gradient of a linear tiny controller from the reduced M records (cc_lin_prep.cuh: cc_lin_cgrad_kernel) ----
struct CcLinCgradArgs {
  CcNodeDev C;
  const float* theta_c;
  const float* pack;     // prep pack (fp64 A_c, Gam_c at CC_LIN_S_CTAB)
  const float* msum;     // [P] reduced records (M space)
  float* theta_bar;      // [P]
};
