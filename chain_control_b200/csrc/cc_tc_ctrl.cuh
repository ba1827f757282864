// cc_tc_ctrl.cuh -- register-resident "tiny" feedback controller for the tcgen05 kernels (thread = trajectory).
//
// Restates NeuralOdeController.step for single-layer f / g (neural_ode_controller.py:108-149, compact variant
// neural_ode_controller_compact_example.py:106-129 with out_scale = sigmoid(alpha) and the zero secondary) and its
// reverse pass.  Differences to cc_tiny_step / cc_tiny_bwd of the FFMA kernels: the controller state,
// the stage inputs and the adjoint live in REGISTERS, inputs use a fixed 16-column layout
//     [ x_0 .. x_7 | t | v_0 .. v_3 | 1 | 0 0 ]
// (weights staged accordingly, zero where the module has no such input), and the parameter gradient is
// accumulated per THREAD in lane-private shared-memory columns (no cross-lane work inside the time loop).
#pragma once
#include "cc_kernels.cuh"

// the launch plan is read straight from the __grid_constant__ kernel parameter (constant bank): a copy in shared memory
// (what the FFMA kernels use for their non-inlined functions) would have to be re-read after every shared-memory store,
// because the compiler cannot prove that the stage vectors / accumulators do not alias it
#define CC_DIRECT_PARAMS(gp)            \
  float* cc_smem = CC_SMEM_PTR;         \
  const CcKParams& p = (gp);            \
  float* cta_smem = cc_smem + CC_PARAM_FLOATS;

#define CC_TC_LANES 128  // threads per CTA of the tcgen05 kernels = stride of the lane-private shared-memory columns
#define CC_CK 16      // staged weight row length
#define CC_CKU 14     // columns that can be non-zero
#define CC_C_T 8      // column of t
#define CC_C_V 9      // first column of v
#define CC_C_ONE 13   // column of the constant 1

// the controller's shape and options as seen by the device functions below.  For the common compact controller
// (S_c = 5, [ref ; obs] = 2 inputs, 1 output, time-invariant, no activations: CFG = 521) the sizes are compile-time constants, so the
// per-row `if (n < S)` branches fold and the rows of a stage become one basic block the scheduler can interleave.
struct CcCtrlDims {
  int S, I, O, has_t, pre, integrator, ut, f_act, g_act;
  float u_scale, out_scale, dt;
};
template <int CFG>
CC_DEV CcCtrlDims cc_ctrl_dims(const CcNodeDev& C) {
  CcCtrlDims d;
  d.S = CFG ? CFG / 100 : C.S;
  d.I = CFG ? (CFG / 10) % 10 : C.I;
  d.O = CFG ? CFG % 10 : C.O;
  d.has_t = CFG ? 0 : C.has_t;
  d.pre = C.pre; d.integrator = C.integrator;
  // CFG shapes are only selected for controllers without activations / input transform (the reference's defaults), so
  // that the per-element activation dispatch folds away too
  d.ut = CFG ? (int)CC_UT_IDENTITY : C.ut;
  d.f_act = CFG ? (int)CC_ACT_IDENTITY : C.f.layer[0].act;
  d.g_act = CFG ? (int)CC_ACT_IDENTITY : C.g.layer[0].act;
  d.u_scale = C.u_scale; d.out_scale = C.out_scale; d.dt = C.dt;
  return d;
}

// node-layout physical row r of [x ; t ; v ; 1]  ->  fixed column
CC_DEV int cc_ctrl_col_of_row(const CcNodeDev& C, int r) {
  if (r < C.S) return r;
  r -= C.S;
  if (C.has_t) {
    if (r == 0) return CC_C_T;
    r -= 1;
  }
  if (r < C.I) return CC_C_V + r;
  return CC_C_ONE;
}

// WF[8][16], WG[4][16] (CTA-shared), zero padded
CC_DEV void cc_ctrl_stage_weights(const CcNodeDev& C, const float* theta, float* WF, float* WG, int tid, int nthreads) {
  for (int i = tid; i < (CC_TINY_S + CC_TINY_O) * CC_CK; i += nthreads) WF[i] = 0.f;  // WG follows WF
  __syncthreads();
  const CcLayerDev& F = C.f.layer[0];
  const CcLayerDev& G = C.g.layer[0];
  for (int i = tid; i < C.S * C.K0; i += nthreads) {
    const int n = i / C.K0, r = i % C.K0;
    WF[n * CC_CK + cc_ctrl_col_of_row(C, r)] = cc_weight_at(C, F, theta, n, r);
  }
  for (int i = tid; i < C.O * C.K0; i += nthreads) {
    const int o = i / C.K0, r = i % C.K0;
    WG[o * CC_CK + cc_ctrl_col_of_row(C, r)] = cc_weight_at(C, G, theta, o, r);
  }
}

struct CcCtrlIn {
  float c[CC_CKU];
};
CC_DEV void cc_ctrl_make_in(const float (&x)[CC_TINY_S], float t, const float (&vt)[CC_TINY_O], CcCtrlIn& in) {
#pragma unroll
  for (int i = 0; i < CC_TINY_S; ++i) in.c[i] = x[i];
  in.c[CC_C_T] = t;
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i) in.c[CC_C_V + i] = vt[i];
  in.c[CC_C_ONE] = 1.f;
}
CC_DEV void cc_ctrl_load_row(const float* row, float (&w)[CC_CK]) {
#pragma unroll
  for (int q = 0; q < CC_CK / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(row + 4 * q);
    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
  }
}
// is column k fed by this controller? (compile-time for CFG shapes)
CC_DEV bool cc_ctrl_live(const CcCtrlDims& C, int k) {
  return k < C.S || (k == CC_C_T && C.has_t) || (k >= CC_C_V && k < CC_C_V + C.I) || k == CC_C_ONE;
}
CC_DEV float cc_ctrl_dot(const CcCtrlDims& C, const float* row, const CcCtrlIn& in) {
  float w[CC_CK];
  cc_ctrl_load_row(row, w);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int k = 0; k < CC_CKU; k += 2) {
    if (cc_ctrl_live(C, k)) s0 = fmaf(w[k], in.c[k], s0);
    if (cc_ctrl_live(C, k + 1)) s1 = fmaf(w[k + 1], in.c[k + 1], s1);
  }
  return s0 + s1;
}
// k = act(WF . in) for the S state components (zero beyond)
CC_DEV void cc_ctrl_rhs(const CcCtrlDims& C, const float* WF, const CcCtrlIn& in, float (&k)[CC_TINY_S]) {
  const int S = C.S, fact = C.f_act;
#pragma unroll
  for (int n = 0; n < CC_TINY_S; ++n) {
    k[n] = 0.f;
    if (n < S) k[n] = cc_act(fact, cc_ctrl_dot(C, WF + n * CC_CK, in));
  }
}
CC_DEV void cc_ctrl_readout(const CcCtrlDims& C, const float* WG, const CcCtrlIn& in, float (&out)[CC_TINY_O]) {
  const int gact = C.g_act;
#pragma unroll
  for (int o = 0; o < CC_TINY_O; ++o) {
    out[o] = 0.f;
    if (o < C.O) out[o] = C.out_scale * cc_act(gact, cc_ctrl_dot(C, WG + o * CC_CK, in));
  }
}

// stage record kept for the reverse pass: lane-private shared-memory columns KEEP[(slot * 8 + n) * 128 + tid]
//   slots 0..3: stage inputs z_st (state part; z_0 = x), slot 4: x_next, slots 5..8: stage outputs k_st
//   (slots 5.. only when f's final activation is not the identity)
#define CC_CKEEP_Z 0
#define CC_CKEEP_XN 4
#define CC_CKEEP_K 5
CC_DEV int cc_ctrl_keep_floats(const CcCtrlDims& C) { return (C.f_act != CC_ACT_IDENTITY ? 9 : 5) * CC_TINY_S * 128; }

// pre-step readout only:  out = out_scale * g([x ; t ; v~])      (compact controller: before the state update)
CC_DEV void cc_ctrl_readout_pre(const CcCtrlDims& C, const float* WG, const float (&x)[CC_TINY_S], float t, const float (&v)[CC_TINY_O], float (&out)[CC_TINY_O]) {
  float vt[CC_TINY_O];
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i) vt[i] = (i < C.I) ? cc_ut(C.ut, C.u_scale, v[i]) : 0.f;
  CcCtrlIn in;
  cc_ctrl_make_in(x, t, vt, in);
  cc_ctrl_readout(C, WG, in, out);
}

// state update x <- integrate(f, x, t, dt) (integrate.py:1-28); post-step readout if !C.pre.
// KEEP (reverse pass): the stages are recorded in `keep` (lane-private column) and x is left untouched.
// One stage body, executed n_stages times with run-time coefficients (small code: the instruction cache matters).
template <bool KEEP>
CC_DEV void cc_ctrl_update(const CcCtrlDims& C, const float* WF, const float* WG, float (&x)[CC_TINY_S], float t, const float (&v)[CC_TINY_O],
                           float (&out)[CC_TINY_O], float* keep) {
  const float h = C.dt;
  const int fact = C.f_act;
  const bool rk4 = C.integrator == CC_INT_RK4;
  const int nst = rk4 ? 4 : 1;
  float vt[CC_TINY_O];
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i) vt[i] = (i < C.I) ? cc_ut(C.ut, C.u_scale, v[i]) : 0.f;
  CcCtrlIn in;
  cc_ctrl_make_in(x, t, vt, in);
  float ks[CC_TINY_S], xn[CC_TINY_S];
#pragma unroll
  for (int n = 0; n < CC_TINY_S; ++n) { ks[n] = 0.f; xn[n] = 0.f; }
#pragma unroll 1
  for (int st = 0; st < nst; ++st) {
    float k[CC_TINY_S];
    if (KEEP) {
#pragma unroll
      for (int n = 0; n < CC_TINY_S; ++n) keep[((CC_CKEEP_Z + st) * CC_TINY_S + n) * 128] = in.c[n];
    }
    cc_ctrl_rhs(C, WF, in, k);
    if (KEEP && fact != CC_ACT_IDENTITY) {
#pragma unroll
      for (int n = 0; n < CC_TINY_S; ++n) keep[((CC_CKEEP_K + st) * CC_TINY_S + n) * 128] = k[n];
    }
    if (rk4) {
      // ks = k1 + 2 k2 + 2 k3 + k4 ; stage inputs x + h k / 2, x + h k / 2, x + h k ; x_next = x + h (1/6 ks)
      const float wk = (st == 0 || st == 3) ? 1.f : 2.f;
      const float cz = st == 2 ? 1.f : 0.5f;
#pragma unroll
      for (int n = 0; n < CC_TINY_S; ++n) {
        ks[n] = fmaf(wk, k[n], ks[n]);
        in.c[n] = x[n] + (h * k[n]) * cz;
      }
      in.c[CC_C_T] = st == 2 ? t + h : t + h / 2.f;  // stage times t, t+h/2, t+h/2, t+h
      if (st == 3) {
#pragma unroll
        for (int n = 0; n < CC_TINY_S; ++n) xn[n] = x[n] + h * ((1.f / 6.f) * ks[n]);
      }
    } else {
#pragma unroll
      for (int n = 0; n < CC_TINY_S; ++n) xn[n] = C.integrator == CC_INT_EE ? x[n] + h * k[n] : k[n];
    }
  }
  if (!C.pre) {  // post-step readout g([x_next ; t ; v~]) with the PRE-step time
    cc_ctrl_make_in(xn, t, vt, in);
    cc_ctrl_readout(C, WG, in, out);
  }
  if (KEEP) {
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) keep[(CC_CKEEP_XN * CC_TINY_S + n) * 128] = xn[n];
  } else {
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) x[n] = xn[n];
  }
}

// the same update one stage at a time (forward kernel: the stages of the compact controller are placed behind the
// publishes of the model's stages, where they overlap the MMA batches):
//   cc_ctrl_begin -> cc_ctrl_stage(st = 0 .. n_stages-1) -> cc_ctrl_finish
struct CcCtrlRun {
  CcCtrlIn in;              // current stage input
  float ks[CC_TINY_S];      // k1 + 2 k2 + 2 k3 (+ k4)
  float xn[CC_TINY_S];      // x_next (valid after the last stage)
  float vt[CC_TINY_O];      // transformed input
};
CC_DEV int cc_ctrl_nstages(const CcCtrlDims& C) { return C.integrator == CC_INT_RK4 ? 4 : 1; }
CC_DEV void cc_ctrl_begin(const CcCtrlDims& C, const float (&x)[CC_TINY_S], float t, const float (&v)[CC_TINY_O], CcCtrlRun& r) {
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i) r.vt[i] = (i < C.I) ? cc_ut(C.ut, C.u_scale, v[i]) : 0.f;
  cc_ctrl_make_in(x, t, r.vt, r.in);
#pragma unroll
  for (int n = 0; n < CC_TINY_S; ++n) { r.ks[n] = 0.f; r.xn[n] = 0.f; }
}
template <bool KEEP>
CC_DEV void cc_ctrl_stage(const CcCtrlDims& C, const float* WF, int st, const float (&x)[CC_TINY_S], float t, CcCtrlRun& r, float* keep) {
  const float h = C.dt;
  const int fact = C.f_act;
  float k[CC_TINY_S];
  if (KEEP) {
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) keep[((CC_CKEEP_Z + st) * CC_TINY_S + n) * CC_TC_LANES] = r.in.c[n];
  }
  cc_ctrl_rhs(C, WF, r.in, k);
  if (KEEP && fact != CC_ACT_IDENTITY) {
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) keep[((CC_CKEEP_K + st) * CC_TINY_S + n) * CC_TC_LANES] = k[n];
  }
  if (C.integrator == CC_INT_RK4) {
    // ks = k1 + 2 k2 + 2 k3 + k4 ; stage inputs x + h k / 2, x + h k / 2, x + h k ; x_next = x + h (1/6 ks)
    const float wk = (st == 0 || st == 3) ? 1.f : 2.f;
    const float cz = st == 2 ? 1.f : 0.5f;
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) {
      r.ks[n] = fmaf(wk, k[n], r.ks[n]);
      r.in.c[n] = x[n] + (h * k[n]) * cz;
    }
    r.in.c[CC_C_T] = st == 2 ? t + h : t + h / 2.f;  // stage times t, t+h/2, t+h/2, t+h
    if (st == 3) {
#pragma unroll
      for (int n = 0; n < CC_TINY_S; ++n) r.xn[n] = x[n] + h * ((1.f / 6.f) * r.ks[n]);
    }
  } else {
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) r.xn[n] = C.integrator == CC_INT_EE ? x[n] + h * k[n] : k[n];
  }
}
template <bool KEEP>
CC_DEV void cc_ctrl_finish(const CcCtrlDims& C, const float* WG, float (&x)[CC_TINY_S], float t, CcCtrlRun& r, float (&out)[CC_TINY_O], float* keep) {
  if (!C.pre) {  // post-step readout g([x_next ; t ; v~]) with the PRE-step time
    cc_ctrl_make_in(r.xn, t, r.vt, r.in);
    cc_ctrl_readout(C, WG, r.in, out);
  }
  if (KEEP) {
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) keep[(CC_CKEEP_XN * CC_TINY_S + n) * CC_TC_LANES] = r.xn[n];
  } else {
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) x[n] = r.xn[n];
  }
}

// lane-private gradient accumulators: GP[(row * CC_CKU + col) * 128 + tid], rows 0..S-1 = f outputs, S.. = g outputs
CC_DEV void cc_ctrl_gacc(float* GP, int row, const float d, const CcCtrlIn& in, int S, bool has_t, int I) {
  float* g = GP + row * CC_CKU * 128;
#pragma unroll
  for (int k = 0; k < CC_TINY_S; ++k)
    if (k < S) g[k * 128] = fmaf(d, in.c[k], g[k * 128]);
  if (has_t) g[CC_C_T * 128] = fmaf(d, in.c[CC_C_T], g[CC_C_T * 128]);
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i)
    if (i < I) g[(CC_C_V + i) * 128] = fmaf(d, in.c[CC_C_V + i], g[(CC_C_V + i) * 128]);
  g[CC_C_ONE * 128] += d;
}

// reverse pass of one controller step.  keep: record of cc_ctrl_update<true>; out: the step's output; ob: its
// cotangent; lam: cotangent of x_next (in) / of x (out); vb: cotangent of the raw input
CC_DEV void cc_ctrl_bwd(const CcCtrlDims& C, const float* WF, const float* WG, float* GP, float t, const float (&v)[CC_TINY_O],
                        const float* keep, const float (&out)[CC_TINY_O], const float (&ob)[CC_TINY_O], float (&lam)[CC_TINY_S],
                        float (&vb)[CC_TINY_O]) {
  const float h = C.dt;
  const int fact = C.f_act, gact = C.g_act;
  float vt[CC_TINY_O];
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i) { vt[i] = (i < C.I) ? cc_ut(C.ut, C.u_scale, v[i]) : 0.f; vb[i] = 0.f; }
  float utb[CC_TINY_O];
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i) utb[i] = 0.f;
  // readout
  float xrb[CC_TINY_S];
#pragma unroll
  for (int n = 0; n < CC_TINY_S; ++n) xrb[n] = 0.f;
  {
    float xin[CC_TINY_S];
    const int slot = C.pre ? CC_CKEEP_Z : CC_CKEEP_XN;
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) xin[n] = keep[(slot * CC_TINY_S + n) * 128];
    CcCtrlIn in;
    cc_ctrl_make_in(xin, t, vt, in);
#pragma unroll 1
    for (int o = 0; o < C.O; ++o) {
      float d = C.out_scale * ob[0];
#pragma unroll
      for (int i = 1; i < CC_TINY_O; ++i)
        if (i == o) d = C.out_scale * ob[i];
      float oo = out[0];
#pragma unroll
      for (int i = 1; i < CC_TINY_O; ++i)
        if (i == o) oo = out[i];
      if (gact != CC_ACT_IDENTITY) d *= cc_act_grad(gact, C.out_scale != 0.f ? oo / C.out_scale : 0.f);
      float w[CC_CK];
      cc_ctrl_load_row(WG + o * CC_CK, w);
#pragma unroll
      for (int n = 0; n < CC_TINY_S; ++n) xrb[n] = fmaf(w[n], d, xrb[n]);
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) utb[i] = fmaf(w[CC_C_V + i], d, utb[i]);  // feed-through columns (transformed input)
      cc_ctrl_gacc(GP, C.S + o, d, in, C.S, C.has_t != 0, C.I);
    }
  }
  if (!C.pre) {
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) lam[n] += xrb[n];
  }
  const bool rk4 = C.integrator == CC_INT_RK4;
  const int nst = rk4 ? 4 : 1;
  const float c6 = h / 6.f, c3 = h / 3.f, c2 = h / 2.f;
  float zb[CC_TINY_S], zsum[CC_TINY_S];
#pragma unroll
  for (int n = 0; n < CC_TINY_S; ++n) { zb[n] = 0.f; zsum[n] = 0.f; }
#pragma unroll 1
  for (int st = nst - 1; st >= 0; --st) {
    const float ca = rk4 ? ((st == 3 || st == 0) ? c6 : c3) : (C.integrator == CC_INT_EE ? h : 1.f);
    const float cb = rk4 ? (st == 3 ? 0.f : (st == 2 ? h : c2)) : 0.f;
    const float ts = rk4 ? (st == 0 ? t : (st == 3 ? t + h : t + h / 2.f)) : t;
    float zin[CC_TINY_S];
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) zin[n] = keep[((CC_CKEEP_Z + st) * CC_TINY_S + n) * 128];
    CcCtrlIn in;
    cc_ctrl_make_in(zin, ts, vt, in);
    float acc[CC_CKU];
#pragma unroll
    for (int k = 0; k < CC_CKU; ++k) acc[k] = 0.f;
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) {
      if (n < C.S) {
        float kb = ca * lam[n] + cb * zb[n];
        if (fact != CC_ACT_IDENTITY) kb *= cc_act_grad(fact, keep[((CC_CKEEP_K + st) * CC_TINY_S + n) * 128]);
        float w[CC_CK];
        cc_ctrl_load_row(WF + n * CC_CK, w);
#pragma unroll
        for (int k = 0; k < CC_CKU; ++k)
          if (cc_ctrl_live(C, k)) acc[k] = fmaf(w[k], kb, acc[k]);
        cc_ctrl_gacc(GP, n, kb, in, C.S, C.has_t != 0, C.I);
      }
    }
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) { zb[n] = acc[n]; zsum[n] += acc[n]; }
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) utb[i] += acc[CC_C_V + i];
  }
#pragma unroll
  for (int n = 0; n < CC_TINY_S; ++n) {
    float l = C.integrator == CC_INT_NONE ? zb[n] : lam[n] + zsum[n];
    if (C.pre) l += xrb[n];
    lam[n] = l;
  }
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i)
    if (i < C.I) vb[i] = utb[i] * cc_ut_grad(C.ut, C.u_scale, v[i]);
}
