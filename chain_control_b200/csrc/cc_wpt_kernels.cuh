// cc_wpt_kernels.cuh -- latency kernels: ONE WARP PER TRAJECTORY, for small batches of depth-0 models.
//
// The reference's own configurations are tiny (BASELINE.json configs[0]/[1]: 8 and 64 trajectories x 1001 steps): a
// train step is then a 1001-step dependent chain and only latency matters.  The tcgen05 kernels need ~1.3 us per RK4
// stage (MMA batch + TMEM round trip + barriers) whatever the batch; here a stage is a 50x52 mat-vec spread over the 32
// lanes of a warp:
//   * lane l owns the state components l and l+32 and holds the two matching ROWS of f's weight matrix (forward) or
//     of its transpose (reverse pass) in registers (2 x 56);
//   * the stage input is a 64-float vector in the warp's shared memory, read back as broadcast LDS.128;
//   * __syncwarp() is the only synchronisation; readout dot products are warp shuffles;
//   * the tiny controller is evaluated redundantly by every lane with the register-resident code of cc_tc_ctrl.cuh
//     (no communication; lane 0's copy of the parameter gradient is the one that is kept).
// Same reference functions, tape layout and gradient records as the other kernel families (cc_tc_kernels.cuh).
#pragma once
#include "cc_tc_ctrl.cuh"

#define CC_WPT_KW 56      // rows of the stage input vector held per lane: [x (S) ; v (I) ; 1] zero padded
#define CC_WPT_THREADS 128
#define CC_WPT_ZB 64      // floats of the per-warp stage vector

struct CcWptShared {
  float* wfc; float* wgc; float* gmod; float* zbuf; float* gp;
};
// floats after the parameter copy: controller weights, model readout rows [4][64] (+bias at [o][63]), 4 stage vectors
#define CC_WPT_REGION_FLOATS ((CC_TINY_S + CC_TINY_O) * CC_CK + CC_TINY_O * 64 + 4 * CC_WPT_ZB)
CC_DEV CcWptShared cc_wpt_carve(float* smem) {
  CcWptShared s;
  s.wfc = smem + ((CC_PARAM_FLOATS + 15) / 16 * 16);
  s.wgc = s.wfc + CC_TINY_S * CC_CK;
  s.gmod = s.wgc + CC_TINY_O * CC_CK;
  s.zbuf = s.gmod + CC_TINY_O * 64;
  s.gp = s.zbuf + 4 * CC_WPT_ZB;
  return s;
}
CC_DEV float cc_warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(CC_FULL, v, off);
  return v;
}
// two dot products of register rows against the broadcast stage vector (four partial sums each: latency)
CC_DEV void cc_wpt_dot2(const float (&r0)[CC_WPT_KW], const float (&r1)[CC_WPT_KW], const float* z, float& k0, float& k1) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
  for (int q = 0; q < CC_WPT_KW / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(z + 4 * q);
    a0 = fmaf(r0[4 * q], v.x, a0); a1 = fmaf(r0[4 * q + 1], v.y, a1); a2 = fmaf(r0[4 * q + 2], v.z, a2); a3 = fmaf(r0[4 * q + 3], v.w, a3);
    b0 = fmaf(r1[4 * q], v.x, b0); b1 = fmaf(r1[4 * q + 1], v.y, b1); b2 = fmaf(r1[4 * q + 2], v.z, b2); b3 = fmaf(r1[4 * q + 3], v.w, b3);
  }
  k0 = (a0 + a1) + (a2 + a3);
  k1 = (b0 + b1) + (b2 + b3);
}
// y_o = out_scale * act(g_o . x + b_o): lane partials + shuffle reduction (every lane gets the result)
CC_DEV void cc_wpt_readout(const float* gmod, int lane, float x0, float x1, int O, int gact, float out_scale, float (&y)[CC_TINY_O]) {
#pragma unroll
  for (int o = 0; o < CC_TINY_O; ++o) {
    y[o] = 0.f;
    if (o < O) {
      const float s = cc_warp_sum(fmaf(gmod[o * 64 + lane], x0, gmod[o * 64 + 32 + lane] * x1));
      y[o] = out_scale * cc_act(gact, s + gmod[CC_TINY_O * 64 - CC_TINY_O + o]);
    }
  }
}

#define CC_WPT_PROLOGUE(gp)                                                                       \
  CC_DIRECT_PARAMS(gp)                                                                            \
  (void)cta_smem;                                                                                 \
  const bool closed = CM != CC_CTRL_NONE;                                                         \
  const CcNodeDev& M = p.nd[closed ? 1 : 0];                                                      \
  const CcNodeDev& C = p.nd[0];                                                                   \
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                                  \
  const CcWptShared sh = cc_wpt_carve(cc_smem);                                                   \
  if (closed) cc_ctrl_stage_weights(C, p.theta[0], sh.wfc, sh.wgc, tid, blockDim.x);              \
  {                                                                                               \
    const CcLayerDev& G_ = M.g.layer[0];                                                          \
    for (int i = tid; i < CC_TINY_O * 64; i += blockDim.x) {                                      \
      const int o = i / 64, n = i % 64;                                                           \
      sh.gmod[i] = (o < M.O && n < M.S) ? cc_weight_at(M, G_, p.theta[closed ? 1 : 0], o, n) : 0.f; \
    }                                                                                             \
    __syncthreads();                                                                              \
    if (tid < CC_TINY_O) sh.gmod[CC_TINY_O * 64 - CC_TINY_O + tid] = tid < M.O ? cc_weight_at(M, G_, p.theta[closed ? 1 : 0], tid, M.row_one) : 0.f; \
  }                                                                                               \
  __syncthreads();                                                                                \
  const CcCtrlDims CD = cc_ctrl_dims<CFG>(C);                                                     \
  const long long b = (long long)blockIdx.x * 4 + warp;                                           \
  const bool actw = b < p.B;                                                                      \
  const bool io = actw && lane == 0; /* the lane that touches global memory */                    \
  float* zb = sh.zbuf + warp * CC_WPT_ZB;                                                         \
  const int n0 = lane, n1 = lane + 32;                                                            \
  const bool v0 = n0 < M.S, v1 = n1 < M.S;                                                        \
  const float h = M.dt;                                                                           \
  const int nst = M.integrator == CC_INT_RK4 ? 4 : 1;

// ------------------------------------------------------------------------------------------------
// forward rollout (open loop: CM = CC_CTRL_NONE, closed loop: CC_CTRL_TINY)
// ------------------------------------------------------------------------------------------------
template <int CM, int CFG>
__global__ void __launch_bounds__(CC_WPT_THREADS) cc_wpt_fwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  CC_WPT_PROLOGUE(gp)
  const CcLayerDev& F = M.f.layer[0];
  const CcLayerDev& G = M.g.layer[0];
  const bool tape = p.tape != nullptr;
  const bool fnl = F.act != CC_ACT_IDENTITY;  // non-linear final activation of f
  const float* thm = p.theta[closed ? 1 : 0];
  // rows n0 / n1 of f's weight matrix in the physical input order [x ; v ; 1]
  float r0[CC_WPT_KW], r1[CC_WPT_KW];
#pragma unroll
  for (int k = 0; k < CC_WPT_KW; ++k) {
    r0[k] = (v0 && k < M.K0) ? cc_weight_at(M, F, thm, n0, k) : 0.f;
    r1[k] = (v1 && k < M.K0) ? cc_weight_at(M, F, thm, n1, k) : 0.f;
  }
  zb[lane] = 0.f; zb[lane + 32] = 0.f;
  __syncwarp();
  if (lane == 0) zb[M.row_one] = 1.f;
  float x0 = (!closed && p.x0 && actw && v0) ? __ldg(p.x0 + b * M.S + n0) : 0.f;
  float x1 = (!closed && p.x0 && actw && v1) ? __ldg(p.x0 + b * M.S + n1) : 0.f;
  float xc[CC_TINY_S];
#pragma unroll
  for (int i = 0; i < CC_TINY_S; ++i) xc[i] = 0.f;
  float yprev[CC_TINY_O];
  cc_wpt_readout(sh.gmod, lane, x0, x1, M.O, G.act, M.out_scale, yprev);  // y0 = g(x0)
  if (!closed && io)
    for (int o = 0; o < M.O; ++o) p.out[b * (long long)(p.T + 1) * M.O + o] = yprev[o];
  float tc = closed ? C.t0 : 0.f;
  const int Win = closed ? p.R : M.I;
  float vin[CC_TINY_O];
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i) vin[i] = (i < Win && actw && p.T > 0) ? __ldg(p.in + b * (long long)p.T * Win + i) : 0.f;
  for (int k = 0; k < p.T; ++k) {
    const bool wt = tape && (k % p.ckpt_every == 0);
    const int ck = k / p.ckpt_every;
    if (tape && b == 0 && lane == 0) {
      float* ttab = p.tape + (long long)p.n_ckpt * p.tape_rows * p.Bpad;
      if (closed) { ttab[k] = tc; ttab[p.T + k] = 0.f; } else { ttab[k] = 0.f; }
    }
    float vcur[CC_TINY_O];
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) vcur[i] = vin[i];
    if (k + 1 < p.T) {
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i)
        if (i < Win && actw) vin[i] = __ldg(p.in + b * (long long)p.T * Win + (long long)(k + 1) * Win + i);
    }
    float a[CC_TINY_O], vc[CC_TINY_O];
    if (closed) {
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) {
        vc[i] = vcur[i];
#pragma unroll
        for (int o = 0; o < CC_TINY_O; ++o)
          if (o < M.O && i == p.R + o) vc[i] = yprev[o];
      }
      if (wt && io) {
        for (int o = 0; o < M.O; ++o) p.tape[((long long)ck * p.tape_rows + p.tape_yrow + o) * p.Bpad + b] = yprev[o];
#pragma unroll
        for (int s = 0; s < CC_TINY_S; ++s)
          if (s < C.S) p.tape[((long long)ck * p.tape_rows + C.tape_row + s) * p.Bpad + b] = xc[s];
      }
      if (wt && actw && M.tape_row >= 0) {  // non-linear frozen model: its state is part of the carry
        if (v0) p.tape[((long long)ck * p.tape_rows + M.tape_row + n0) * p.Bpad + b] = x0;
        if (v1) p.tape[((long long)ck * p.tape_rows + M.tape_row + n1) * p.Bpad + b] = x1;
      }
      if (C.pre) cc_ctrl_readout_pre(CD, sh.wgc, xc, tc, vc, a);
      else cc_ctrl_update<false>(CD, sh.wfc, sh.wgc, xc, tc, vc, a, nullptr);
      if (p.u_out && io)
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i)
          if (i < M.I) p.u_out[b * (long long)p.T * M.I + (long long)k * M.I + i] = a[i];
    } else {
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) a[i] = vcur[i];
      if (wt && actw) {
        if (v0) p.tape[((long long)ck * p.tape_rows + n0) * p.Bpad + b] = x0;
        if (v1) p.tape[((long long)ck * p.tape_rows + n1) * p.Bpad + b] = x1;
      }
    }
    float y[CC_TINY_O];
    if (M.pre) cc_wpt_readout(sh.gmod, lane, x0, x1, M.O, G.act, M.out_scale, y);  // compact: y = g(x) BEFORE the update
    // stage vector [x ; v~ ; 1]
    if (v0) zb[n0] = x0;
    if (v1) zb[n1] = x1;
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i)
      if (i < M.I && lane == i) zb[M.row_v + i] = cc_ut(M.ut, M.u_scale, a[i]);
    __syncwarp();
    float s0 = 0.f, s1 = 0.f;
    // the compact controller's state update is independent of the model's step: its RK4 stages are issued next to the
    // model's stages so that the scheduler interleaves the two dependent chains
    CcCtrlRun crun;
    const bool cpipe = closed && C.pre;
    const int nstc = closed ? cc_ctrl_nstages(CD) : 0;
    if (cpipe) cc_ctrl_begin(CD, xc, tc, vc, crun);
#pragma unroll 1
    for (int st = 0; st < nst; ++st) {
      if (cpipe && st < nstc) cc_ctrl_stage<false>(CD, sh.wfc, st, xc, tc, crun, nullptr);
      float k0, k1;
      cc_wpt_dot2(r0, r1, zb, k0, k1);
      if (fnl) { k0 = cc_act(F.act, k0); k1 = cc_act(F.act, k1); }
      __syncwarp();  // every lane has read the stage vector
      if (nst == 4) {
        if (st < 3) {
          const float wk = st == 0 ? 1.f : 2.f, cz = st == 2 ? 1.f : 0.5f;
          s0 = st == 0 ? k0 : s0 + wk * k0;
          s1 = st == 0 ? k1 : s1 + wk * k1;
          if (v0) zb[n0] = x0 + (h * k0) * cz;
          if (v1) zb[n1] = x1 + (h * k1) * cz;
          __syncwarp();
        } else {
          x0 = x0 + h * ((1.f / 6.f) * (s0 + k0));
          x1 = x1 + h * ((1.f / 6.f) * (s1 + k1));
        }
      } else {
        x0 = x0 + h * k0;
        x1 = x1 + h * k1;
      }
    }
    if (!v0) x0 = 0.f;
    if (!v1) x1 = 0.f;
    if (closed) {
      if (cpipe) {
#pragma unroll 1
        for (int st = nst; st < nstc; ++st) cc_ctrl_stage<false>(CD, sh.wfc, st, xc, tc, crun, nullptr);
        float dummy[CC_TINY_O];
        cc_ctrl_finish<false>(CD, sh.wgc, xc, tc, crun, dummy, nullptr);
      }
      if (C.has_t) tc = tc + C.dt;
    }
    if (!M.pre) cc_wpt_readout(sh.gmod, lane, x0, x1, M.O, G.act, M.out_scale, y);  // full model: y = g(x_next)
    if (closed) {
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o)
        if (o < M.O) {
          if (p.noise && actw) y[o] += __ldg(p.noise + b * (long long)p.T * M.O + (long long)k * M.O + o);
          yprev[o] = y[o];
          if (io) p.out[b * (long long)p.T * M.O + (long long)k * M.O + o] = y[o];
        }
    } else if (io) {
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o)
        if (o < M.O) p.out[b * (long long)(p.T + 1) * M.O + (long long)(k + 1) * M.O + o] = y[o];
    }
  }
  if (!closed && p.x_final && actw) {
    if (v0) p.x_final[b * M.S + n0] = x0;
    if (v1) p.x_final[b * M.S + n1] = x1;
  }
}

// ------------------------------------------------------------------------------------------------
// reverse pass of the closed loop, frozen LINEAR model + tiny trainable controller
// ------------------------------------------------------------------------------------------------
template <int CM, int CFG>
__global__ void __launch_bounds__(CC_WPT_THREADS) cc_wpt_bwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  CC_WPT_PROLOGUE(gp)
  const CcLayerDev& F = M.f.layer[0];
  const float* thm = p.theta[1];
  const float* ttab = p.tape + (long long)p.n_ckpt * p.tape_rows * p.Bpad;
  // lane l owns the INPUTS j = l, l+32 of f (state components, then the raw inputs v): rows of the transpose
  const int SI = M.S + M.I;
  const bool w0 = n0 < SI, w1 = n1 < SI;
  float c0[CC_WPT_KW], c1[CC_WPT_KW];
#pragma unroll
  for (int m = 0; m < CC_WPT_KW; ++m) {
    c0[m] = (w0 && m < M.S) ? cc_weight_at(M, F, thm, m, n0) : 0.f;
    c1[m] = (w1 && m < M.S) ? cc_weight_at(M, F, thm, m, n1) : 0.f;
  }
  const float g0 = sh.gmod[lane], g1 = sh.gmod[32 + lane];  // O == 1 fast path below uses these; general O reads gmod
  (void)g0; (void)g1;
  zb[lane] = 0.f; zb[lane + 32] = 0.f;
  __syncwarp();
  float* GP = sh.gp + tid;
  float* keep = sh.gp + (C.S + C.O) * CC_CKU * 128 + tid;
  for (int e = 0; e < (C.S + C.O) * CC_CKU; ++e) GP[e * 128] = 0.f;
  float lamc[CC_TINY_S];
#pragma unroll
  for (int i = 0; i < CC_TINY_S; ++i) lamc[i] = 0.f;
  float lam0 = 0.f, lam1 = 0.f;
  float yfb[CC_TINY_O];
#pragma unroll
  for (int o = 0; o < CC_TINY_O; ++o) yfb[o] = 0.f;
  const long long yrow = (long long)p.T * M.O;
  const float os = M.out_scale;
  float n_vin[CC_TINY_O], n_yb[CC_TINY_O], n_yp[CC_TINY_O], n_xc[CC_TINY_S];
  auto fetch = [&](int k) {
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) n_vin[i] = (i < p.R && actw) ? __ldg(p.in + b * (long long)p.T * p.R + (long long)k * p.R + i) : 0.f;
#pragma unroll
    for (int o = 0; o < CC_TINY_O; ++o) {
      n_yb[o] = (o < M.O && actw) ? __ldg(p.ybar + b * yrow + (long long)k * M.O + o) : 0.f;
      n_yp[o] = (o < M.O && actw) ? p.tape[((long long)k * p.tape_rows + p.tape_yrow + o) * p.Bpad + b] : 0.f;
    }
#pragma unroll
    for (int s = 0; s < CC_TINY_S; ++s) n_xc[s] = (s < C.S && actw) ? p.tape[((long long)k * p.tape_rows + C.tape_row + s) * p.Bpad + b] : 0.f;
  };
  if (p.T > 0) fetch(p.T - 1);
  for (int k = p.T - 1; k >= 0; --k) {
    float vc[CC_TINY_O], yb[CC_TINY_O], xcv[CC_TINY_S];
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) {
      vc[i] = n_vin[i];
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o)
        if (o < M.O && i == p.R + o) vc[i] = n_yp[o];
    }
#pragma unroll
    for (int o = 0; o < CC_TINY_O; ++o) yb[o] = n_yb[o] + yfb[o];  // cotangent of y_k: loss + feedback into step k+1
#pragma unroll
    for (int s = 0; s < CC_TINY_S; ++s) xcv[s] = n_xc[s];
    if (k > 0) fetch(k - 1);
    const float tc = C.has_t ? ttab[k] : 0.f;
    // cotangent of the readout (post-step: before the integrator's reverse pass, pre-step: after it)
    float gy0 = 0.f, gy1 = 0.f;
#pragma unroll
    for (int o = 0; o < CC_TINY_O; ++o)
      if (o < M.O) { gy0 = fmaf(sh.gmod[o * 64 + lane], os * yb[o], gy0); gy1 = fmaf(sh.gmod[o * 64 + 32 + lane], os * yb[o], gy1); }
    if (!M.pre) { lam0 += gy0; lam1 += gy1; }
    // recompute the controller step (stage record kept for its reverse pass)
    float a[CC_TINY_O];
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) a[i] = 0.f;
    if (C.pre) cc_ctrl_readout_pre(CD, sh.wgc, xcv, tc, vc, a);
    CcCtrlRun crun;
    const int nstc = cc_ctrl_nstages(CD);
    cc_ctrl_begin(CD, xcv, tc, vc, crun);
    int cst = 0;
    // integrator's reverse pass: kbar vectors go through the warp's stage vector; the controller's
    // recomputation (independent of it) is issued stage by stage next to it
    float zs0 = 0.f, zs1 = 0.f, us0 = 0.f, us1 = 0.f;
    float kb0 = (nst == 4 ? h / 6.f : h) * lam0, kb1 = (nst == 4 ? h / 6.f : h) * lam1;
#pragma unroll 1
    for (int st = nst - 1; st >= 0; --st) {
      if (cst < nstc) { cc_ctrl_stage<true>(CD, sh.wfc, cst, xcv, tc, crun, keep); ++cst; }
      if (v0) zb[n0] = kb0;
      if (v1) zb[n1] = kb1;
      __syncwarp();
      float z0, z1;
      cc_wpt_dot2(c0, c1, zb, z0, z1);  // [zbar ; vbar] = kbar . [W_x ; W_v]
      __syncwarp();
      us0 += z0; us1 += z1;  // (the v entries of the sum are the input cotangent)
      if (nst == 4 && st > 0) {
        const float ca = st == 1 ? h / 6.f : h / 3.f, cb = st == 3 ? h : h / 2.f;
        zs0 += z0; zs1 += z1;
        kb0 = ca * lam0 + cb * z0;
        kb1 = ca * lam1 + cb * z1;
      } else {
        if (v0) lam0 = lam0 + (zs0 + z0);
        if (v1) lam1 = lam1 + (zs1 + z1);
      }
    }
#pragma unroll 1
    for (; cst < nstc; ++cst) cc_ctrl_stage<true>(CD, sh.wfc, cst, xcv, tc, crun, keep);
    cc_ctrl_finish<true>(CD, sh.wgc, xcv, tc, crun, a, keep);
    if (M.pre) { lam0 += gy0; lam1 += gy1; }
    if (!v0) lam0 = 0.f;
    if (!v1) lam1 = 0.f;
    // cotangent of the controller output = cotangent of the raw model input (held by the lane that owns input S + i)
    float ob[CC_TINY_O], vb[CC_TINY_O];
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) {
      ob[i] = 0.f;
      if (i < M.I) {
        const int j = M.S + i;
        const float u = __shfl_sync(CC_FULL, j < 32 ? us0 : us1, j & 31);
        ob[i] = u * cc_ut_grad(M.ut, M.u_scale, a[i]);
      }
    }
    cc_ctrl_bwd(CD, sh.wfc, sh.wgc, GP, tc, vc, keep, a, ob, lamc, vb);
#pragma unroll
    for (int o = 0; o < CC_TINY_O; ++o) {
      yfb[o] = 0.f;
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i)
        if (o < M.O && i == p.R + o) yfb[o] = vb[i];
    }
  }
  // one gradient record per trajectory (the plan sizes the workspace accordingly); lane 0's accumulators are the warp's
  if (io) {
    float* rec = p.gpart + b * p.g_rec;
    for (int row = 0; row < C.S + C.O; ++row) {
      const int goff = row < C.S ? C.f.layer[0].g_off + row * C.K0 : C.g.layer[0].g_off + (row - C.S) * C.K0;
      for (int r = 0; r < C.K0; ++r) rec[C.g_base + goff + r] = GP[(row * CC_CKU + cc_ctrl_col_of_row(C, r)) * 128];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// reverse pass of the OPEN loop with a TRAINABLE depth-0 model (ModelTrainer, step_fn.py:133-146): theta_bar and, on
// request, ubar.  Lane l owns rows l / l+32 of f's weight matrix (recomputation) AND of its gradient (register
// accumulators: the outer products need no cross-lane reduction); the transposed product zbar = W^T kbar reads the
// CTA-shared copy W[m][64] column-wise (consecutive lanes -> consecutive addresses).  One gradient record per trajectory.
// ------------------------------------------------------------------------------------------------
#define CC_WPT_OB_FLOATS(S) (CC_WPT_REGION_FLOATS + 4 * 3 * CC_WPT_ZB + (((S) + 3) / 4 * 4) * 64)  // + 3 more stage vectors per warp + W[m][64]
template <int KW>
CC_DEV void cc_wpt_dot2k(const float (&r0)[KW], const float (&r1)[KW], const float* z, float& k0, float& k1) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
  for (int q = 0; q < KW / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(z + 4 * q);
    a0 = fmaf(r0[4 * q], v.x, a0); a1 = fmaf(r0[4 * q + 1], v.y, a1); a2 = fmaf(r0[4 * q + 2], v.z, a2); a3 = fmaf(r0[4 * q + 3], v.w, a3);
    b0 = fmaf(r1[4 * q], v.x, b0); b1 = fmaf(r1[4 * q + 1], v.y, b1); b2 = fmaf(r1[4 * q + 2], v.z, b2); b3 = fmaf(r1[4 * q + 3], v.w, b3);
  }
  k0 = (a0 + a1) + (a2 + a3);
  k1 = (b0 + b1) + (b2 + b3);
}
template <int KW>
__global__ void __launch_bounds__(CC_WPT_THREADS) cc_wpt_obwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  constexpr int CM = CC_CTRL_NONE, CFG = 0;
  CC_WPT_PROLOGUE(gp)
  (void)CD;
  const CcLayerDev& F = M.f.layer[0];
  const CcLayerDev& G = M.g.layer[0];
  const float* thm = p.theta[0];
  const bool fnl = F.act != CC_ACT_IDENTITY, gnl = G.act != CC_ACT_IDENTITY;
  const int K0 = M.K0, SI = M.S + M.I;
  float* Z = sh.gp + warp * 4 * CC_WPT_ZB;            // this warp's four stage-input vectors [x ; v~ ; 1]
  float* Wsm = sh.gp + 4 * 4 * CC_WPT_ZB;              // CTA-shared W[m][64] (output m, physical input j)
  const int S4 = (M.S + 3) / 4 * 4;
  // columns permuted so that a lane's two inputs (lane, lane + 32) are adjacent: one LDS.64 per row
  for (int i = tid; i < S4 * 64; i += blockDim.x) {
    const int m = i / 64, j = i % 64;
    Wsm[m * 64 + 2 * (j & 31) + (j >> 5)] = (j < K0 && m < M.S) ? cc_weight_at(M, F, thm, m, j) : 0.f;
  }
  float r0[KW], r1[KW], wb0[KW], wb1[KW];
#pragma unroll
  for (int k = 0; k < KW; ++k) {
    r0[k] = (v0 && k < K0) ? cc_weight_at(M, F, thm, n0, k) : 0.f;
    r1[k] = (v1 && k < K0) ? cc_weight_at(M, F, thm, n1, k) : 0.f;
    wb0[k] = 0.f; wb1[k] = 0.f;
  }
  for (int s = 0; s < 4; ++s) { Z[s * CC_WPT_ZB + lane] = 0.f; Z[s * CC_WPT_ZB + lane + 32] = 0.f; }
  zb[lane] = 0.f; zb[lane + 32] = 0.f;  // kbar vector
  __syncthreads();
  if (lane == 0)
    for (int s = 0; s < 4; ++s) Z[s * CC_WPT_ZB + M.row_one] = 1.f;
  __syncwarp();
  float ga0[CC_TINY_O], ga1[CC_TINY_O], gab[CC_TINY_O];  // gradient of g: my two state columns, bias (every lane)
#pragma unroll
  for (int o = 0; o < CC_TINY_O; ++o) { ga0[o] = 0.f; ga1[o] = 0.f; gab[o] = 0.f; }
  float lam0 = 0.f, lam1 = 0.f;
  const long long yrow = (long long)(p.T + 1) * M.O;
  const float os = M.out_scale;
  const bool w0 = n0 < SI, w1 = n1 < SI;  // input index j owned for the transposed product
  // cotangent of one readout y = os * act(g . xr + b) at the state (xr0, xr1): accumulates g's gradient, returns g^T d
  auto readout_bwd = [&](float xr0, float xr1, long long yoff, float& gx0, float& gx1) {
    gx0 = 0.f; gx1 = 0.f;
    float yv[CC_TINY_O];
    if (gnl) cc_wpt_readout(sh.gmod, lane, xr0, xr1, M.O, G.act, os, yv);
#pragma unroll
    for (int o = 0; o < CC_TINY_O; ++o) {
      if (o < M.O) {
        float d = os * (actw ? __ldg(p.ybar + b * yrow + yoff + o) : 0.f);
        if (gnl) d *= cc_act_grad(G.act, os != 0.f ? yv[o] / os : 0.f);
        ga0[o] = fmaf(d, xr0, ga0[o]); ga1[o] = fmaf(d, xr1, ga1[o]); gab[o] += d;
        gx0 = fmaf(sh.gmod[o * 64 + lane], d, gx0); gx1 = fmaf(sh.gmod[o * 64 + 32 + lane], d, gx1);
      }
    }
  };
  // the step's carry (x_k from the tape) and input u_k are fetched one step ahead
  float nx0 = 0.f, nx1 = 0.f, nu = 0.f;
  auto fetch = [&](int k) {
    nx0 = (v0 && actw) ? p.tape[((long long)k * p.tape_rows + n0) * p.Bpad + b] : 0.f;
    nx1 = (v1 && actw) ? p.tape[((long long)k * p.tape_rows + n1) * p.Bpad + b] : 0.f;
    nu = (lane < M.I && actw) ? __ldg(p.in + b * (long long)p.T * M.I + (long long)k * M.I + lane) : 0.f;  // lane i < I holds input i
  };
  if (p.T > 0) fetch(p.T - 1);
  for (int k = p.T - 1; k >= 0; --k) {
    const float x0 = nx0, x1 = nx1, uraw = nu;
    if (k > 0) fetch(k - 1);
    const float ut = cc_ut(M.ut, M.u_scale, uraw);
    // ---- recompute the stages (stage inputs stay in Z[0..3], stage outputs in registers) ----
    if (v0) Z[n0] = x0;
    if (v1) Z[n1] = x1;
    if (lane < M.I)
      for (int s = 0; s < nst; ++s) Z[s * CC_WPT_ZB + M.row_v + lane] = ut;
    __syncwarp();
    float ks0[4], ks1[4];
    float s0 = 0.f, s1 = 0.f, xn0 = x0, xn1 = x1;
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      ks0[st] = 0.f; ks1[st] = 0.f;
      if (st < nst) {
        float k0, k1;
        cc_wpt_dot2k<KW>(r0, r1, Z + st * CC_WPT_ZB, k0, k1);
        if (fnl) { k0 = cc_act(F.act, k0); k1 = cc_act(F.act, k1); }
        ks0[st] = k0; ks1[st] = k1;
        if (nst == 4) {
          if (st < 3) {
            const float wk = st == 0 ? 1.f : 2.f, cz = st == 2 ? 1.f : 0.5f;
            s0 = st == 0 ? k0 : s0 + wk * k0;
            s1 = st == 0 ? k1 : s1 + wk * k1;
            if (v0) Z[(st + 1) * CC_WPT_ZB + n0] = x0 + (h * k0) * cz;
            if (v1) Z[(st + 1) * CC_WPT_ZB + n1] = x1 + (h * k1) * cz;
            __syncwarp();
          } else {
            xn0 = x0 + h * ((1.f / 6.f) * (s0 + k0));
            xn1 = x1 + h * ((1.f / 6.f) * (s1 + k1));
          }
        } else {
          xn0 = x0 + h * k0;
          xn1 = x1 + h * k1;
        }
      }
    }
    if (!v0) xn0 = 0.f;
    if (!v1) xn1 = 0.f;
    // ---- readout of this step: out[k+1] = g(x_{k+1}) (full model) or g(x_k) (compact) ----
    float gx0, gx1;
    readout_bwd(M.pre ? x0 : xn0, M.pre ? x1 : xn1, (long long)(k + 1) * M.O, gx0, gx1);
    if (!M.pre) { lam0 += gx0; lam1 += gx1; }
    // ---- integrator's reverse pass ----
    float zs0 = 0.f, zs1 = 0.f, us0 = 0.f, us1 = 0.f, z0 = 0.f, z1 = 0.f;
#pragma unroll
    for (int st = 3; st >= 0; --st) {
      if (st < nst) {
        const float ca = nst == 4 ? ((st == 3 || st == 0) ? h / 6.f : h / 3.f) : h;
        const float cb = nst == 4 ? (st == 3 ? 0.f : (st == 2 ? h : h / 2.f)) : 0.f;
        float kb0 = ca * lam0 + cb * z0, kb1 = ca * lam1 + cb * z1;
        if (fnl) { kb0 *= cc_act_grad(F.act, ks0[st]); kb1 *= cc_act_grad(F.act, ks1[st]); }
        if (!v0) kb0 = 0.f;
        if (!v1) kb1 = 0.f;
        // weight gradient rows: Wbar[n][:] += kbar[n] * [z_st ; v~ ; 1]
        const float* zst = Z + st * CC_WPT_ZB;
#pragma unroll
        for (int q = 0; q < KW / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(zst + 4 * q);
          wb0[4 * q] = fmaf(kb0, v.x, wb0[4 * q]); wb0[4 * q + 1] = fmaf(kb0, v.y, wb0[4 * q + 1]); wb0[4 * q + 2] = fmaf(kb0, v.z, wb0[4 * q + 2]); wb0[4 * q + 3] = fmaf(kb0, v.w, wb0[4 * q + 3]);
          wb1[4 * q] = fmaf(kb1, v.x, wb1[4 * q]); wb1[4 * q + 1] = fmaf(kb1, v.y, wb1[4 * q + 1]); wb1[4 * q + 2] = fmaf(kb1, v.z, wb1[4 * q + 2]); wb1[4 * q + 3] = fmaf(kb1, v.w, wb1[4 * q + 3]);
        }
        // zbar = W^T kbar (inputs j = lane, lane + 32): kbar goes through the warp's vector, W is read column-wise
        zb[n0] = kb0; zb[n1] = kb1;
        __syncwarp();
        float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll 2
        for (int m = 0; m < S4; m += 4) {  // W has S rounded up to a multiple of 4 rows (zero padded), kbar beyond S is 0
          const float4 q = *reinterpret_cast<const float4*>(zb + m);
          const float2 wa = *reinterpret_cast<const float2*>(Wsm + m * 64 + 2 * lane);
          const float2 wb = *reinterpret_cast<const float2*>(Wsm + (m + 1) * 64 + 2 * lane);
          const float2 wc = *reinterpret_cast<const float2*>(Wsm + (m + 2) * 64 + 2 * lane);
          const float2 wd = *reinterpret_cast<const float2*>(Wsm + (m + 3) * 64 + 2 * lane);
          a0 = fmaf(wa.x, q.x, a0); c0 = fmaf(wa.y, q.x, c0);
          a1 = fmaf(wb.x, q.y, a1); c1 = fmaf(wb.y, q.y, c1);
          a0 = fmaf(wc.x, q.z, a0); c0 = fmaf(wc.y, q.z, c0);
          a1 = fmaf(wd.x, q.w, a1); c1 = fmaf(wd.y, q.w, c1);
        }
        __syncwarp();
        z0 = w0 ? a0 + a1 : 0.f;
        z1 = w1 ? c0 + c1 : 0.f;
        us0 += z0; us1 += z1;
        if (nst == 4 && st > 0) { zs0 += z0; zs1 += z1; }
        else {
          if (v0) lam0 = lam0 + (zs0 + z0);
          if (v1) lam1 = lam1 + (zs1 + z1);
        }
      }
    }
    if (M.pre) { lam0 += gx0; lam1 += gx1; }
    if (!v0) lam0 = 0.f;
    if (!v1) lam1 = 0.f;
    if (p.ubar) {  // cotangent of the raw input i: held by the lane that owns input S + i; lane i knows u_i
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) {
        if (i < M.I) {
          const int j = M.S + i;
          const float uu = __shfl_sync(CC_FULL, j < 32 ? us0 : us1, j & 31);
          if (lane == i && actw) p.ubar[b * (long long)p.T * M.I + (long long)k * M.I + i] = uu * cc_ut_grad(M.ut, M.u_scale, uraw);
        }
      }
    }
  }
  // out[0] = y0 = g(x0): only g's parameters see it (x0 is not trained)
  {
    const float x00 = (p.x0 && actw && v0) ? __ldg(p.x0 + b * M.S + n0) : 0.f;
    const float x01 = (p.x0 && actw && v1) ? __ldg(p.x0 + b * M.S + n1) : 0.f;
    float gx0, gx1;
    readout_bwd(x00, x01, 0, gx0, gx1);
  }
  // gradient record of this trajectory: f's rows are lane-owned, g's columns are lane-owned
  if (actw) {
    float* rec = p.gpart + b * p.g_rec + M.g_base;
#pragma unroll
    for (int k = 0; k < KW; ++k) {
      if (k < K0) {
        if (v0) rec[F.g_off + n0 * K0 + k] = wb0[k];
        if (v1) rec[F.g_off + n1 * K0 + k] = wb1[k];
      }
    }
#pragma unroll
    for (int o = 0; o < CC_TINY_O; ++o) {
      if (o < M.O) {
        if (n0 < K0) rec[G.g_off + o * K0 + n0] = v0 ? ga0[o] : (n0 == M.row_one ? gab[o] : 0.f);
        if (n1 < K0) rec[G.g_off + o * K0 + n1] = v1 ? ga1[o] : (n1 == M.row_one ? gab[o] : 0.f);
      }
    }
  }
}
