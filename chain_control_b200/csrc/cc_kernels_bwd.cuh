// cc_kernels_bwd.cuh — BPTT of the fused rollout (K2 / K4) for sm_100a, one autonomous warp per 32
// trajectories (see cc_plan.h).
//
// What eqx.filter_value_and_grad derives for cc/train/step_fn.py:133-146 (open loop, gradient w.r.t.
// the model) and :249-283 (closed loop, gradient w.r.t. the controller only; the model is frozen).
// The reverse time loop reads the carry (module states, fed-back output) of every checkpointed step
// from the tape the forward kernel wrote, recomputes the RK4 stage intermediates of that step in
// shared memory only where the reverse pass needs them (trainable or non-linear networks; a frozen
// linear model needs none), and runs the reverse pass as register-tiled SGEMMs out of shared memory:
//   forward recompute   C[traj][n]  = sum_k  A[k][traj] * Wt[k][n]
//   input gradient      C[traj][r]  = sum_n  D[n][traj] * W[n][r]
//   weight gradient     G[n][r]    += sum_traj D[n][traj] * A[r][traj]   (per-warp accumulator in smem)
// Gradient records are written once per warp tile and reduced in fixed order (deterministic).
#pragma once
#include "cc_kernels.cuh"

// weight-gradient thread tile: WN outputs x WK inputs per lane; 4 x 8 for wide layers, smaller tiles for the narrow layers of
// the reference's default networks (width 10: a 4 x 8 tiling gives 6 tiles, i.e. 6 busy lanes of 32 -- ncu on the CLI
// architecture showed 13 k of the 21 k FFMA warp-instructions per step in this routine at 19 % lane utilisation)
template <int WN, int WK>
struct CcWFrag {
  float4 d[WN], a[WK];
};
template <int WN, int WK>
CC_DEV void cc_wfrag_load(CcWFrag<WN, WK>& f, const float* const (&dptr)[WN], const float* const (&aptr)[WK], int q) {
#pragma unroll
  for (int i = 0; i < WN; ++i) f.d[i] = *reinterpret_cast<const float4*>(dptr[i] + q);
#pragma unroll
  for (int j = 0; j < WK; ++j) f.a[j] = *reinterpret_cast<const float4*>(aptr[j] + q);
}
template <int WN, int WK>
CC_DEV void cc_wfrag_fma(const CcWFrag<WN, WK>& f, float (&acc)[WN][WK]) {
#pragma unroll
  for (int i = 0; i < WN; ++i)
#pragma unroll
    for (int j = 0; j < WK; ++j) {
      acc[i][j] = fmaf(f.d[i].x, f.a[j].x, acc[i][j]);
      acc[i][j] = fmaf(f.d[i].y, f.a[j].y, acc[i][j]);
      acc[i][j] = fmaf(f.d[i].z, f.a[j].z, acc[i][j]);
      acc[i][j] = fmaf(f.d[i].w, f.a[j].w, acc[i][j]);
    }
}

// G[n][r] += sum_{traj<32} D[n][traj] * A[r][traj].  WN x WK output tiles are dealt to the 32 lanes; each
// lane walks the 8 trajectory quads in a lane-rotated order so that the LDS.128 of a quarter-warp hit 8
// different bank quads whatever rows they touch; the operands of quad q+1 are fetched while the FMAs of
// quad q issue.  (Run-to-run deterministic; the rotated quad order depends on the owning lane, so different tile shapes
// differ in the last bit.)
template <int WN, int WK>
CC_DEV void cc_wgrad_t(const CcLayerDev& L, const float* D, const float* A, const CcWarp& w) {
  const int NT = (L.N + WN - 1) / WN;
  const int KT = (L.K + WK - 1) / WK;
  float* G = cc_ws(w) + L.w_gw;
#pragma unroll 1
  for (int tile = w.lane; tile < NT * KT; tile += 32) {
    const int nt = tile / KT, kt = tile % KT;
    float acc[WN][WK];
#pragma unroll
    for (int i = 0; i < WN; ++i)
#pragma unroll
      for (int j = 0; j < WK; ++j) acc[i][j] = 0.f;
    const float* dptr[WN];
    const float* aptr[WK];
    bool nok[WN], rok[WK];
#pragma unroll
    for (int i = 0; i < WN; ++i) {
      const int n = nt + NT * i;
      nok[i] = n < L.N;
      dptr[i] = D + (nok[i] ? n : 0) * CC_LD;
    }
#pragma unroll
    for (int j = 0; j < WK; ++j) {
      const int r = kt + KT * j;
      rok[j] = r < L.K;
      aptr[j] = A + (rok[j] ? r : 0) * CC_LD;
    }
    CcWFrag<WN, WK> cur, nxt;
    cc_wfrag_load<WN, WK>(cur, dptr, aptr, 4 * (w.lane & 7));
#pragma unroll
    for (int qq = 1; qq < 8; ++qq) {
      cc_wfrag_load<WN, WK>(nxt, dptr, aptr, 4 * ((qq + w.lane) & 7));
      cc_wfrag_fma<WN, WK>(cur, acc);
      cur = nxt;
    }
    cc_wfrag_fma<WN, WK>(cur, acc);
#pragma unroll
    for (int i = 0; i < WN; ++i)
#pragma unroll
      for (int j = 0; j < WK; ++j)
        if (nok[i] && rok[j]) G[(nt + NT * i) * L.K + kt + KT * j] += acc[i][j];
  }
}
CC_DEV void cc_wgrad(const CcLayerDev& L, const float* D, const float* A, const CcWarp& w) {
  // the largest tile that still gives (nearly) every lane work
  if (((L.N + 3) / 4) * ((L.K + 7) / 8) >= 24) cc_wgrad_t<4, 8>(L, D, A, w);
  else if (((L.N + 1) / 2) * ((L.K + 3) / 4) >= 24) cc_wgrad_t<2, 4>(L, D, A, w);
  else cc_wgrad_t<2, 2>(L, D, A, w);
}

// VB[iv][traj] += sum_n WV[iv][n] * d[traj][n]  with d the register tile of layer-0 cotangents:
// per-thread partial over its columns, then the 4 column groups meet through two shuffles.
template <int TN>
CC_DEV void cc_vgrad(const CcNodeDev& nd, const CcLayerDev& L, const float (&d)[CC_TM][TN], const CcWarp& w) {
  const float* WV = cc_cs(w) + L.s_wv;
  for (int iv = 0; iv < nd.I; ++iv) {
    float part[CC_TM] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = w.cg * L.cpt + j;
      if (j < L.cpt && n < L.N) {
        const float wv = WV[iv * L.N + n];
#pragma unroll
        for (int i = 0; i < CC_TM; ++i) part[i] = fmaf(wv, d[i][j], part[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < CC_TM; ++i) {
      part[i] += __shfl_xor_sync(CC_FULL, part[i], 8);
      part[i] += __shfl_xor_sync(CC_FULL, part[i], 16);
    }
    if (w.cg == 0) {
      float4* dst = reinterpret_cast<float4*>(cc_ws(w) + nd.w_VB + iv * CC_LD + 4 * w.rg);
      float4 v = *dst;
      v.x += part[0]; v.y += part[1]; v.z += part[2]; v.w += part[3];
      *dst = v;
    }
  }
}

// Reverse pass of one MLP.  d: cotangent of the last layer's PRE-activation (register tile, column
// mapping of that layer).  in0: input buffer of layer 0; hid[l]: offset of layer l's hidden output.
// din: cotangent of the first Kin input rows (column mapping of the source buffer).  Input-row
// cotangents are accumulated into VB when the first layer consumes v.  Ends with __syncwarp().
template <int TN>
CC_DEV void cc_mlp_bwd(const CcNodeDev& nd, const CcMlpDev& mlp, float (&d)[CC_TM][TN], const float* in0,
                       const int* hid, const CcWarp& w, float (&din)[CC_TM][TN]) {
  float* DA = cc_ws(w) + nd.w_DA;
#pragma unroll 1
  for (int l = mlp.L - 1; l >= 0; --l) {
    const CcLayerDev& L = mlp.layer[l];
    cc_store_tile<TN>(DA, L.N, L.cpt, w, d);
    __syncwarp();
    if (nd.trainable) cc_wgrad(L, DA, l == 0 ? in0 : cc_ws(w) + hid[l - 1], w);
    if (l == 0 && L.s_wv >= 0 && L.use_v) cc_vgrad<TN>(nd, L, d, w);
    cc_zero_tile<TN>(din);
    cc_gemm_tile<TN>(DA + 4 * w.rg, cc_cs(w) + L.s_w + w.cg * L.GSk, L.ldk, L.N, din);
    if (l > 0) {
      const CcLayerDev& Lp = mlp.layer[l - 1];
      float hv[CC_TM][TN];
      cc_load_tile<TN>(cc_ws(w) + hid[l - 1], Lp.N, Lp.cpt, w, hv);
#pragma unroll
      for (int i = 0; i < CC_TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) d[i][j] = din[i][j];
      cc_act_grad_tile<TN>(Lp.act, d, hv);
    }
    __syncwarp();  // DA is rewritten by the next layer
  }
}

// cotangent of the readout: OB[O][32] -> xrb (state rows) and VB (feed-through rows).
// gsrc = the buffer g read ([x or x_next ; t ; v ; 1]).
template <int TN>
CC_DEV void cc_readout_bwd(const CcNodeDev& nd, const float* gsrc, const CcWarp& w, float (&xrb)[CC_TM][TN]) {
  const CcLayerDev& LL = nd.g.layer[nd.g.L - 1];
  float d[CC_TM][TN];
  cc_load_tile<TN>(cc_ws(w) + nd.w_OB, LL.N, LL.cpt, w, d);
  if (LL.act != CC_ACT_IDENTITY) {
    float y[CC_TM][TN];
    cc_load_tile<TN>(cc_ws(w) + nd.w_OUT, LL.N, LL.cpt, w, y);
    const float inv = nd.out_scale != 0.f ? 1.f / nd.out_scale : 0.f;
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) d[i][j] *= cc_act_grad(LL.act, y[i][j] * inv);
  }
#pragma unroll
  for (int i = 0; i < CC_TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) d[i][j] *= nd.out_scale;
  cc_mlp_bwd<TN>(nd, nd.g, d, gsrc, nd.w_GH, w, xrb);
}

// Reverse pass of one tile-path node step.  On entry: OB = cotangent of the node
// output, LAM = cotangent of the next state, stage intermediates recomputed (if nd.recompute) by
// cc_node_step(keep=true).  vraw[i*32 + traj] = raw node input (for u_transform'), or null.
// On exit: LAM = cotangent of the state at step start, VB = cotangent of the raw input.
template <int TN>
CC_DEV_NOINLINE void cc_node_bwd(int node, const CcWarp w, int vraw_off) {
  const CcNodeDev& nd = cc_params().nd[node];
  const float* vraw = vraw_off >= 0 ? cc_ws(w) + vraw_off : nullptr;
  const float h = nd.dt;
  float* X = cc_ws(w) + nd.w_X;
  float* LAM = cc_ws(w) + nd.w_LAM;
  for (int i = 0; i < nd.I; ++i) cc_ws(w)[nd.w_VB + i * CC_LD + w.lane] = 0.f;
  __syncwarp();
  float xrb[CC_TM][TN], lam[CC_TM][TN];
  cc_load_tile<TN>(LAM, nd.S, nd.cptS, w, lam);  // LAM shares its buffer with DA: read it before DA is written
  __syncwarp();
  cc_readout_bwd<TN>(nd, nd.pre ? X : (nd.w_XN >= 0 ? cc_ws(w) + nd.w_XN : X), w, xrb);
  if (!nd.pre) {
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) lam[i][j] += xrb[i][j];
  }
  const CcLayerDev& FL = nd.f.layer[nd.f.L - 1];
  if (nd.integrator == CC_INT_RK4) {
    const float c6 = h / 6.f, c3 = h / 3.f, c2 = h / 2.f;
    float zsum[CC_TM][TN], zb[CC_TM][TN];
    cc_zero_tile<TN>(zsum);
    cc_zero_tile<TN>(zb);
#pragma unroll 1
    for (int s = 3; s >= 0; --s) {
      float d[CC_TM][TN];
      // k_bar_s = a_s * lam + b_s * z_bar_{s+1}     
      const float ca = (s == 3 || s == 0) ? c6 : c3;
      const float cb = s == 3 ? 0.f : (s == 2 ? h : c2);
#pragma unroll
      for (int i = 0; i < CC_TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) d[i][j] = ca * lam[i][j] + cb * zb[i][j];
      if (FL.act != CC_ACT_IDENTITY) {
        float ks[CC_TM][TN];
        cc_load_tile<TN>(cc_ws(w) + nd.w_KS[s], nd.S, nd.cptS, w, ks);
        cc_act_grad_tile<TN>(FL.act, d, ks);
      }
      cc_mlp_bwd<TN>(nd, nd.f, d, nd.recompute ? cc_ws(w) + nd.w_ZS[s] : X, nd.w_HS[s], w, zb);
#pragma unroll
      for (int i = 0; i < CC_TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) zsum[i][j] += zb[i][j];
    }
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) lam[i][j] += zsum[i][j];
  } else {
    float d[CC_TM][TN], ks[CC_TM][TN], zb[CC_TM][TN];
    if (FL.act != CC_ACT_IDENTITY) cc_load_tile<TN>(cc_ws(w) + nd.w_KS[0], nd.S, nd.cptS, w, ks);
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        float kb = nd.integrator == CC_INT_EE ? h * lam[i][j] : lam[i][j];
        if (FL.act != CC_ACT_IDENTITY) kb *= cc_act_grad(FL.act, ks[i][j]);
        d[i][j] = kb;
      }
    cc_mlp_bwd<TN>(nd, nd.f, d, X, nd.w_HS[0], w, zb);
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) lam[i][j] = nd.integrator == CC_INT_EE ? lam[i][j] + zb[i][j] : zb[i][j];
  }
  if (nd.pre) {
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) lam[i][j] += xrb[i][j];
  }
  cc_store_tile<TN>(LAM, nd.S, nd.cptS, w, lam);
  __syncwarp();
  if (vraw && nd.ut != CC_UT_IDENTITY)
    for (int i = 0; i < nd.I; ++i)
      cc_ws(w)[nd.w_VB + i * CC_LD + w.lane] *= cc_ut_grad(nd.ut, nd.u_scale, vraw[i * CC_LD + w.lane]);
  __syncwarp();
}

// state rows of X <- tape record `ck` (coalesced float4 loads through the state tile)
template <int TN>
CC_DEV void cc_tape_load(const CcKParams& p, const CcNodeDev& nd, int ck, const CcWarp& w) {
  const float* rec = p.tape + ((long long)ck * p.tape_rows + nd.tape_row) * p.Bpad + w.b0;
  float* X = cc_ws(w) + nd.w_X;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = w.cg * nd.cptS + j;
    if (j < nd.cptS && n < nd.S) {
      float4 v = *reinterpret_cast<const float4*>(rec + (long long)n * p.Bpad + 4 * w.rg);
      // trajectories beyond B carry whatever the forward pass left there: force them to zero
      const int t0 = 4 * w.rg;
      if (t0 + 0 >= w.nact) v.x = 0.f;
      if (t0 + 1 >= w.nact) v.y = 0.f;
      if (t0 + 2 >= w.nact) v.z = 0.f;
      if (t0 + 3 >= w.nact) v.w = 0.f;
      *reinterpret_cast<float4*>(X + n * CC_LD + 4 * w.rg) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// reverse pass of a "tiny" node for this lane's trajectory.  Expects cc_tiny_step<true> to have just
// run for this step (stage inputs in ZS, stage outputs in KS, readout input in ZS[0] / XN, `out`).
//   v: raw input; ob: cotangent of the output; LAM (smem, lane-private): cotangent of the next state
//   (in) / of x (out); vb: cotangent of the raw input (out).  Parameter gradients go to the
//   lane-private accumulators GP[entry][lane] (reduced over lanes once, at the end of the kernel).
// ------------------------------------------------------------------------------------------------
// G[e] += sum_{lane} D[n][lane] * IN[k][lane] for the entries e = n*K0 + k of one tiny layer: the
// outer products of all 32 trajectories are contracted across the warp (entries dealt to lanes,
// lane-rotated quad order -> conflict-free LDS.128); one owner lane per entry -> deterministic.
CC_DEV void cc_tiny_wgrad(const float* D, const float* IN, int N, int K0, float* G, int lane) {
  for (int e = lane; e < N * K0; e += 32) {
    const float* dr = D + (e / K0) * CC_LD;
    const float* ir = IN + (e % K0) * CC_LD;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four independent chains
#pragma unroll
    for (int qq = 0; qq < 8; ++qq) {
      const int q = 4 * ((qq + lane) & 7);
      const float4 a = *reinterpret_cast<const float4*>(dr + q);
      const float4 b = *reinterpret_cast<const float4*>(ir + q);
      s0 = fmaf(a.x, b.x, s0); s1 = fmaf(a.y, b.y, s1); s2 = fmaf(a.z, b.z, s2); s3 = fmaf(a.w, b.w, s3);
    }
    G[e] += (s0 + s1) + (s2 + s3);
  }
}

CC_DEV void cc_tiny_bwd(const CcNodeDev& nd, const CcWarp& w, const float (&v)[CC_TINY_O], const float (&out)[CC_TINY_O],
                        const float (&ob)[CC_TINY_O], float (&vb)[CC_TINY_O]) {
  float* wsb = cc_ws(w);
  float* ws = wsb + w.lane;
  const float* cs = cc_cs(w);
  const CcLayerDev& F = nd.f.layer[0];
  const CcLayerDev& G = nd.g.layer[0];
  const float* WF = cs + F.s_wt;
  const float* WG = cs + G.s_wt;
  float* GP = wsb + nd.w_GP;  // [g_total] per warp
  float* LAM = ws + nd.w_LAM;
  float* D = ws + nd.w_DA;
  float* ZB = ws + nd.w_OB;
  float* ZSUM = ZB + nd.S * CC_LD;
  float* XRB = ZSUM + nd.S * CC_LD;
  const float h = nd.dt;
  const int K0 = nd.K0;
  float acc[CC_TINY_K], wr[CC_TINY_K];
  float utb[CC_TINY_O];
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i) { vb[i] = 0.f; utb[i] = 0.f; }
  // readout
  const int gin_off = nd.pre ? nd.w_ZS[0] : nd.w_XN;
#pragma unroll
  for (int k = 0; k < CC_TINY_K; ++k) acc[k] = 0.f;
#pragma unroll
  for (int o = 0; o < CC_TINY_O; ++o) {
    if (o < nd.O) {
      float d = nd.out_scale * ob[o];
      if (G.act != CC_ACT_IDENTITY) d *= cc_act_grad(G.act, nd.out_scale != 0.f ? out[o] / nd.out_scale : 0.f);
      D[o * CC_LD] = d;
      cc_tiny_load_row(WG + o * CC_TINY_K, wr);
#pragma unroll
      for (int k = 0; k < CC_TINY_K; ++k) acc[k] = fmaf(wr[k], d, acc[k]);
    }
  }
  __syncwarp();
  cc_tiny_wgrad(wsb + nd.w_DA, wsb + gin_off, nd.O, K0, GP + G.g_off, w.lane);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < CC_TINY_K; ++k) {
    if (k < nd.S) XRB[k * CC_LD] = acc[k];
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i)
      if (i < nd.I && k == nd.row_v + i) vb[i] += acc[k];  // feed-through rows
  }
  if (!nd.pre)
    for (int n = 0; n < nd.S; ++n) LAM[n * CC_LD] += XRB[n * CC_LD];
  const int nst = nd.integrator == CC_INT_RK4 ? 4 : 1;
  const float c6 = h / 6.f, c3 = h / 3.f, c2 = h / 2.f;
  for (int n = 0; n < nd.S; ++n) { ZB[n * CC_LD] = 0.f; ZSUM[n * CC_LD] = 0.f; }
#pragma unroll 1
  for (int st = nst - 1; st >= 0; --st) {
#pragma unroll
    for (int k = 0; k < CC_TINY_K; ++k) acc[k] = 0.f;
    const float* KS = ws + nd.w_KS[st];
    const float ca = nd.integrator == CC_INT_RK4 ? ((st == 3 || st == 0) ? c6 : c3) : (nd.integrator == CC_INT_EE ? h : 1.f);
    const float cb = nd.integrator == CC_INT_RK4 ? (st == 3 ? 0.f : (st == 2 ? h : c2)) : 0.f;
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) {
      if (n < nd.S) {
        float kb = ca * LAM[n * CC_LD] + cb * ZB[n * CC_LD];
        if (F.act != CC_ACT_IDENTITY) kb *= cc_act_grad(F.act, KS[n * CC_LD]);
        D[n * CC_LD] = kb;
        cc_tiny_load_row(WF + n * CC_TINY_K, wr);
#pragma unroll
        for (int k = 0; k < CC_TINY_K; ++k) acc[k] = fmaf(wr[k], kb, acc[k]);
      }
    }
    __syncwarp();
    cc_tiny_wgrad(wsb + nd.w_DA, wsb + nd.w_ZS[st], nd.S, K0, GP + F.g_off, w.lane);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < CC_TINY_K; ++k) {
      if (k < nd.S) {
        ZB[k * CC_LD] = acc[k];
        ZSUM[k * CC_LD] += acc[k];
      }
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i)
        if (i < nd.I && k == nd.row_v + i) utb[i] += acc[k];
    }
  }
  for (int n = 0; n < nd.S; ++n) {
    float l = nd.integrator == CC_INT_NONE ? ZB[n * CC_LD] : LAM[n * CC_LD] + ZSUM[n * CC_LD];
    if (nd.pre) l += XRB[n * CC_LD];
    LAM[n * CC_LD] = l;
  }
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i)
    if (i < nd.I) vb[i] += utb[i] * cc_ut_grad(nd.ut, nd.u_scale, v[i]);
}

// ------------------------------------------------------------------------------------------------
// K2 + K4: reverse time loop
// ------------------------------------------------------------------------------------------------
template <int TN, int CM>
__global__ void __launch_bounds__(256, 1) cc_rollout_bwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  CC_KERNEL_PROLOGUE(gp)
  const bool closed = CM != CC_CTRL_NONE;
  const CcNodeDev& M = p.nd[closed ? 1 : 0];
  const CcNodeDev& C = p.nd[0];
  const bool m4 = cc_node_fits4(M), c4 = CM == CC_CTRL_TILE && cc_node_fits4(C);  // narrow [4][4] register tiles
  for (int n = 0; n < p.n_nodes; ++n) {
    cc_stage_mlp(p.nd[n], p.nd[n].f, p.theta[n], cta_smem, threadIdx.x, blockDim.x);
    cc_stage_mlp(p.nd[n], p.nd[n].g, p.theta[n], cta_smem, threadIdx.x, blockDim.x);
  }
  __syncthreads();
  CcWarp w;
  cc_make_warp(p, cta_smem, w);
  if (w.nact == 0) return;
  const bool act = w.lane < w.nact;
  const long long b = w.b0 + w.lane;
  const float* ttab = p.ttab;  // [2][T] step times
  for (int n = 0; n < p.n_nodes; ++n) cc_init_node_buffers(p.nd[n], w, true);
  if (closed)
    for (int o = 0; o < M.O; ++o) cc_ws(w)[p.w_yfb + o * CC_LD + w.lane] = 0.f;
  __syncwarp();
  // checkpoint segments (ckpt_every > 1): this launch walks the steps [k_beg, k_end) backwards on a segment-local tape;
  // the adjoints of the later segment and the gradient sums so far come in through lam_io / the records
  const int k_beg = p.seg_k0, k_end = p.seg_k1;
  float* rec = p.gpart + w.wt * p.g_rec;
  if (p.lam_io && k_end < p.T) {
    int row = 0;
    for (int n = 0; n < p.n_nodes; ++n) {
      const CcNodeDev& nd = p.nd[n];
      for (int s = 0; s < nd.S; ++s) cc_ws(w)[nd.w_LAM + s * CC_LD + w.lane] = act ? p.lam_io[(long long)(row + s) * p.Bpad + b] : 0.f;
      row += nd.S;
    }
    if (closed)
      for (int o = 0; o < M.O; ++o) cc_ws(w)[p.w_yfb + o * CC_LD + w.lane] = act ? p.lam_io[(long long)(row + o) * p.Bpad + b] : 0.f;
  }
  if (!p.seg_first) {
    for (int n = 0; n < p.n_nodes; ++n) {
      const CcNodeDev& nd = p.nd[n];
      if (!nd.trainable) continue;
      if (nd.tiny) {
        for (int e = w.lane; e < nd.g_total; e += 32) cc_ws(w)[nd.w_GP + e] = rec[nd.g_base + e];
      } else {
        for (int which = 0; which < 2; ++which) {
          const CcMlpDev& mlp = which ? nd.g : nd.f;
          for (int l = 0; l < mlp.L; ++l) {
            const CcLayerDev& L = mlp.layer[l];
            for (int i = w.lane; i < L.N * L.K; i += 32) cc_ws(w)[L.w_gw + i] = rec[nd.g_base + L.g_off + i];
          }
        }
      }
    }
  }
  __syncwarp();

  const int Win = closed ? p.R : M.I;
  const long long yrow = closed ? (long long)p.T * M.O : (long long)(p.T + 1) * M.O;

  // per-lane scalars of a step (input, cotangent, fed-back output, tiny-controller state) are
  // fetched one step ahead so that their global-memory latency hides behind a whole step of work
  float n_vin[CC_TINY_K], n_yb[CC_TINY_O], n_yp[CC_TINY_O], n_xc[CC_TINY_S];
  auto fetch = [&](int k) {
    const long long yoff = closed ? (long long)k * M.O : (long long)(k + 1) * M.O;
#pragma unroll
    for (int i = 0; i < CC_TINY_K; ++i) n_vin[i] = (i < Win && act) ? __ldg(p.in + b * (long long)p.T * Win + (long long)k * Win + i) : 0.f;
#pragma unroll
    for (int o = 0; o < CC_TINY_O; ++o) {
      n_yb[o] = (o < M.O && act) ? __ldg(p.ybar + b * yrow + yoff + o) : 0.f;
      n_yp[o] = (closed && o < M.O && act) ? p.tape[((long long)(k - k_beg) * p.tape_rows + p.tape_yrow + o) * p.Bpad + b] : 0.f;
    }
#pragma unroll
    for (int s = 0; s < CC_TINY_S; ++s)
      n_xc[s] = (CM == CC_CTRL_TINY && s < C.S && act) ? p.tape[((long long)(k - k_beg) * p.tape_rows + C.tape_row + s) * p.Bpad + b] : 0.f;
  };
  if (k_end > k_beg) fetch(k_end - 1);
  for (int k = k_end - 1; k >= k_beg; --k) {  // the (segment-local) tape holds the carry of every step
    float vin[CC_TINY_K], yb[CC_TINY_O], ypv[CC_TINY_O], xcv[CC_TINY_S];
#pragma unroll
    for (int i = 0; i < CC_TINY_K; ++i) vin[i] = n_vin[i];
#pragma unroll
    for (int o = 0; o < CC_TINY_O; ++o) { yb[o] = n_yb[o]; ypv[o] = n_yp[o]; }
#pragma unroll
    for (int s = 0; s < CC_TINY_S; ++s) xcv[s] = n_xc[s];
    if (k > k_beg) fetch(k - 1);
    if (closed) {
      const float tc = C.has_t ? ttab[k] : 0.f;
      const float tm = M.has_t ? ttab[p.T + k] : 0.f;
      // carry of step k: (x_c, x_m, y_prev)
      float vc[CC_TINY_K];
#pragma unroll
      for (int i = 0; i < CC_TINY_K; ++i) vc[i] = vin[i];
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o)
#pragma unroll
        for (int i = 0; i < CC_TINY_K; ++i)
          if (o < M.O && i == p.R + o) vc[i] = ypv[o];
      float a[CC_TINY_O];
      float vcc[CC_TINY_O];
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) vcc[i] = vc[i];
      if (CM == CC_CTRL_TINY) {
#pragma unroll
        for (int s = 0; s < CC_TINY_S; ++s)
          if (s < C.S) cc_ws(w)[C.w_X + s * CC_LD + w.lane] = xcv[s];
        cc_tiny_step<true>(C, w, tc, vcc, a);  // recompute with the stage intermediates kept; a_k feeds u_transform'
      } else {
        cc_tape_load<8>(p, C, k - k_beg, w);
        cc_set_inputs(C, w, vc, tc, 1);
        __syncwarp();
        CC_TILE_DISPATCH(8, c4, cc_node_step, 0, w, tc, true, true, false, 0);
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i) a[i] = (i < M.I) ? cc_ws(w)[C.w_OUT + i * CC_LD + w.lane] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i)
        if (i < M.I) cc_ws(w)[p.w_araw + i * CC_LD + w.lane] = a[i];
      if (M.recompute) {
        cc_tape_load<TN>(p, M, k - k_beg, w);
        float am[CC_TINY_K];
#pragma unroll
        for (int i = 0; i < CC_TINY_K; ++i) am[i] = (i < CC_TINY_O) ? a[i < CC_TINY_O ? i : 0] : 0.f;
        cc_set_inputs(M, w, am, tm, 1);
      }
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o)
        if (o < M.O) cc_ws(w)[M.w_OB + o * CC_LD + w.lane] = yb[o] + cc_ws(w)[p.w_yfb + o * CC_LD + w.lane];
      __syncwarp();
      if (M.recompute) CC_TILE_DISPATCH(TN, m4, cc_node_step, closed ? 1 : 0, w, tm, true, true, false, 0);
      CC_TILE_DISPATCH(TN, m4, cc_node_bwd, closed ? 1 : 0, w, p.w_araw);
      // cotangent of the controller output = cotangent of the model input
      if (CM == CC_CTRL_TINY) {
        float ob[CC_TINY_O], vb[CC_TINY_O];
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i) ob[i] = (i < M.I) ? cc_ws(w)[M.w_VB + i * CC_LD + w.lane] : 0.f;
        cc_tiny_bwd(C, w, vcc, a, ob, vb);
        for (int o = 0; o < M.O; ++o) {
          float fb = 0.f;
#pragma unroll
          for (int i = 0; i < CC_TINY_O; ++i)
            if (i == p.R + o) fb = vb[i];
          cc_ws(w)[p.w_yfb + o * CC_LD + w.lane] = fb;
        }
      } else {
        for (int i = 0; i < M.I; ++i) cc_ws(w)[C.w_OB + i * CC_LD + w.lane] = cc_ws(w)[M.w_VB + i * CC_LD + w.lane];
        __syncwarp();
        CC_TILE_DISPATCH(8, c4, cc_node_bwd, 0, w, -1);
        for (int o = 0; o < M.O; ++o) cc_ws(w)[p.w_yfb + o * CC_LD + w.lane] = cc_ws(w)[C.w_VB + (p.R + o) * CC_LD + w.lane];
      }
      __syncwarp();
    } else {
      const float tm = M.has_t ? ttab[k] : 0.f;
      cc_tape_load<TN>(p, M, k - k_beg, w);
      cc_set_inputs(M, w, vin, tm, 1);
#pragma unroll
      for (int i = 0; i < CC_TINY_K; ++i)
        if (i < M.I) cc_ws(w)[p.w_araw + i * CC_LD + w.lane] = vin[i];
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o)
        if (o < M.O) cc_ws(w)[M.w_OB + o * CC_LD + w.lane] = yb[o];
      __syncwarp();
      if (M.recompute) CC_TILE_DISPATCH(TN, m4, cc_node_step, closed ? 1 : 0, w, tm, true, true, false, 0);
      CC_TILE_DISPATCH(TN, m4, cc_node_bwd, closed ? 1 : 0, w, p.w_araw);
      if (p.ubar && act)
        for (int i = 0; i < M.I; ++i) p.ubar[b * (long long)p.T * M.I + (long long)k * M.I + i] = cc_ws(w)[M.w_VB + i * CC_LD + w.lane];
    }
  }

  if (p.lam_io && k_beg > 0) {  // adjoint carry for the earlier segment
    __syncwarp();
    int row = 0;
    for (int n = 0; n < p.n_nodes; ++n) {
      const CcNodeDev& nd = p.nd[n];
      if (act)
        for (int s = 0; s < nd.S; ++s) p.lam_io[(long long)(row + s) * p.Bpad + b] = cc_ws(w)[nd.w_LAM + s * CC_LD + w.lane];
      row += nd.S;
    }
    if (closed && act)
      for (int o = 0; o < M.O; ++o) p.lam_io[(long long)(row + o) * p.Bpad + b] = cc_ws(w)[p.w_yfb + o * CC_LD + w.lane];
  }
  if (!closed && k_beg == 0) {
    // y0 = out_scale * g([x0 ; 0]) only feeds g's parameters (x0 is not trained, neural_ode_model.py:151-158)
    for (int s = 0; s < M.S; ++s) cc_ws(w)[M.w_X + s * CC_LD + w.lane] = (p.x0 && act) ? __ldg(p.x0 + b * M.S + s) : 0.f;
    float v0[CC_TINY_K];
#pragma unroll
    for (int i = 0; i < CC_TINY_K; ++i) v0[i] = 0.f;
    cc_set_inputs(M, w, v0, M.t0, 2);  // u_transform(0) in the v rows is harmless: models have no feed-through
    for (int o = 0; o < M.O; ++o) cc_ws(w)[M.w_OB + o * CC_LD + w.lane] = act ? __ldg(p.ybar + b * yrow + o) : 0.f;
    __syncwarp();
    cc_readout<TN>(0, M.w_X, true, w);
    float dummy[CC_TM][TN];
    cc_readout_bwd<TN>(M, cc_ws(w) + M.w_X, w, dummy);
  }

  // gradient record of this warp tile -> workspace (reduced over tiles in fixed order afterwards)
  __syncwarp();
  for (int n = 0; n < p.n_nodes; ++n) {
    const CcNodeDev& nd = p.nd[n];
    if (!nd.trainable) continue;
    if (nd.tiny) {
      for (int e = w.lane; e < nd.g_total; e += 32) rec[nd.g_base + e] = cc_ws(w)[nd.w_GP + e];
    } else {
      for (int which = 0; which < 2; ++which) {
        const CcMlpDev& mlp = which ? nd.g : nd.f;
        for (int l = 0; l < mlp.L; ++l) {
          const CcLayerDev& L = mlp.layer[l];
          for (int i = w.lane; i < L.N * L.K; i += 32) rec[nd.g_base + L.g_off + i] = cc_ws(w)[L.w_gw + i];
        }
      }
    }
  }
}
