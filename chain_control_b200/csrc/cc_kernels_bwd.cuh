// cc_kernels_bwd.cuh — BPTT of the fused rollout (K2 / K4) for sm_100a.
//
// What eqx.filter_value_and_grad derives for cc/train/step_fn.py:133-146 (open loop, gradient w.r.t.
// the model) and :249-283 (closed loop, gradient w.r.t. the controller only; the model is frozen).
// The reverse time loop reads the carry (module states, fed-back output) of every checkpointed step
// from the tape the forward kernel wrote, recomputes the RK4 stage intermediates of that step in
// shared memory where the reverse pass needs them (trainable or non-linear networks), and runs the
// reverse pass as three kinds of register-tiled SGEMMs out of shared memory:
//   forward recompute   C[traj][n]  = sum_k  A[k][traj] * Wt[k][n]
//   input gradient      C[traj][k]  = sum_n  D[n][traj] * W[n][k]
//   weight gradient     G[n][k]    += sum_traj D[n][traj] * A[k][traj]      (accumulated in smem)
// Per-CTA gradient blocks are written once at the end and reduced in fixed tile order.
#pragma once
#include "cc_kernels.cuh"

#define CC_WN 2 /* weight-gradient thread tile: 2 outputs x 8 inputs */
#define CC_WK 8

// G[n][r] += sum_traj D[n][traj] * A[r][traj]; rows are owned interleaved (n = nt + NT*i,
// r = kt + KT*j) so that the LDS.128 of a warp hit distinct bank quads.
CC_DEV void cc_wgrad(const CcKParams& p, const CcLayerDev& L, const float* D, const float* A, float* smem,
                     const CcThread& th) {
  const int NT = (L.N + CC_WN - 1) / CC_WN;
  const int KT = (L.K + CC_WK - 1) / CC_WK;
  float* G = smem + L.s_gw;
  for (int tile = th.tid; tile < NT * KT; tile += p.nthreads) {
    const int nt = tile / KT, kt = tile % KT;
    float acc[CC_WN][CC_WK];
#pragma unroll
    for (int i = 0; i < CC_WN; ++i)
#pragma unroll
      for (int j = 0; j < CC_WK; ++j) acc[i][j] = 0.f;
    const float* dptr[CC_WN];
    const float* aptr[CC_WK];
    bool nok[CC_WN], rok[CC_WK];
#pragma unroll
    for (int i = 0; i < CC_WN; ++i) {
      const int n = nt + NT * i;
      nok[i] = n < L.N;
      dptr[i] = D + (nok[i] ? n : 0) * p.BTP;
    }
#pragma unroll
    for (int j = 0; j < CC_WK; ++j) {
      const int r = kt + KT * j;
      rok[j] = r < L.K;
      aptr[j] = A + (rok[j] ? r : 0) * p.BTP;
    }
#pragma unroll 2
    for (int q = 0; q < p.BT; q += 4) {
      float4 d[CC_WN], a[CC_WK];
#pragma unroll
      for (int i = 0; i < CC_WN; ++i) d[i] = *reinterpret_cast<const float4*>(dptr[i] + q);
#pragma unroll
      for (int j = 0; j < CC_WK; ++j) a[j] = *reinterpret_cast<const float4*>(aptr[j] + q);
#pragma unroll
      for (int i = 0; i < CC_WN; ++i)
#pragma unroll
        for (int j = 0; j < CC_WK; ++j) {
          acc[i][j] = fmaf(d[i].x, a[j].x, acc[i][j]);
          acc[i][j] = fmaf(d[i].y, a[j].y, acc[i][j]);
          acc[i][j] = fmaf(d[i].z, a[j].z, acc[i][j]);
          acc[i][j] = fmaf(d[i].w, a[j].w, acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < CC_WN; ++i)
#pragma unroll
      for (int j = 0; j < CC_WK; ++j)
        if (nok[i] && rok[j]) G[(nt + NT * i) * L.Kp + kt + KT * j] += acc[i][j];
  }
}

// dst[r][traj] = (sum_n W[n][r] * D[n][traj]) * act'(H[r][traj])   for r < K  (H may be null)
// tile over (traj, r) with CGk column groups.  No trailing barrier.
template <int TN>
CC_DEV void cc_igrad(const CcKParams& p, const CcLayerDev& L, const float* D, float* dst, const float* H, int hact,
                     const float* smem, const CcThread& th) {
  if (!(th.rg_ok && th.cg < L.CGk)) return;
  float acc[CC_TM][TN];
  cc_zero_tile<TN>(acc);
  cc_gemm_tile<TN>(D + 4 * th.rg, p.BTP, smem + L.s_w, L.Kp, L.N, L.CGk, th.cg, acc);
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int r = cc_col<TN>(L.CGk, th.cg, j);
    if (r < L.K) {
      float4 v = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      if (H && r < L.in_log) {
        const float4 hv = *reinterpret_cast<const float4*>(H + r * p.BTP + 4 * th.rg);
        v.x *= cc_act_grad(hact, hv.x); v.y *= cc_act_grad(hact, hv.y);
        v.z *= cc_act_grad(hact, hv.z); v.w *= cc_act_grad(hact, hv.w);
      }
      *reinterpret_cast<float4*>(dst + r * p.BTP + 4 * th.rg) = v;
    }
  }
}

// reverse pass of one MLP.  `cur` (DA) holds the cotangent of the last layer's PRE-activation
// [N_L][BTP].  in0 = input buffer of layer 0; hid[l] = smem offset of layer l's hidden output.
// Returns the buffer holding the cotangent of the MLP input (rows < K of layer 0).  Ends with a barrier.
template <int TN>
CC_DEV float* cc_mlp_bwd(const CcKParams& p, const CcNodeDev& nd, const CcMlpDev& mlp, float* cur, float* other,
                         const float* in0, const int* hid, float* smem, const CcThread& th) {
  for (int l = mlp.L - 1; l >= 0; --l) {
    const CcLayerDev& L = mlp.layer[l];
    const float* in_l = l == 0 ? in0 : smem + hid[l - 1];
    if (nd.trainable) cc_wgrad(p, L, cur, in_l, smem, th);
    const float* H = l > 0 ? smem + hid[l - 1] : nullptr;
    cc_igrad<TN>(p, L, cur, other, H, l > 0 ? mlp.layer[l - 1].act : CC_ACT_IDENTITY, smem, th);
    __syncthreads();
    float* t = cur; cur = other; other = t;
  }
  return cur;
}

// cotangent of the readout: OB[O][BTP] -> adds to xrb tile (state part) and VB (input rows).
// gsrc = the buffer g read ([x or x_next ; t ; v ; 1]).  Ends with a barrier.
template <int TN>
CC_DEV void cc_readout_bwd(const CcKParams& p, const CcNodeDev& nd, const float* gsrc, float* smem, const CcThread& th,
                           float (&xrb)[CC_TM][TN], bool want_input_grad) {
  const CcLayerDev& LL = nd.g.layer[nd.g.L - 1];
  float* DA = smem + nd.s_DA;
  float* DB = smem + nd.s_DB;
  const float* OB = smem + nd.s_OB;
  const float* OUT = smem + nd.s_OUT;
  for (int i = th.tid; i < nd.O * p.BT; i += p.nthreads) {
    const int o = i / p.BT, traj = i % p.BT;
    float d = nd.out_scale * OB[o * p.BTP + traj];
    if (LL.act != CC_ACT_IDENTITY) {
      const float y = nd.out_scale != 0.f ? OUT[o * p.BTP + traj] / nd.out_scale : 0.f;
      d *= cc_act_grad(LL.act, y);
    }
    DA[o * p.BTP + traj] = d;
  }
  __syncthreads();
  float* cur = cc_mlp_bwd<TN>(p, nd, nd.g, DA, DB, gsrc, nd.s_GH, smem, th);
  if (want_input_grad) {
    const bool own = th.rg_ok && th.cg < nd.CGs;
    cc_load_tile<TN>(cur, p.BTP, own ? nd.S : 0, nd.CGs, th, xrb);
    if (th.tid < p.BT)
      for (int i = 0; i < nd.I; ++i) smem[nd.s_VB + i * p.BTP + th.tid] += cur[(nd.row_v + i) * p.BTP + th.tid];
  }
  __syncthreads();
}

// Reverse pass of one node step.  On entry: OB = cotangent of the node output,
// LAM = cotangent of the next state, stage intermediates recomputed (if nd.recompute) by
// cc_node_step(keep=true).  vraw[i*BTP + traj] = raw node input (for u_transform'), or null.
// On exit: LAM = cotangent of the state at step start, VB = cotangent of the raw input.
template <int TN>
CC_DEV_NOINLINE void cc_node_bwd(const CcKParams& p, const CcNodeDev& nd, float* smem, const CcThread& th, const float* vraw) {
  const float h = nd.dt;
  const bool own = th.rg_ok && th.cg < nd.CGs;
  float* X = smem + nd.s_X;
  float* LAM = smem + nd.s_LAM;
  float* DA = smem + nd.s_DA;
  float* DB = smem + nd.s_DB;
  for (int i = th.tid; i < nd.I * p.BT; i += p.nthreads) smem[nd.s_VB + (i / p.BT) * p.BTP + (i % p.BT)] = 0.f;
  __syncthreads();
  float xrb[CC_TM][TN], lam[CC_TM][TN];
  cc_zero_tile<TN>(xrb);
  cc_readout_bwd<TN>(p, nd, nd.pre ? X : smem + nd.s_ZA, smem, th, xrb, true);
  cc_load_tile<TN>(LAM, p.BTP, own ? nd.S : 0, nd.CGs, th, lam);
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
    for (int s = 3; s >= 0; --s) {
      if (own) {
        float d[CC_TM][TN];
        float ks[CC_TM][TN];
        if (FL.act != CC_ACT_IDENTITY) cc_load_tile<TN>(smem + nd.s_KS[s], p.BTP, nd.S, nd.CGs, th, ks);
#pragma unroll
        for (int i = 0; i < CC_TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            float kb;
            if (s == 3) kb = c6 * lam[i][j];
            else if (s == 2) kb = c3 * lam[i][j] + h * zb[i][j];
            else if (s == 1) kb = c3 * lam[i][j] + c2 * zb[i][j];
            else kb = c6 * lam[i][j] + c2 * zb[i][j];
            if (FL.act != CC_ACT_IDENTITY) kb *= cc_act_grad(FL.act, ks[i][j]);
            d[i][j] = kb;
          }
        cc_store_tile<TN>(DA, p.BTP, nd.S, nd.CGs, th, d);
      }
      __syncthreads();
      const float* in0 = nd.recompute ? smem + nd.s_Z[s] : X;
      float* cur = cc_mlp_bwd<TN>(p, nd, nd.f, DA, DB, in0, nd.s_HS[s], smem, th);
      cc_load_tile<TN>(cur, p.BTP, own ? nd.S : 0, nd.CGs, th, zb);
      if (th.tid < p.BT)
        for (int i = 0; i < nd.I; ++i) smem[nd.s_VB + i * p.BTP + th.tid] += cur[(nd.row_v + i) * p.BTP + th.tid];
#pragma unroll
      for (int i = 0; i < CC_TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) zsum[i][j] += zb[i][j];
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) lam[i][j] += zsum[i][j];
  } else {
    if (own) {
      float d[CC_TM][TN], ks[CC_TM][TN];
      if (FL.act != CC_ACT_IDENTITY) cc_load_tile<TN>(smem + nd.s_KS[0], p.BTP, nd.S, nd.CGs, th, ks);
#pragma unroll
      for (int i = 0; i < CC_TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          float kb = nd.integrator == CC_INT_EE ? h * lam[i][j] : lam[i][j];
          if (FL.act != CC_ACT_IDENTITY) kb *= cc_act_grad(FL.act, ks[i][j]);
          d[i][j] = kb;
        }
      cc_store_tile<TN>(DA, p.BTP, nd.S, nd.CGs, th, d);
    }
    __syncthreads();
    float* cur = cc_mlp_bwd<TN>(p, nd, nd.f, DA, DB, X, nd.s_HS[0], smem, th);
    float zb[CC_TM][TN];
    cc_load_tile<TN>(cur, p.BTP, own ? nd.S : 0, nd.CGs, th, zb);
    if (th.tid < p.BT)
      for (int i = 0; i < nd.I; ++i) smem[nd.s_VB + i * p.BTP + th.tid] += cur[(nd.row_v + i) * p.BTP + th.tid];
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) lam[i][j] = nd.integrator == CC_INT_EE ? lam[i][j] + zb[i][j] : zb[i][j];
    __syncthreads();
  }
  if (nd.pre) {
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) lam[i][j] += xrb[i][j];
  }
  if (own) cc_store_tile<TN>(LAM, p.BTP, nd.S, nd.CGs, th, lam);
  if (vraw && nd.ut != CC_UT_IDENTITY && th.tid < p.BT)
    for (int i = 0; i < nd.I; ++i)
      smem[nd.s_VB + i * p.BTP + th.tid] *= cc_ut_grad(nd.ut, nd.u_scale, vraw[i * p.BTP + th.tid]);
  __syncthreads();
}

// state rows of X <- tape record `ck`
CC_DEV void cc_tape_load(const CcKParams& p, const CcNodeDev& nd, long long b0, int ck, float* smem, int tid) {
  const float* rec = p.tape + ((long long)ck * p.tape_rows + nd.tape_row) * p.Bpad + b0;
  for (int i = tid; i < nd.S * p.BT; i += p.nthreads) {
    const int s = i / p.BT, traj = i % p.BT;
    smem[nd.s_X + s * p.BTP + traj] = (b0 + traj < p.B) ? rec[(long long)s * p.Bpad + traj] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// K2 + K4: reverse time loop
// ------------------------------------------------------------------------------------------------
template <int TN>
__global__ void __launch_bounds__(320, 1) cc_rollout_bwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  CC_KERNEL_PROLOGUE(gp)
  CcThread th;
  th.tid = threadIdx.x;
  th.rg = th.tid / p.CGT;
  th.cg = th.tid % p.CGT;
  th.rg_ok = th.rg < p.RG;
  const long long tile = blockIdx.x;
  const long long b0 = tile * p.BT;
  const bool closed = p.n_nodes == 2;
  const CcNodeDev& M = p.nd[p.n_nodes - 1];
  const CcNodeDev& C = p.nd[0];
  const float* ttab = p.tape + (long long)p.n_ckpt * p.tape_rows * p.Bpad;  // [2][T] step times

  for (int n = 0; n < p.n_nodes; ++n) {
    const CcNodeDev& nd = p.nd[n];
    cc_stage_mlp(nd, nd.f, p.theta[n], smem, th.tid, p.nthreads);
    cc_stage_mlp(nd, nd.g, p.theta[n], smem, th.tid, p.nthreads);
    cc_init_node_buffers(p, nd, smem, th.tid, true);
    for (int i = th.tid; i < nd.S * p.BTP; i += p.nthreads) smem[nd.s_LAM + i] = 0.f;
    for (int i = th.tid; i < nd.O * p.BTP; i += p.nthreads) smem[nd.s_OUT + i] = 0.f;
  }
  if (closed)
    for (int i = th.tid; i < M.O * p.BTP; i += p.nthreads) smem[p.s_yfb + i] = 0.f;
  __syncthreads();

  const int Win = closed ? p.R : M.I;
  const int nchunk = (p.T + CC_TC - 1) / CC_TC;
  for (int c = nchunk - 1; c >= 0; --c) {
    const int k0 = c * CC_TC;
    const int kend = (p.T - k0) < CC_TC ? (p.T - k0) : CC_TC;
    if (Win > 0)
      cc_stage_in(p, p.in, CC_TC * Win, (long long)p.T * Win, (long long)k0 * Win, kend * Win, b0, smem + p.s_in, th.tid);
    if (closed)
      cc_stage_in(p, p.ybar, CC_TC * M.O, (long long)p.T * M.O, (long long)k0 * M.O, kend * M.O, b0, smem + p.s_yb, th.tid);
    else
      cc_stage_in(p, p.ybar, CC_TC * M.O, (long long)(p.T + 1) * M.O, (long long)(k0 + 1) * M.O, kend * M.O, b0, smem + p.s_yb, th.tid);
    __syncthreads();
    for (int kk = kend - 1; kk >= 0; --kk) {
      const int k = k0 + kk;  // ckpt_every == 1: the tape holds the carry of every step
      if (closed) {
        const float tc = C.has_t ? ttab[k] : 0.f;
        const float tm = M.has_t ? ttab[p.T + k] : 0.f;
        cc_tape_load(p, C, b0, k, smem, th.tid);
        if (M.recompute) cc_tape_load(p, M, b0, k, smem, th.tid);
        if (th.tid < p.BT) {
          float v[16];
          for (int r = 0; r < p.R; ++r) v[r] = smem[p.s_in + (kk * p.R + r) * p.BTP + th.tid];
          for (int o = 0; o < M.O; ++o)
            v[p.R + o] = (b0 + th.tid < p.B) ? p.tape[((long long)k * p.tape_rows + C.S + M.S + o) * p.Bpad + b0 + th.tid] : 0.f;
          cc_set_inputs(p, C, smem, th.tid, v, tc, true);
        }
        __syncthreads();
        cc_node_step<TN>(p, C, smem, th, tc, true, true, false, 0, 0);  // controller always recomputes: a_k is needed
        if (th.tid < p.BT) {
          float a[16];
          for (int i = 0; i < M.I; ++i) {
            a[i] = smem[C.s_OUT + i * p.BTP + th.tid];
            smem[p.s_araw + i * p.BTP + th.tid] = a[i];
          }
          cc_set_inputs(p, M, smem, th.tid, a, tm, M.recompute != 0);
          for (int o = 0; o < M.O; ++o)
            smem[M.s_OB + o * p.BTP + th.tid] = smem[p.s_yb + (kk * M.O + o) * p.BTP + th.tid] + smem[p.s_yfb + o * p.BTP + th.tid];
        }
        __syncthreads();
        if (M.recompute) cc_node_step<TN>(p, M, smem, th, tm, true, true, false, 0, 0);
        cc_node_bwd<TN>(p, M, smem, th, smem + p.s_araw);
        for (int i = th.tid; i < M.I * p.BT; i += p.nthreads) {
          const int r = i / p.BT, traj = i % p.BT;
          smem[C.s_OB + r * p.BTP + traj] = smem[M.s_VB + r * p.BTP + traj];
        }
        __syncthreads();
        cc_node_bwd<TN>(p, C, smem, th, nullptr);
        for (int i = th.tid; i < M.O * p.BT; i += p.nthreads) {
          const int o = i / p.BT, traj = i % p.BT;
          smem[p.s_yfb + o * p.BTP + traj] = smem[C.s_VB + (p.R + o) * p.BTP + traj];
        }
        __syncthreads();
      } else {
        const float tm = M.has_t ? ttab[k] : 0.f;
        cc_tape_load(p, M, b0, k, smem, th.tid);
        if (th.tid < p.BT) {
          float v[16];
          for (int i = 0; i < M.I; ++i) v[i] = smem[p.s_in + (kk * M.I + i) * p.BTP + th.tid];
          cc_set_inputs(p, M, smem, th.tid, v, tm, M.recompute != 0);
          for (int o = 0; o < M.O; ++o) smem[M.s_OB + o * p.BTP + th.tid] = smem[p.s_yb + (kk * M.O + o) * p.BTP + th.tid];
        }
        __syncthreads();
        if (M.recompute) cc_node_step<TN>(p, M, smem, th, tm, true, true, false, 0, 0);
        cc_node_bwd<TN>(p, M, smem, th, smem + p.s_in + kk * M.I * p.BTP);
        if (p.ubar && th.tid < p.BT)
          for (int i = 0; i < M.I; ++i) smem[p.s_ub + (kk * M.I + i) * p.BTP + th.tid] = smem[M.s_VB + i * p.BTP + th.tid];
      }
    }
    __syncthreads();
    if (!closed && p.ubar)
      cc_stage_out(p, p.ubar, CC_TC * M.I, (long long)p.T * M.I, (long long)k0 * M.I, kend * M.I, b0, smem + p.s_ub, th.tid);
    __syncthreads();
  }

  if (!closed) {
    // y0 = out_scale * g([x0 ; 0]) only feeds g's parameters (x0 is not trained, neural_ode_model.py:151-158)
    for (int i = th.tid; i < M.S * p.BT; i += p.nthreads) {
      const int traj = i / M.S, s = i % M.S;
      smem[M.s_X + s * p.BTP + traj] = (p.x0 && b0 + traj < p.B) ? __ldg(p.x0 + (b0 + traj) * M.S + s) : 0.f;
    }
    if (th.tid < p.BT) {
      float v[16];
      for (int i = 0; i < M.I; ++i) v[i] = 0.f;
      cc_set_inputs(p, M, smem, th.tid, v, 0.f, false);
      // cc_set_inputs applies u_transform(0), harmless: models have no feed-through
      for (int o = 0; o < M.O; ++o)
        smem[M.s_OB + o * p.BTP + th.tid] = (b0 + th.tid < p.B) ? __ldg(p.ybar + (b0 + th.tid) * (long long)(p.T + 1) * M.O + o) : 0.f;
    }
    __syncthreads();
    cc_readout<TN>(p, M, smem + M.s_X, smem, M.s_GH, th);
    float dummy[CC_TM][TN];
    cc_readout_bwd<TN>(p, M, smem + M.s_X, smem, th, dummy, false);
  }

  // per-CTA gradient block -> workspace (reduced over tiles in fixed order by cc_reduce_grad_kernel)
  for (int n = 0; n < p.n_nodes; ++n) {
    const CcNodeDev& nd = p.nd[n];
    if (!nd.trainable) continue;
    int off = nd.gw_off;
    for (int w = 0; w < 2; ++w) {
      const CcMlpDev& mlp = w ? nd.g : nd.f;
      for (int l = 0; l < mlp.L; ++l) {
        const CcLayerDev& L = mlp.layer[l];
        const int sz = L.N * L.Kp;
        for (int i = th.tid; i < sz; i += p.nthreads) p.gpart[tile * p.gw_total + off + i] = smem[L.s_gw + i];
        off += sz;
      }
    }
  }
}
