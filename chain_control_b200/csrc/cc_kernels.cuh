// cc_kernels.cuh — device code of the fused neural-ODE rollout (forward) for sm_100a.
//
// One persistent CTA per trajectory tile runs the whole time loop: weights are staged once in
// shared memory, the state tile and the RK4 stage sum live in registers, every layer evaluation is
// a register-tiled FP32 SGEMM out of shared memory (see cc_plan.h for the layout).
//
// Restates (no code shared with the reference, which is JAX):
//   AbstractModel.unroll            cc/core/abstract.py:54-70
//   AbstractController.unroll       cc/core/abstract.py:97-127
//   NeuralOdeModel.step / y0        cc/examples/neural_ode_model.py:106-149,160-175 (+ compact :94-110)
//   NeuralOdeController.step        cc/examples/neural_ode_controller.py:108-149 (+ compact :106-129)
//   integrate                       cc/examples/nn_lib/integrate.py:1-28
//   mlp_network                     cc/examples/nn_lib/mlp_network.py:5-35
#pragma once
#include "cc_plan.h"

#ifndef CC_EMULATE
#define CC_DEV __device__ __forceinline__
#else
#define CC_DEV static inline
#endif

// ------------------------------------------------------------------------------------------------
// elementwise pieces
// ------------------------------------------------------------------------------------------------
CC_DEV float cc_act(int kind, float x) {
  if (kind == CC_ACT_RELU) return x > 0.f ? x : 0.f;
  if (kind == CC_ACT_TANH) return tanhf(x);
  return x;
}
// derivative expressed through the activation OUTPUT (relu: 0 at exactly 0, like JAX)
CC_DEV float cc_act_grad(int kind, float y) {
  if (kind == CC_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (kind == CC_ACT_TANH) return 1.f - y * y;
  return 1.f;
}
CC_DEV float cc_ut(int kind, float scale, float u) {
  if (kind == CC_UT_ARCTAN) return atanf(u);
  if (kind == CC_UT_KTANH) return scale * tanhf(u / scale);
  return u;
}
CC_DEV float cc_ut_grad(int kind, float scale, float u) {
  if (kind == CC_UT_ARCTAN) return 1.f / (1.f + u * u);
  if (kind == CC_UT_KTANH) {
    const float th = tanhf(u / scale);
    return 1.f - th * th;
  }
  return 1.f;
}

struct CcThread {
  int tid, rg, cg;
  bool rg_ok;
};

// physical row of a first-layer source buffer -> logical input column of the MLP
// (>= 0), -1 = not consumed (zero weight), -2 = bias row.
CC_DEV int cc_row_to_logical(const CcNodeDev& nd, const CcLayerDev& L, int r) {
  if (!L.first) return r < L.in_log ? r : (r == L.in_log ? -2 : -1);
  if (r < nd.S) return r;
  r -= nd.S;
  if (nd.has_t) {
    if (r == 0) return L.use_t ? nd.S : -1;
    r -= 1;
  }
  if (r < nd.I) return L.use_v ? nd.S + L.use_t + r : -1;
  r -= nd.I;
  return r == 0 ? -2 : -1;
}

CC_DEV float cc_weight_at(const CcNodeDev& nd, const CcLayerDev& L, const float* theta, int n, int r) {
  if (n >= L.N || r >= L.K) return 0.f;
  const int lc = cc_row_to_logical(nd, L, r);
  if (lc >= 0) return __ldg(theta + L.w_off + n * L.in_log + lc);
  if (lc == -2 && L.b_off >= 0) return __ldg(theta + L.b_off + n);
  return 0.f;
}

// stage the weights of one MLP: Wt[K][Np] always, W[N][Kp] and zeroed Wbar[N][Kp] when present
CC_DEV void cc_stage_mlp(const CcNodeDev& nd, const CcMlpDev& mlp, const float* theta, float* smem, int tid,
                         int nthreads) {
  for (int l = 0; l < mlp.L; ++l) {
    const CcLayerDev& L = mlp.layer[l];
    float* Wt = smem + L.s_wt;
    for (int i = tid; i < L.K * L.Np; i += nthreads) Wt[i] = cc_weight_at(nd, L, theta, i % L.Np, i / L.Np);
    if (L.s_w >= 0) {
      float* W = smem + L.s_w;
      for (int i = tid; i < L.N * L.Kp; i += nthreads) W[i] = cc_weight_at(nd, L, theta, i / L.Kp, i % L.Kp);
    }
    if (L.s_gw >= 0) {
      float* G = smem + L.s_gw;
      for (int i = tid; i < L.N * L.Kp; i += nthreads) G[i] = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// register-tiled SGEMM out of shared memory:  acc[i][j] += sum_k A[k][4rg+i] * Bm[k][col(j)]
//   col(j) = 4cg + j (j<4) | 4CG + 4cg + (j-4)   -> every LDS.128 of a warp is conflict-free
// ------------------------------------------------------------------------------------------------
CC_DEV void cc_gemm_tile(const float* A, int lda, const float* Bm, int ldb, int K, int cgoff,
                         float (&acc)[CC_TM][CC_TN]) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(A + k * lda);
    const float4 b0 = *reinterpret_cast<const float4*>(Bm + k * ldb);
    const float4 b1 = *reinterpret_cast<const float4*>(Bm + k * ldb + cgoff);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < CC_TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

CC_DEV int cc_col(int CG, int cg, int j) { return j < 4 ? 4 * cg + j : 4 * CG + 4 * cg + (j - 4); }

CC_DEV void cc_zero_tile(float (&t)[CC_TM][CC_TN]) {
#pragma unroll
  for (int i = 0; i < CC_TM; ++i)
#pragma unroll
    for (int j = 0; j < CC_TN; ++j) t[i][j] = 0.f;
}

// tile <-> feature-major buffer rows [n][BTP]
CC_DEV void cc_load_tile(const float* buf, int BTP, int N, int CG, const CcThread& th, float (&t)[CC_TM][CC_TN]) {
#pragma unroll
  for (int j = 0; j < CC_TN; ++j) {
    const int n = cc_col(CG, th.cg, j);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) v = *reinterpret_cast<const float4*>(buf + n * BTP + 4 * th.rg);
    t[0][j] = v.x; t[1][j] = v.y; t[2][j] = v.z; t[3][j] = v.w;
  }
}
CC_DEV void cc_store_tile(float* buf, int BTP, int N, int CG, const CcThread& th, const float (&t)[CC_TM][CC_TN]) {
#pragma unroll
  for (int j = 0; j < CC_TN; ++j) {
    const int n = cc_col(CG, th.cg, j);
    if (n < N) *reinterpret_cast<float4*>(buf + n * BTP + 4 * th.rg) = make_float4(t[0][j], t[1][j], t[2][j], t[3][j]);
  }
}

// one layer as a tile GEMM; leaves the raw accumulators (bias included via the 1-row) in acc
CC_DEV void cc_layer_gemm(const CcKParams& p, const CcLayerDev& L, const float* src, const float* smem,
                          const CcThread& th, float (&acc)[CC_TM][CC_TN]) {
  cc_zero_tile(acc);
  if (th.rg_ok && th.cg < L.CG)
    cc_gemm_tile(src + 4 * th.rg, p.BTP, smem + L.s_wt + 4 * th.cg, L.Np, L.K, 4 * L.CG, acc);
}

// split-K path for very narrow layers (N <= CC_MAX_SKINNY): every thread sums a strided slice of K
// for its 4 trajectories, partials meet in P[cg][n][BTP], then a fixed-order reduction.
// dst[n][traj] = scale * act(sum).  Ends with a barrier.
CC_DEV void cc_layer_skinny(const CcKParams& p, const CcLayerDev& L, const float* src, float* smem, float* P,
                            float* dst, float scale, const CcThread& th) {
  float part[CC_MAX_SKINNY][CC_TM];
#pragma unroll
  for (int n = 0; n < CC_MAX_SKINNY; ++n)
#pragma unroll
    for (int i = 0; i < CC_TM; ++i) part[n][i] = 0.f;
  const float* Wt = smem + L.s_wt;
  if (th.rg_ok) {
    for (int k = th.cg; k < L.K; k += p.CGT) {
      const float4 a = *reinterpret_cast<const float4*>(src + k * p.BTP + 4 * th.rg);
#pragma unroll
      for (int n = 0; n < CC_MAX_SKINNY; ++n) {
        if (n < L.N) {
          const float w = Wt[k * L.Np + n];
          part[n][0] = fmaf(a.x, w, part[n][0]);
          part[n][1] = fmaf(a.y, w, part[n][1]);
          part[n][2] = fmaf(a.z, w, part[n][2]);
          part[n][3] = fmaf(a.w, w, part[n][3]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < CC_MAX_SKINNY; ++n)
      if (n < L.N)
        *reinterpret_cast<float4*>(P + (th.cg * CC_MAX_SKINNY + n) * p.BTP + 4 * th.rg) =
            make_float4(part[n][0], part[n][1], part[n][2], part[n][3]);
  }
  __syncthreads();
  for (int idx = th.tid; idx < L.N * p.RG; idx += p.nthreads) {
    const int n = idx / p.RG, r = idx % p.RG;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < p.CGT; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(P + (c * CC_MAX_SKINNY + n) * p.BTP + 4 * r);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x = scale * cc_act(L.act, s.x);
    s.y = scale * cc_act(L.act, s.y);
    s.z = scale * cc_act(L.act, s.z);
    s.w = scale * cc_act(L.act, s.w);
    *reinterpret_cast<float4*>(dst + n * p.BTP + 4 * r) = s;
  }
  __syncthreads();
}

// hidden layers of an MLP (all but the last): src -> hidden buffers.  returns the input buffer of
// the last layer.  hbuf[l] overrides the hidden buffer of layer l (backward keeps them per stage).
CC_DEV const float* cc_mlp_hidden(const CcKParams& p, const CcMlpDev& mlp, const float* src, float* smem,
                                  const int* hbuf, const CcThread& th) {
  float acc[CC_TM][CC_TN];
  for (int l = 0; l < mlp.L - 1; ++l) {
    const CcLayerDev& L = mlp.layer[l];
    float* dst = smem + (hbuf ? hbuf[l] : L.s_out);
    if (L.N <= CC_MAX_SKINNY) {
      cc_layer_skinny(p, L, src, smem, smem + p.nd[0].s_P, dst, 1.f, th);
    } else {
      cc_layer_gemm(p, L, src, smem, th, acc);
      if (th.rg_ok && th.cg < L.CG) {
#pragma unroll
        for (int i = 0; i < CC_TM; ++i)
#pragma unroll
          for (int j = 0; j < CC_TN; ++j) acc[i][j] = cc_act(L.act, acc[i][j]);
        cc_store_tile(dst, p.BTP, L.N, L.CG, th, acc);
      }
      __syncthreads();
    }
    src = dst;
  }
  return src;
}

// readout: OUT = out_scale * g(src rows).  Ends with a barrier.
CC_DEV void cc_readout(const CcKParams& p, const CcNodeDev& nd, const float* src, float* smem, const int* ghbuf,
                       const CcThread& th) {
  const float* in = cc_mlp_hidden(p, nd.g, src, smem, ghbuf, th);
  const CcLayerDev& L = nd.g.layer[nd.g.L - 1];
  float* OUT = smem + nd.s_OUT;
  if (L.N <= CC_MAX_SKINNY) {
    cc_layer_skinny(p, L, in, smem, smem + p.nd[0].s_P, OUT, nd.out_scale, th);
  } else {
    float acc[CC_TM][CC_TN];
    cc_layer_gemm(p, L, in, smem, th, acc);
    if (th.rg_ok && th.cg < L.CG) {
#pragma unroll
      for (int i = 0; i < CC_TM; ++i)
#pragma unroll
        for (int j = 0; j < CC_TN; ++j) acc[i][j] = nd.out_scale * cc_act(L.act, acc[i][j]);
      cc_store_tile(OUT, p.BTP, L.N, L.CG, th, acc);
    }
    __syncthreads();
  }
}

// k = f(src): hidden layers to smem, last layer into registers (activation applied)
CC_DEV void cc_rhs(const CcKParams& p, const CcNodeDev& nd, const float* src, float* smem, const int* hbuf,
                   const CcThread& th, float (&k)[CC_TM][CC_TN]) {
  const float* in = cc_mlp_hidden(p, nd.f, src, smem, hbuf, th);
  const CcLayerDev& L = nd.f.layer[nd.f.L - 1];
  cc_layer_gemm(p, L, in, smem, th, k);
  if (L.act != CC_ACT_IDENTITY) {
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < CC_TN; ++j) k[i][j] = cc_act(L.act, k[i][j]);
  }
}

// write the node's state tile into the tape record of checkpoint `ck`
CC_DEV void cc_tape_store(const CcKParams& p, const CcNodeDev& nd, long long tile, int ck, const CcThread& th,
                          const float (&x)[CC_TM][CC_TN]) {
  float* rec = p.tape + ((long long)ck * p.tape_rows + nd.tape_row) * p.Bpad + tile * p.BT;
  if (th.rg_ok && th.cg < nd.CGs) {
#pragma unroll
    for (int j = 0; j < CC_TN; ++j) {
      const int n = cc_col(nd.CGs, th.cg, j);
      if (n < nd.S) *reinterpret_cast<float4*>(rec + (long long)n * p.Bpad + 4 * th.rg) = make_float4(x[0][j], x[1][j], x[2][j], x[3][j]);
    }
  }
}

// One step of a node.  On entry X holds [x ; t ; v ; 1], ZA/ZB hold the same t(+h/2)/v/1 rows.
// On exit X holds x_next, OUT holds the readout.  `stage_bufs`: when non-null (backward recompute)
// the stage inputs are kept in s_Z[1..3] instead of the ZA/ZB ping-pong and hidden activations /
// k_i are kept per stage.  Ends with a barrier.
CC_DEV void cc_node_step(const CcKParams& p, const CcNodeDev& nd, float* smem, const CcThread& th, float t,
                         bool keep, bool do_readout, bool write_tape, long long tile, int ck) {
  float* X = smem + nd.s_X;
  const float h = nd.dt;
  const bool own = th.rg_ok && th.cg < nd.CGs;
  if (nd.pre && do_readout) cc_readout(p, nd, X, smem, keep ? nd.s_GH : nullptr, th);
  float xr[CC_TM][CC_TN], k[CC_TM][CC_TN];
  cc_load_tile(X, p.BTP, own ? nd.S : 0, nd.CGs, th, xr);
  if (write_tape) cc_tape_store(p, nd, tile, ck, th, xr);
  if (nd.integrator == CC_INT_RK4) {
    float ks[CC_TM][CC_TN];
    for (int s = 0; s < 4; ++s) {
      float* src;
      float* dst;
      if (keep) {
        src = smem + nd.s_Z[s];
        dst = s < 3 ? smem + nd.s_Z[s + 1] : X;
      } else {
        src = s == 0 ? X : (s == 2 ? smem + nd.s_ZB : smem + nd.s_ZA);
        dst = s == 0 ? smem + nd.s_ZA : (s == 1 ? smem + nd.s_ZB : (s == 2 ? smem + nd.s_ZA : X));
        // stage 3 reads ZA again, now at time t + h (ZA is idle during stage 2)
        if (s == 2 && nd.has_t && th.tid < p.BT) smem[nd.s_ZA + nd.row_t * p.BTP + th.tid] = t + h;
      }
      cc_rhs(p, nd, src, smem, keep ? nd.s_HS[s] : nullptr, th, k);
      if (own) {
        if (keep && nd.s_KS[s] >= 0) cc_store_tile(smem + nd.s_KS[s], p.BTP, nd.S, nd.CGs, th, k);
#pragma unroll
        for (int i = 0; i < CC_TM; ++i)
#pragma unroll
          for (int j = 0; j < CC_TN; ++j) {
            const float kk = k[i][j];
            if (s == 0) {
              ks[i][j] = kk;
              k[i][j] = xr[i][j] + h * kk / 2.f;
            } else if (s == 1) {
              ks[i][j] = ks[i][j] + 2.f * kk;
              k[i][j] = xr[i][j] + h * kk / 2.f;
            } else if (s == 2) {
              ks[i][j] = ks[i][j] + 2.f * kk;
              k[i][j] = xr[i][j] + h * kk;
            } else {
              k[i][j] = xr[i][j] + h * ((1.f / 6.f) * (ks[i][j] + kk));
            }
          }
        if (!(keep && s == 3)) cc_store_tile(dst, p.BTP, nd.S, nd.CGs, th, k);
      }
      __syncthreads();
    }
  } else {
    cc_rhs(p, nd, X, smem, keep ? nd.s_HS[0] : nullptr, th, k);
    if (own) {
      if (keep && nd.s_KS[0] >= 0) cc_store_tile(smem + nd.s_KS[0], p.BTP, nd.S, nd.CGs, th, k);
#pragma unroll
      for (int i = 0; i < CC_TM; ++i)
#pragma unroll
        for (int j = 0; j < CC_TN; ++j)
          k[i][j] = nd.integrator == CC_INT_EE ? xr[i][j] + h * k[i][j] : k[i][j];
    }
    __syncthreads();  // everyone finished reading X
    if (own && !keep) cc_store_tile(X, p.BTP, nd.S, nd.CGs, th, k);
    __syncthreads();
  }
  if (!nd.pre && do_readout) {
    // post-step readout g([x_next ; t ; v]); in `keep` mode X must stay x (the reverse pass needs
    // it), so x_next goes to ZA's state rows, whose t / v / 1 rows equal X's up to t (set by caller)
    if (keep) {
      float* XN = smem + nd.s_ZA;
      if (own) cc_store_tile(XN, p.BTP, nd.S, nd.CGs, th, k);
      __syncthreads();
      cc_readout(p, nd, XN, smem, nd.s_GH, th);
    } else {
      cc_readout(p, nd, X, smem, nullptr, th);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// I/O staging: global [B][Tlen][W] (trajectory-major, as the reference lays it out) <-> smem
// [e][BTP] with e = (step - k0) * W + w.  Global accesses are coalesced along e.
// ------------------------------------------------------------------------------------------------
CC_DEV void cc_stage_in(const CcKParams& p, const float* g, int ne, long long rowlen, long long eoff, int nvalid,
                        long long b0, float* s, int tid) {
  for (int idx = tid; idx < p.BT * ne; idx += p.nthreads) {
    const int traj = idx / ne, e = idx % ne;
    float v = 0.f;
    if (b0 + traj < p.B && e < nvalid) v = __ldg(g + (b0 + traj) * rowlen + eoff + e);
    s[e * p.BTP + traj] = v;
  }
}
CC_DEV void cc_stage_out(const CcKParams& p, float* g, int ne, long long rowlen, long long eoff, int nvalid,
                         long long b0, const float* s, int tid) {
  for (int idx = tid; idx < p.BT * ne; idx += p.nthreads) {
    const int traj = idx / ne, e = idx % ne;
    if (b0 + traj < p.B && e < nvalid) g[(b0 + traj) * rowlen + eoff + e] = s[e * p.BTP + traj];
  }
}

// zero the stage buffers, set the constant-1 rows
CC_DEV void cc_init_node_buffers(const CcKParams& p, const CcNodeDev& nd, float* smem, int tid, bool bwd) {
  const int nbuf = 3;
  const int bufs[3] = {nd.s_X, nd.s_ZA, nd.s_ZB};
  for (int b = 0; b < nbuf; ++b) {
    float* buf = smem + bufs[b];
    for (int i = tid; i < nd.K0 * p.BTP; i += p.nthreads) buf[i] = (i / p.BTP == nd.row_one) ? 1.f : 0.f;
  }
  if (bwd) {
    for (int s = 1; s < 4; ++s) {
      if (nd.s_Z[s] < 0) continue;
      float* buf = smem + nd.s_Z[s];
      for (int i = tid; i < nd.K0 * p.BTP; i += p.nthreads) buf[i] = (i / p.BTP == nd.row_one) ? 1.f : 0.f;
    }
  }
  for (int which = 0; which < 2; ++which) {
    const CcMlpDev& mlp = which ? nd.g : nd.f;
    for (int l = 0; l < mlp.L - 1; ++l) {
      const CcLayerDev& L = mlp.layer[l];
      const int rows = L.N + 1;
      const int nst = bwd ? (which ? 1 : nd.n_stages) : 1;
      for (int s = 0; s < nst; ++s) {
        const int off = bwd ? (which ? nd.s_GH[l] : nd.s_HS[s][l]) : L.s_out;
        float* buf = smem + off;
        for (int i = tid; i < rows * p.BTP; i += p.nthreads) buf[i] = (i / p.BTP == L.N) ? 1.f : 0.f;
      }
    }
  }
}

// set the t / v rows of a node's stage buffers for the coming step (thread `traj` of the tile)
CC_DEV void cc_set_inputs(const CcKParams& p, const CcNodeDev& nd, float* smem, int traj, const float* v, float t,
                          bool keep) {
  const float h = nd.dt;
  float* X = smem + nd.s_X;
  float* B1 = smem + (keep ? nd.s_Z[1] : nd.s_ZA);
  float* B2 = smem + (keep ? nd.s_Z[2] : nd.s_ZB);
  for (int i = 0; i < nd.I; ++i) {
    const float ut = cc_ut(nd.ut, nd.u_scale, v[i]);
    const int o = (nd.row_v + i) * p.BTP + traj;
    X[o] = ut;
    if (nd.integrator == CC_INT_RK4) {
      B1[o] = ut;
      B2[o] = ut;
      if (keep) smem[nd.s_Z[3] + o] = ut;
    }
    if (keep && !nd.pre) smem[nd.s_ZA + o] = ut;  // post-step readout source in keep mode
  }
  if (nd.has_t) {
    const int o = nd.row_t * p.BTP + traj;
    X[o] = t;
    if (nd.integrator == CC_INT_RK4) {
      B1[o] = t + h / 2.f;
      B2[o] = t + h / 2.f;
      if (keep) smem[nd.s_Z[3] + o] = t + h;
    }
    if (keep && !nd.pre) smem[nd.s_ZA + o] = t;
  }
}

#ifndef CC_EMULATE
#define CC_SMEM_DECL extern __shared__ __align__(16) float cc_smem[]; float* smem = cc_smem;
#else
#define CC_SMEM_DECL float* smem = ccemu::st().smem;
#endif

// ------------------------------------------------------------------------------------------------
// K1 + K3: forward rollout (open loop: n_nodes == 1; closed loop: nd[0] controller, nd[1] model)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) cc_rollout_fwd_kernel(const CcKParams p) {
  CC_SMEM_DECL
  CcThread th;
  th.tid = threadIdx.x;
  th.rg = th.tid / p.CGT;
  th.cg = th.tid % p.CGT;
  th.rg_ok = th.rg < p.RG;
  const long long tile = blockIdx.x;
  const long long b0 = tile * p.BT;
  const bool closed = p.n_nodes == 2;
  const CcNodeDev& M = p.nd[p.n_nodes - 1];  // the model
  const CcNodeDev& C = p.nd[0];              // the controller (closed loop only)
  const bool tape = p.tape != nullptr;

  for (int n = 0; n < p.n_nodes; ++n) {
    cc_stage_mlp(p.nd[n], p.nd[n].f, p.theta[n], smem, th.tid, p.nthreads);
    cc_stage_mlp(p.nd[n], p.nd[n].g, p.theta[n], smem, th.tid, p.nthreads);
    cc_init_node_buffers(p, p.nd[n], smem, th.tid, false);
  }
  __syncthreads();
  if (!closed && p.x0) {
    for (int i = th.tid; i < M.S * p.BT; i += p.nthreads) {
      const int traj = i / M.S, s = i % M.S;
      smem[M.s_X + s * p.BTP + traj] = (b0 + traj < p.B) ? __ldg(p.x0 + (b0 + traj) * M.S + s) : 0.f;
    }
    __syncthreads();
  }
  // y0 = out_scale * g([x0 ; 0 ; (no feed-through for models)])     neural_ode_model.py:160-175
  cc_readout(p, M, smem + M.s_X, smem, nullptr, th);
  if (th.tid < p.BT) {
    for (int o = 0; o < M.O; ++o) {
      const float y0 = smem[M.s_OUT + o * p.BTP + th.tid];
      if (closed) smem[p.s_yprev + o * p.BTP + th.tid] = y0;
      else if (b0 + th.tid < p.B) p.out[(b0 + th.tid) * (long long)(p.T + 1) * M.O + o] = y0;
    }
  }
  __syncthreads();

  float tc = 0.f, tm = 0.f;
  const int Win = closed ? p.R : M.I;
  for (int k0 = 0; k0 < p.T; k0 += CC_TC) {
    const int kend = (p.T - k0) < CC_TC ? (p.T - k0) : CC_TC;
    cc_stage_in(p, p.in, CC_TC * Win, (long long)p.T * Win, (long long)k0 * Win, kend * Win, b0, smem + p.s_in, th.tid);
    if (closed && p.noise)
      cc_stage_in(p, p.noise, CC_TC * M.O, (long long)p.T * M.O, (long long)k0 * M.O, kend * M.O, b0, smem + p.s_noise, th.tid);
    __syncthreads();
    for (int kk = 0; kk < kend; ++kk) {
      const int k = k0 + kk;
      const bool wt = tape && (k % p.ckpt_every == 0);
      const int ck = k / p.ckpt_every;
      if (tape && tile == 0 && th.tid == 0) {  // step times for the reverse pass (fp32 accumulation is not invertible)
        float* ttab = p.tape + (long long)p.n_ckpt * p.tape_rows * p.Bpad;
        if (closed) { ttab[k] = tc; ttab[p.T + k] = tm; } else { ttab[k] = tm; }
      }
      if (closed) {
        if (th.tid < p.BT) {
          float v[16];
          for (int r = 0; r < p.R; ++r) v[r] = smem[p.s_in + (kk * p.R + r) * p.BTP + th.tid];
          for (int o = 0; o < M.O; ++o) {
            const float yp = smem[p.s_yprev + o * p.BTP + th.tid];
            v[p.R + o] = yp;
            if (wt) p.tape[((long long)ck * p.tape_rows + C.S + M.S + o) * p.Bpad + b0 + th.tid] = yp;
          }
          cc_set_inputs(p, C, smem, th.tid, v, tc, false);
        }
        __syncthreads();
        cc_node_step(p, C, smem, th, tc, false, true, wt, tile, ck);
        if (th.tid < p.BT) {
          float a[16];
          for (int i = 0; i < M.I; ++i) {
            a[i] = smem[C.s_OUT + i * p.BTP + th.tid];
            if (p.u_out) smem[p.s_uo + (kk * M.I + i) * p.BTP + th.tid] = a[i];
          }
          cc_set_inputs(p, M, smem, th.tid, a, tm, false);
        }
        __syncthreads();
        cc_node_step(p, M, smem, th, tm, false, true, wt, tile, ck);
        if (th.tid < p.BT) {
          for (int o = 0; o < M.O; ++o) {
            float y = smem[M.s_OUT + o * p.BTP + th.tid];
            if (p.noise) y += smem[p.s_noise + (kk * M.O + o) * p.BTP + th.tid];
            smem[p.s_yprev + o * p.BTP + th.tid] = y;
            smem[p.s_out + (kk * M.O + o) * p.BTP + th.tid] = y;
          }
        }
        if (C.has_t) tc = tc + C.dt;
        if (M.has_t) tm = tm + M.dt;
      } else {
        if (th.tid < p.BT) {
          float v[16];
          for (int i = 0; i < M.I; ++i) v[i] = smem[p.s_in + (kk * M.I + i) * p.BTP + th.tid];
          cc_set_inputs(p, M, smem, th.tid, v, tm, false);
        }
        __syncthreads();
        cc_node_step(p, M, smem, th, tm, false, true, wt, tile, ck);
        if (th.tid < p.BT)
          for (int o = 0; o < M.O; ++o) smem[p.s_out + (kk * M.O + o) * p.BTP + th.tid] = smem[M.s_OUT + o * p.BTP + th.tid];
        if (M.has_t) tm = tm + M.dt;
      }
    }
    __syncthreads();
    if (closed) {
      cc_stage_out(p, p.out, CC_TC * M.O, (long long)p.T * M.O, (long long)k0 * M.O, kend * M.O, b0, smem + p.s_out, th.tid);
      if (p.u_out)
        cc_stage_out(p, p.u_out, CC_TC * M.I, (long long)p.T * M.I, (long long)k0 * M.I, kend * M.I, b0, smem + p.s_uo, th.tid);
    } else {
      // step k produces output index k + 1
      cc_stage_out(p, p.out, CC_TC * M.O, (long long)(p.T + 1) * M.O, (long long)(k0 + 1) * M.O, kend * M.O, b0, smem + p.s_out, th.tid);
    }
    __syncthreads();
  }
  if (!closed && p.x_final) {
    for (int i = th.tid; i < M.S * p.BT; i += p.nthreads) {
      const int traj = i / M.S, s = i % M.S;
      if (b0 + traj < p.B) p.x_final[(b0 + traj) * M.S + s] = smem[M.s_X + s * p.BTP + traj];
    }
  }
}
