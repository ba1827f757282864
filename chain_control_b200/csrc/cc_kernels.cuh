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
#define CC_DEV_NOINLINE __device__ __noinline__
#define CC_GRID_CONSTANT __grid_constant__
#else
#define CC_DEV static inline
#define CC_DEV_NOINLINE static
#define CC_GRID_CONSTANT
#endif
#define CC_TILE float (&)[CC_TM][TN]

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
//   col(j) = 4cg + j (j<4) | 4CG + 4cg + (j-4) (j<8) | 8CG + 2cg + (j-8) (TN == 10)
// so that every LDS.128 / LDS.64 of a warp reads one contiguous, conflict-free span.  The operands
// of step k+1 are fetched while the FMAs of step k issue (register double buffering).
// ------------------------------------------------------------------------------------------------
template <int TN>
struct CcFrag {
  float4 a, b0, b1;
  float b2x, b2y;
};
#ifdef CC_EMULATE
#define CC_ASSERT_ALIGNED(ptr, n)                                                          \
  do {                                                                                     \
    if (reinterpret_cast<uintptr_t>(ptr) % (n)) {                                          \
      fprintf(stderr, "ccemu: misaligned %d-byte access at %s:%d\n", (n), __FILE__, __LINE__); \
      abort();                                                                             \
    }                                                                                      \
  } while (0)
#else
#define CC_ASSERT_ALIGNED(ptr, n)
#endif
template <int TN>
CC_DEV void cc_frag_load(CcFrag<TN>& f, const float* A, const float* B0, const float* B1, const float* B2) {
  CC_ASSERT_ALIGNED(A, 16); CC_ASSERT_ALIGNED(B0, 16); CC_ASSERT_ALIGNED(B1, 16);
  if (TN == 10) CC_ASSERT_ALIGNED(B2, 8);
  f.a = *reinterpret_cast<const float4*>(A);
  f.b0 = *reinterpret_cast<const float4*>(B0);
  f.b1 = *reinterpret_cast<const float4*>(B1);
  if (TN == 10) {
    f.b2x = B2[0];
    f.b2y = B2[1];
  }
}
template <int TN>
CC_DEV void cc_frag_fma(const CcFrag<TN>& f, float (&acc)[CC_TM][TN]) {
  const float av[4] = {f.a.x, f.a.y, f.a.z, f.a.w};
  float bv[10] = {f.b0.x, f.b0.y, f.b0.z, f.b0.w, f.b1.x, f.b1.y, f.b1.z, f.b1.w, 0.f, 0.f};
  if (TN == 10) { bv[8] = f.b2x; bv[9] = f.b2y; }
#pragma unroll
  for (int i = 0; i < CC_TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
}
// A -> &A[0][4rg]; Bm -> &Bm[0][0]; CG = column groups of Bm; cg = this thread's group
template <int TN>
CC_DEV void cc_gemm_tile(const float* A, int lda, const float* Bm, int ldb, int K, int CG, int cg,
                         float (&acc)[CC_TM][TN]) {
  const float* B0 = Bm + 4 * cg;
  const float* B1 = Bm + 4 * CG + 4 * cg;
  const float* B2 = Bm + 8 * CG + 2 * cg;
  CcFrag<TN> cur, nxt;
  cc_frag_load<TN>(cur, A, B0, B1, B2);
#pragma unroll 2
  for (int k = 1; k < K; ++k) {
    cc_frag_load<TN>(nxt, A + k * lda, B0 + k * ldb, B1 + k * ldb, B2 + k * ldb);
    cc_frag_fma<TN>(cur, acc);
    cur = nxt;
  }
  cc_frag_fma<TN>(cur, acc);
}

template <int TN>
CC_DEV int cc_col(int CG, int cg, int j) {
  return j < 4 ? 4 * cg + j : (j < 8 ? 4 * CG + 4 * cg + (j - 4) : 8 * CG + 2 * cg + (j - 8));
}

template <int TN>
CC_DEV void cc_zero_tile(float (&t)[CC_TM][TN]) {
#pragma unroll
  for (int i = 0; i < CC_TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) t[i][j] = 0.f;
}

// tile <-> feature-major buffer rows [n][BTP]
template <int TN>
CC_DEV void cc_load_tile(const float* buf, int BTP, int N, int CG, const CcThread& th, float (&t)[CC_TM][TN]) {
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = cc_col<TN>(CG, th.cg, j);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) { CC_ASSERT_ALIGNED(buf + n * BTP + 4 * th.rg, 16); v = *reinterpret_cast<const float4*>(buf + n * BTP + 4 * th.rg); }
    t[0][j] = v.x; t[1][j] = v.y; t[2][j] = v.z; t[3][j] = v.w;
  }
}
template <int TN>
CC_DEV void cc_store_tile(float* buf, int BTP, int N, int CG, const CcThread& th, const float (&t)[CC_TM][TN]) {
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = cc_col<TN>(CG, th.cg, j);
    if (n < N) *reinterpret_cast<float4*>(buf + n * BTP + 4 * th.rg) = make_float4(t[0][j], t[1][j], t[2][j], t[3][j]);
  }
}

// one layer as a tile GEMM; leaves the raw accumulators (bias included via the 1-row) in acc
template <int TN>
CC_DEV void cc_layer_gemm(const CcKParams& p, const CcLayerDev& L, const float* src, const float* smem,
                          const CcThread& th, float (&acc)[CC_TM][TN]) {
  cc_zero_tile<TN>(acc);
  if (th.rg_ok && th.cg < L.CG)
    cc_gemm_tile<TN>(src + 4 * th.rg, p.BTP, smem + L.s_wt, L.Np, L.K, L.CG, th.cg, acc);
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
template <int TN>
CC_DEV const float* cc_mlp_hidden(const CcKParams& p, const CcMlpDev& mlp, const float* src, float* smem,
                                  const int* hbuf, const CcThread& th) {
  float acc[CC_TM][TN];
  for (int l = 0; l < mlp.L - 1; ++l) {
    const CcLayerDev& L = mlp.layer[l];
    float* dst = smem + (hbuf ? hbuf[l] : L.s_out);
    if (L.N <= CC_MAX_SKINNY) {
      cc_layer_skinny(p, L, src, smem, smem + p.nd[0].s_P, dst, 1.f, th);
    } else {
      cc_layer_gemm<TN>(p, L, src, smem, th, acc);
      if (th.rg_ok && th.cg < L.CG) {
#pragma unroll
        for (int i = 0; i < CC_TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = cc_act(L.act, acc[i][j]);
        cc_store_tile<TN>(dst, p.BTP, L.N, L.CG, th, acc);
      }
      __syncthreads();
    }
    src = dst;
  }
  return src;
}

// readout: OUT = out_scale * g(src rows).  Ends with a barrier.
template <int TN>
CC_DEV_NOINLINE void cc_readout(const CcKParams& p, const CcNodeDev& nd, const float* src, float* smem, const int* ghbuf,
                       const CcThread& th) {
  const float* in = cc_mlp_hidden<TN>(p, nd.g, src, smem, ghbuf, th);
  const CcLayerDev& L = nd.g.layer[nd.g.L - 1];
  float* OUT = smem + nd.s_OUT;
  if (L.N <= CC_MAX_SKINNY) {
    cc_layer_skinny(p, L, in, smem, smem + p.nd[0].s_P, OUT, nd.out_scale, th);
  } else {
    float acc[CC_TM][TN];
    cc_layer_gemm<TN>(p, L, in, smem, th, acc);
    if (th.rg_ok && th.cg < L.CG) {
#pragma unroll
      for (int i = 0; i < CC_TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = nd.out_scale * cc_act(L.act, acc[i][j]);
      cc_store_tile<TN>(OUT, p.BTP, L.N, L.CG, th, acc);
    }
    __syncthreads();
  }
}

// k = f(src): hidden layers to smem, last layer into registers (activation applied)
template <int TN>
CC_DEV void cc_rhs(const CcKParams& p, const CcNodeDev& nd, const float* src, float* smem, const int* hbuf,
                   const CcThread& th, float (&k)[CC_TM][TN]) {
  const float* in = cc_mlp_hidden<TN>(p, nd.f, src, smem, hbuf, th);
  const CcLayerDev& L = nd.f.layer[nd.f.L - 1];
  cc_layer_gemm<TN>(p, L, in, smem, th, k);
  if (L.act != CC_ACT_IDENTITY) {
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) k[i][j] = cc_act(L.act, k[i][j]);
  }
}

// write the node's state tile into the tape record of checkpoint `ck`
template <int TN>
CC_DEV void cc_tape_store(const CcKParams& p, const CcNodeDev& nd, long long tile, int ck, const CcThread& th,
                          const float (&x)[CC_TM][TN]) {
  float* rec = p.tape + ((long long)ck * p.tape_rows + nd.tape_row) * p.Bpad + tile * p.BT;
  if (th.rg_ok && th.cg < nd.CGs) {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = cc_col<TN>(nd.CGs, th.cg, j);
      if (n < nd.S) *reinterpret_cast<float4*>(rec + (long long)n * p.Bpad + 4 * th.rg) = make_float4(x[0][j], x[1][j], x[2][j], x[3][j]);
    }
  }
}

// One step of a node.  On entry X holds [x ; t ; v ; 1], ZA/ZB hold the same t(+h/2)/v/1 rows.
// On exit X holds x_next, OUT holds the readout.  `stage_bufs`: when non-null (backward recompute)
// the stage inputs are kept in s_Z[1..3] instead of the ZA/ZB ping-pong and hidden activations /
// k_i are kept per stage.  Ends with a barrier.
template <int TN>
CC_DEV_NOINLINE void cc_node_step(const CcKParams& p, const CcNodeDev& nd, float* smem, const CcThread& th, float t,
                         bool keep, bool do_readout, bool write_tape, long long tile, int ck) {
  float* X = smem + nd.s_X;
  const float h = nd.dt;
  const bool own = th.rg_ok && th.cg < nd.CGs;
  if (nd.pre && do_readout) cc_readout<TN>(p, nd, X, smem, keep ? nd.s_GH : nullptr, th);
  float xr[CC_TM][TN], k[CC_TM][TN];
  cc_load_tile<TN>(X, p.BTP, own ? nd.S : 0, nd.CGs, th, xr);
  if (write_tape) cc_tape_store<TN>(p, nd, tile, ck, th, xr);
  if (nd.integrator == CC_INT_RK4) {
    float ks[CC_TM][TN];
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
      cc_rhs<TN>(p, nd, src, smem, keep ? nd.s_HS[s] : nullptr, th, k);
      if (own) {
        if (keep && nd.s_KS[s] >= 0) cc_store_tile<TN>(smem + nd.s_KS[s], p.BTP, nd.S, nd.CGs, th, k);
#pragma unroll
        for (int i = 0; i < CC_TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
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
        if (!(keep && s == 3)) cc_store_tile<TN>(dst, p.BTP, nd.S, nd.CGs, th, k);
      }
      __syncthreads();
    }
  } else {
    cc_rhs<TN>(p, nd, X, smem, keep ? nd.s_HS[0] : nullptr, th, k);
    if (own) {
      if (keep && nd.s_KS[0] >= 0) cc_store_tile<TN>(smem + nd.s_KS[0], p.BTP, nd.S, nd.CGs, th, k);
#pragma unroll
      for (int i = 0; i < CC_TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j)
          k[i][j] = nd.integrator == CC_INT_EE ? xr[i][j] + h * k[i][j] : k[i][j];
    }
    __syncthreads();  // everyone finished reading X
    if (own && !keep) cc_store_tile<TN>(X, p.BTP, nd.S, nd.CGs, th, k);
    __syncthreads();
  }
  if (!nd.pre && do_readout) {
    // post-step readout g([x_next ; t ; v]); in `keep` mode X must stay x (the reverse pass needs
    // it), so x_next goes to ZA's state rows, whose t / v / 1 rows equal X's up to t (set by caller)
    if (keep) {
      float* XN = smem + nd.s_ZA;
      if (own) cc_store_tile<TN>(XN, p.BTP, nd.S, nd.CGs, th, k);
      __syncthreads();
      cc_readout<TN>(p, nd, XN, smem, nd.s_GH, th);
    } else {
      cc_readout<TN>(p, nd, X, smem, nullptr, th);
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

// The launch plan is copied once into the head of dynamic shared memory; all device code reads it
// from there (the node-level functions are not inlined, and a by-reference kernel parameter would
// otherwise be spilled to local memory or read through generic loads).
#define CC_PARAM_FLOATS ((int)((sizeof(CcKParams) + 15) / 16 * 4))
#ifndef CC_EMULATE
#define CC_SMEM_BASE extern __shared__ __align__(16) float cc_smem[];
#else
#define CC_SMEM_BASE float* cc_smem = ccemu::st().smem;
#endif
#define CC_KERNEL_PROLOGUE(gp)                                                      \
  CC_SMEM_BASE                                                                      \
  {                                                                                 \
    const int* src_ = reinterpret_cast<const int*>(&(gp));                          \
    int* dst_ = reinterpret_cast<int*>(cc_smem);                                    \
    for (int i_ = threadIdx.x; i_ < (int)(sizeof(CcKParams) / 4); i_ += blockDim.x) dst_[i_] = src_[i_]; \
  }                                                                                 \
  __syncthreads();                                                                  \
  const CcKParams& p = *reinterpret_cast<const CcKParams*>(cc_smem);                \
  float* smem = cc_smem + CC_PARAM_FLOATS;

// ------------------------------------------------------------------------------------------------
// K1 + K3: forward rollout (open loop: n_nodes == 1; closed loop: nd[0] controller, nd[1] model)
// ------------------------------------------------------------------------------------------------
template <int TN>
__global__ void __launch_bounds__(320, 1) cc_rollout_fwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  CC_KERNEL_PROLOGUE(gp)
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
  cc_readout<TN>(p, M, smem + M.s_X, smem, nullptr, th);
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
        cc_node_step<TN>(p, C, smem, th, tc, false, true, wt, tile, ck);
        if (th.tid < p.BT) {
          float a[16];
          for (int i = 0; i < M.I; ++i) {
            a[i] = smem[C.s_OUT + i * p.BTP + th.tid];
            if (p.u_out) smem[p.s_uo + (kk * M.I + i) * p.BTP + th.tid] = a[i];
          }
          cc_set_inputs(p, M, smem, th.tid, a, tm, false);
        }
        __syncthreads();
        cc_node_step<TN>(p, M, smem, th, tm, false, true, wt, tile, ck);
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
        cc_node_step<TN>(p, M, smem, th, tm, false, true, wt, tile, ck);
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
