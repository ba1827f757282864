// cc_kernels.cuh — device code of the fused neural-ODE rollout (forward) for sm_100a.
//
// Each warp runs the whole time loop for its 32 trajectories (see cc_plan.h): weights are staged once
// per CTA in shared memory, the state tile and the RK4 stage sum live in registers, every layer
// evaluation is a register-tiled FP32 SGEMM out of shared memory, and __syncwarp() is the only
// synchronisation inside the time loop.
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
#define CC_ASSERT_ALIGNED(ptr, n)
#else
#define CC_DEV static inline
#define CC_DEV_NOINLINE static
#define CC_GRID_CONSTANT
#define CC_ASSERT_ALIGNED(ptr, n)                                                          \
  do {                                                                                     \
    if (reinterpret_cast<uintptr_t>(ptr) % (n)) {                                          \
      fprintf(stderr, "ccemu: misaligned %d-byte access at %s:%d\n", (n), __FILE__, __LINE__); \
      abort();                                                                             \
    }                                                                                      \
  } while (0)
#endif
#define CC_FULL 0xffffffffu

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

// Dynamic shared memory is declared at file scope and every pointer into it is rebuilt from this
// symbol inside the function that uses it: a shared pointer that travels through a struct or a
// non-inlined call degrades to a GENERIC pointer (LD.E instead of LDS, ~2x the latency).
#ifndef CC_EMULATE
extern __shared__ __align__(16) float cc_smem_[];
#define CC_SMEM_PTR cc_smem_
#else
#define CC_SMEM_PTR (ccemu::st().smem)
#endif

// per-warp execution context
struct CcWarp {
  int lane, rg, cg;  // lane = cg * 8 + rg
  long long wt;      // warp-tile index
  long long b0;      // first trajectory of the tile
  int nact;          // active trajectories (<= 32)
  int ws_off;        // this warp's shared-memory region (floats from the start of dynamic smem)
  int cs_off;        // CTA-shared region (weights)
};
CC_DEV const CcKParams& cc_params() { return *reinterpret_cast<const CcKParams*>(CC_SMEM_PTR); }
CC_DEV float* cc_ws(const CcWarp& w) { return CC_SMEM_PTR + w.ws_off; }
CC_DEV float* cc_cs(const CcWarp& w) { return CC_SMEM_PTR + w.cs_off; }

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

// stage the weights of one MLP into the CTA-shared region (all threads of the CTA):
//   Wt[K][ldw]: element (k, n) at k*ldw + (n/cpt)*GS + n%cpt, zero padding elsewhere
//   W[N][ldk] : element (n, r) at n*ldk + (r/kcpt)*GSk + r%kcpt for r < Kin            (backward)
//   WV[I][N]  : weights of the input rows                                               (backward)
CC_DEV void cc_stage_mlp(const CcNodeDev& nd, const CcMlpDev& mlp, const float* theta, float* cs, int tid,
                         int nthreads) {
  for (int l = 0; l < mlp.L; ++l) {
    const CcLayerDev& L = mlp.layer[l];
    if (L.s_wt >= 0 && nd.tiny) {  // row-major W[n][CC_TINY_K], zero padded
      float* W = cs + L.s_wt;
      for (int i = tid; i < L.N * CC_TINY_K; i += nthreads) W[i] = cc_weight_at(nd, L, theta, i / CC_TINY_K, i % CC_TINY_K);
    } else if (L.s_wt >= 0) {
      float* Wt = cs + L.s_wt;
      for (int i = tid; i < L.K * L.ldw; i += nthreads) {
        const int k = i / L.ldw, c = i % L.ldw;
        const int g = c / L.GS, j = c % L.GS;
        Wt[i] = (j < L.cpt) ? cc_weight_at(nd, L, theta, g * L.cpt + j, k) : 0.f;
      }
    }
    if (L.s_w >= 0) {
      float* W = cs + L.s_w;
      for (int i = tid; i < L.N * L.ldk; i += nthreads) {
        const int n = i / L.ldk, c = i % L.ldk;
        const int g = c / L.GSk, j = c % L.GSk;
        const int r = g * L.kcpt + j;
        W[i] = (j < L.kcpt && r < L.Kin) ? cc_weight_at(nd, L, theta, n, r) : 0.f;
      }
    }
    if (L.s_wv >= 0) {
      float* WV = cs + L.s_wv;
      for (int i = tid; i < nd.I * L.N; i += nthreads) WV[i] = cc_weight_at(nd, L, theta, i % L.N, nd.row_v + i / L.N);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// register-tiled SGEMM out of shared memory:  acc[i][j] += sum_k A[k][4rg+i] * Bm[k][cg*GS + j]
// The operands of step k+1 are fetched while the FMAs of step k issue (register double buffering).
// ------------------------------------------------------------------------------------------------
template <int TN>
struct CcFrag {
  float4 a;
  float b[TN];
};
template <int TN>
CC_DEV void cc_frag_load(CcFrag<TN>& f, const float* A, const float* Bp) {
  CC_ASSERT_ALIGNED(A, 16);
  CC_ASSERT_ALIGNED(Bp, 16);
  f.a = *reinterpret_cast<const float4*>(A);
#pragma unroll
  for (int q = 0; q < TN / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(Bp + 4 * q);
    f.b[4 * q + 0] = v.x; f.b[4 * q + 1] = v.y; f.b[4 * q + 2] = v.z; f.b[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int j = TN / 4 * 4; j < TN; ++j) f.b[j] = Bp[j];
}
template <int TN>
CC_DEV void cc_frag_fma(const CcFrag<TN>& f, float (&acc)[CC_TM][TN]) {
  const float av[4] = {f.a.x, f.a.y, f.a.z, f.a.w};
#pragma unroll
  for (int i = 0; i < CC_TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], f.b[j], acc[i][j]);
}
// A -> &A[0][4rg] (row stride 32); Bp -> &Bm[0][cg*GS] (row stride ldb)
template <int TN>
CC_DEV void cc_gemm_tile(const float* A, const float* Bp, int ldb, int K, float (&acc)[CC_TM][TN]) {
  CcFrag<TN> cur, nxt;
  cc_frag_load<TN>(cur, A, Bp);
#pragma unroll 2
  for (int k = 1; k < K; ++k) {
    A += CC_LD;
    Bp += ldb;
    cc_frag_load<TN>(nxt, A, Bp);
    cc_frag_fma<TN>(cur, acc);
    cur = nxt;
  }
  cc_frag_fma<TN>(cur, acc);
}

template <int TN>
CC_DEV void cc_zero_tile(float (&t)[CC_TM][TN]) {
#pragma unroll
  for (int i = 0; i < CC_TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) t[i][j] = 0.f;
}

// tile <-> feature-major buffer rows [n][32]; thread owns rows cg*cpt + j (j < cpt), columns 4rg..4rg+3
template <int TN>
CC_DEV void cc_load_tile(const float* buf, int N, int cpt, const CcWarp& w, float (&t)[CC_TM][TN]) {
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = w.cg * cpt + j;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < cpt && n < N) v = *reinterpret_cast<const float4*>(buf + n * CC_LD + 4 * w.rg);
    t[0][j] = v.x; t[1][j] = v.y; t[2][j] = v.z; t[3][j] = v.w;
  }
}
template <int TN>
CC_DEV void cc_store_tile(float* buf, int N, int cpt, const CcWarp& w, const float (&t)[CC_TM][TN]) {
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = w.cg * cpt + j;
    if (j < cpt && n < N) *reinterpret_cast<float4*>(buf + n * CC_LD + 4 * w.rg) = make_float4(t[0][j], t[1][j], t[2][j], t[3][j]);
  }
}

// tile-wide activation / activation-derivative with the kind test hoisted out of the element loops
template <int TN>
CC_DEV void cc_act_tile(int kind, float (&t)[CC_TM][TN]) {
  if (kind == CC_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) t[i][j] = fmaxf(t[i][j], 0.f);
  } else if (kind == CC_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) t[i][j] = tanhf(t[i][j]);
  }
}
// d *= act'(y) with y the activation output
template <int TN>
CC_DEV void cc_act_grad_tile(int kind, float (&d)[CC_TM][TN], const float (&y)[CC_TM][TN]) {
  if (kind == CC_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) d[i][j] = y[i][j] > 0.f ? d[i][j] : 0.f;
  } else if (kind == CC_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) d[i][j] *= 1.f - y[i][j] * y[i][j];
  }
}

// one layer as a tile GEMM; leaves the raw accumulators (bias included via the 1-row) in acc
template <int TN>
CC_DEV void cc_layer_gemm(const CcLayerDev& L, const float* src, const CcWarp& w, float (&acc)[CC_TM][TN]) {
  cc_zero_tile<TN>(acc);
  cc_gemm_tile<TN>(src + 4 * w.rg, cc_cs(w) + L.s_wt + w.cg * L.GS, L.ldw, L.K, acc);
}

// very narrow layers (N <= CC_MAX_SKINNY): the 4 column groups split K, partial sums meet through
// two shuffles (lane bits 3,4 = cg); fixed order -> deterministic.  dst[n][traj] = scale*act(sum).
// Ends with __syncwarp().
CC_DEV void cc_layer_skinny(const CcLayerDev& L, const float* src, float* dst, float scale, const CcWarp& w) {
  float part[CC_MAX_SKINNY][CC_TM];
#pragma unroll
  for (int n = 0; n < CC_MAX_SKINNY; ++n)
#pragma unroll
    for (int i = 0; i < CC_TM; ++i) part[n][i] = 0.f;
  const float* Wt = cc_cs(w) + L.s_wt;
  for (int k = w.cg; k < L.K; k += 4) {
    const float4 a = *reinterpret_cast<const float4*>(src + k * CC_LD + 4 * w.rg);
#pragma unroll
    for (int n = 0; n < CC_MAX_SKINNY; ++n) {
      if (n < L.N) {
        const float wv = Wt[k * L.ldw + (n / L.cpt) * L.GS + n % L.cpt];
        part[n][0] = fmaf(a.x, wv, part[n][0]);
        part[n][1] = fmaf(a.y, wv, part[n][1]);
        part[n][2] = fmaf(a.z, wv, part[n][2]);
        part[n][3] = fmaf(a.w, wv, part[n][3]);
      }
    }
  }
#pragma unroll
  for (int n = 0; n < CC_MAX_SKINNY; ++n) {
    if (n < L.N) {
#pragma unroll
      for (int i = 0; i < CC_TM; ++i) {
        float v = part[n][i];
        v += __shfl_xor_sync(CC_FULL, v, 8);
        v += __shfl_xor_sync(CC_FULL, v, 16);
        part[n][i] = scale * cc_act(L.act, v);
      }
      if (w.cg == 0)
        *reinterpret_cast<float4*>(dst + n * CC_LD + 4 * w.rg) = make_float4(part[n][0], part[n][1], part[n][2], part[n][3]);
    }
  }
  __syncwarp();
}

// hidden layers of an MLP (all but the last): src -> hidden buffers (per warp).  Returns the input
// buffer of the last layer.  hbuf[l] overrides the hidden buffer of layer l (backward keeps them
// per stage).
template <int TN>
CC_DEV const float* cc_mlp_hidden(const CcMlpDev& mlp, const float* src, const int* hbuf, const CcWarp& w) {
  float acc[CC_TM][TN];
#pragma unroll 1
  for (int l = 0; l < mlp.L - 1; ++l) {
    const CcLayerDev& L = mlp.layer[l];
    float* dst = cc_ws(w) + (hbuf ? hbuf[l] : L.w_out);
    if (L.N <= CC_MAX_SKINNY) {
      cc_layer_skinny(L, src, dst, 1.f, w);
    } else {
      cc_layer_gemm<TN>(L, src, w, acc);
      cc_act_tile<TN>(L.act, acc);
      cc_store_tile<TN>(dst, L.N, L.cpt, w, acc);
      __syncwarp();
    }
    src = dst;
  }
  return src;
}

// readout: OUT = out_scale * g(src rows).  Ends with __syncwarp().
template <int TN>
CC_DEV_NOINLINE void cc_readout(int node, int src_off, bool keep_gh, const CcWarp w) {
  const CcNodeDev& nd = cc_params().nd[node];
  const int* ghbuf = keep_gh ? nd.w_GH : nullptr;
  const float* in = cc_mlp_hidden<TN>(nd.g, cc_ws(w) + src_off, ghbuf, w);
  const CcLayerDev& L = nd.g.layer[nd.g.L - 1];
  float* OUT = cc_ws(w) + nd.w_OUT;
  if (L.N <= CC_MAX_SKINNY) {
    cc_layer_skinny(L, in, OUT, nd.out_scale, w);
  } else {
    float acc[CC_TM][TN];
    cc_layer_gemm<TN>(L, in, w, acc);
    cc_act_tile<TN>(L.act, acc);
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] *= nd.out_scale;
    cc_store_tile<TN>(OUT, L.N, L.cpt, w, acc);
    __syncwarp();
  }
}

// k = f(src): hidden layers to smem, last layer into registers (activation applied)
template <int TN>
CC_DEV void cc_rhs(const CcNodeDev& nd, const float* src, const int* hbuf, const CcWarp& w, float (&k)[CC_TM][TN]) {
  const float* in = cc_mlp_hidden<TN>(nd.f, src, hbuf, w);
  const CcLayerDev& L = nd.f.layer[nd.f.L - 1];
  cc_layer_gemm<TN>(L, in, w, k);
  cc_act_tile<TN>(L.act, k);
}

// the node's state tile -> tape record `ck` (128-byte coalesced float4 stores)
template <int TN>
CC_DEV void cc_tape_store(const CcKParams& p, const CcNodeDev& nd, int ck, const CcWarp& w, const float (&x)[CC_TM][TN]) {
  float* rec = p.tape + ((long long)ck * p.tape_rows + nd.tape_row) * p.Bpad + w.b0;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = w.cg * nd.cptS + j;
    if (j < nd.cptS && n < nd.S) *reinterpret_cast<float4*>(rec + (long long)n * p.Bpad + 4 * w.rg) = make_float4(x[0][j], x[1][j], x[2][j], x[3][j]);
  }
}

// One step of a tile-path node.  On entry X holds [x ; t ; v ; 1] and Z the same v / 1 rows.
// On exit X holds x_next and OUT the readout.  keep (backward recompute): stage inputs stay in
// ZS[0..3], hidden activations / k_i are kept per stage, X keeps x.  Ends with __syncwarp().
// A node whose layers and state need at most 4 columns per column group (widths <= 16: the reference's default networks,
// width 10) runs with a [4][4] register tile instead of [4][8]: the shared-memory layout is the same (group stride 8), the
// register-tiled products issue half the FMAs (ncu on the CLI architecture: 3 of 8 tile columns carried data).
CC_DEV bool cc_node_fits4(const CcNodeDev& nd) {
  if (nd.tiny || nd.cptS > 4) return false;
  for (int l = 0; l < nd.f.L; ++l)
    if (nd.f.layer[l].cpt > 4 || nd.f.layer[l].kcpt > 4) return false;
  for (int l = 0; l < nd.g.L; ++l)
    if (nd.g.layer[l].cpt > 4 || nd.g.layer[l].kcpt > 4) return false;
  return true;
}
// call fn<4> when the node fits the narrow tile (only instantiated next to the 8-wide tile), else fn<TN>
#define CC_TILE_DISPATCH(TNV, small4, fn, ...)   \
  do {                                            \
    if constexpr ((TNV) == 8) {                   \
      if (small4) fn<4>(__VA_ARGS__);             \
      else fn<8>(__VA_ARGS__);                    \
    } else {                                      \
      fn<TNV>(__VA_ARGS__);                       \
    }                                             \
  } while (0)

template <int TN>
CC_DEV_NOINLINE void cc_node_step(int node, const CcWarp w, float t, bool keep, bool do_readout, bool write_tape,
                                  int ck) {
  const CcKParams& p = cc_params();
  const CcNodeDev& nd = p.nd[node];
  float* X = cc_ws(w) + nd.w_X;
  float* Z = cc_ws(w) + nd.w_Z;
  const float h = nd.dt;
  if (nd.pre && do_readout) cc_readout<TN>(node, nd.w_X, keep, w);
  float xr[CC_TM][TN], k[CC_TM][TN];
  cc_load_tile<TN>(X, nd.S, nd.cptS, w, xr);
  if (write_tape && nd.tape_row >= 0) cc_tape_store<TN>(p, nd, ck, w, xr);
  if (nd.integrator == CC_INT_RK4) {
    float ks[CC_TM][TN];
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
      float* src = keep ? cc_ws(w) + nd.w_ZS[s] : ((s & 1) ? Z : X);
      float* dst = keep ? (s < 3 ? cc_ws(w) + nd.w_ZS[s + 1] : nullptr) : ((s & 1) ? X : Z);
      if (nd.has_t && !keep && s > 0) {  // stage times t, t+h/2, t+h/2, t+h  (integrate.py:6-13)
        src[nd.row_t * CC_LD + w.lane] = s == 3 ? t + h : t + h / 2.f;
        __syncwarp();
      }
      cc_rhs<TN>(nd, src, keep ? nd.w_HS[s] : nullptr, w, k);
      if (keep && nd.w_KS[s] >= 0) cc_store_tile<TN>(cc_ws(w) + nd.w_KS[s], nd.S, nd.cptS, w, k);
      // ks = k1 + 2 k2 + 2 k3 (+ k4); stage input x + h k / 2 (s = 0, 1), x + h k (s = 2);
      // x_next = x + h (1/6 (ks + k4))  -- the reference's operation order (integrate.py:6-13)
      if (s < 3) {
        const float wk = s == 0 ? 1.f : 2.f;   // exact scalings: identical rounding to the reference
        const float cz = s == 2 ? 1.f : 0.5f;
#pragma unroll
        for (int i = 0; i < CC_TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            const float kk = k[i][j];
            ks[i][j] = s == 0 ? kk : ks[i][j] + wk * kk;
            k[i][j] = xr[i][j] + (h * kk) * cz;
          }
      } else {
#pragma unroll
        for (int i = 0; i < CC_TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) k[i][j] = xr[i][j] + h * ((1.f / 6.f) * (ks[i][j] + k[i][j]));
      }
      if (dst) cc_store_tile<TN>(dst, nd.S, nd.cptS, w, k);
      __syncwarp();
    }
  } else {
    cc_rhs<TN>(nd, X, keep ? nd.w_HS[0] : nullptr, w, k);
    if (keep && nd.w_KS[0] >= 0) cc_store_tile<TN>(cc_ws(w) + nd.w_KS[0], nd.S, nd.cptS, w, k);
#pragma unroll
    for (int i = 0; i < CC_TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) k[i][j] = nd.integrator == CC_INT_EE ? xr[i][j] + h * k[i][j] : k[i][j];
    __syncwarp();  // every lane finished reading X
    if (!keep) cc_store_tile<TN>(X, nd.S, nd.cptS, w, k);
    __syncwarp();
  }
  if (!nd.pre && do_readout) {
    // post-step readout g([x_next ; t ; v]) with the PRE-step time (neural_ode_model.py:129-133)
    if (keep) {
      float* XN = cc_ws(w) + nd.w_XN;  // t / v / 1 rows were set by cc_set_inputs
      cc_store_tile<TN>(XN, nd.S, nd.cptS, w, k);
      __syncwarp();
      cc_readout<TN>(node, nd.w_XN, true, w);
    } else {
      if (nd.has_t) {
        X[nd.row_t * CC_LD + w.lane] = t;
        __syncwarp();
      }
      cc_readout<TN>(node, nd.w_X, false, w);
    }
  }
}

// zero a node's per-warp buffers and set the constant-1 rows (one warp)
CC_DEV void cc_init_node_buffers(const CcNodeDev& nd, const CcWarp& w, bool bwd) {
  auto init = [&](int off, int rows, int one_row) {
    if (off < 0) return;
    float* buf = cc_ws(w) + off;
    for (int r = 0; r < rows; ++r) buf[r * CC_LD + w.lane] = (r == one_row) ? 1.f : 0.f;
  };
  if (!nd.tiny) {
    init(nd.w_X, nd.K0, nd.row_one);
    init(nd.w_Z, nd.K0, nd.row_one);
    init(nd.w_OUT, nd.O, -1);
    for (int which = 0; which < 2; ++which) {
      const CcMlpDev& mlp = which ? nd.g : nd.f;
      for (int l = 0; l < mlp.L - 1; ++l) init(mlp.layer[l].w_out, mlp.layer[l].N + 1, mlp.layer[l].N);
    }
  }
  if (bwd && !nd.tiny) {
    for (int s = 1; s < 4; ++s) init(nd.w_ZS[s], nd.K0, nd.row_one);
    init(nd.w_XN, nd.K0, nd.row_one);
    for (int s = 0; s < 4; ++s)
      for (int l = 0; l < nd.f.L - 1; ++l) init(nd.w_HS[s][l], nd.f.layer[l].N + 1, nd.f.layer[l].N);
    for (int l = 0; l < nd.g.L - 1; ++l) init(nd.w_GH[l], nd.g.layer[l].N + 1, nd.g.layer[l].N);
    init(nd.w_LAM, nd.S, -1);
    init(nd.w_VB, nd.I > 0 ? nd.I : 1, -1);
    init(nd.w_OB, nd.O, -1);
    for (int which = 0; which < 2; ++which) {
      const CcMlpDev& mlp = which ? nd.g : nd.f;
      for (int l = 0; l < mlp.L; ++l) {
        const CcLayerDev& L = mlp.layer[l];
        if (L.w_gw >= 0)
          for (int i = w.lane; i < L.N * L.K; i += 32) cc_ws(w)[L.w_gw + i] = 0.f;
      }
    }
  }
  if (nd.tiny) {
    init(nd.w_X, nd.S, -1);
    init(nd.w_Z, nd.K0, -1);
    init(nd.w_XN, nd.K0, -1);
    init(nd.w_DB, nd.S, -1);
    if (bwd) {
      for (int s = 0; s < 4; ++s) { init(nd.w_ZS[s], nd.K0, -1); init(nd.w_KS[s], nd.S, -1); }
      init(nd.w_LAM, nd.S, -1);
      init(nd.w_OB, 3 * nd.S, -1);
      init(nd.w_DA, nd.S > nd.O ? nd.S : nd.O, -1);
      for (int e = w.lane; e < nd.g_total; e += 32) cc_ws(w)[nd.w_GP + e] = 0.f;
    }
  }
}

// set the t / v rows of a tile-path node's stage buffers for the coming step (lane = trajectory).
// mode 0: forward (X and Z); 1: keep / backward recompute (X, ZS[1..3], XN); 2: X only
CC_DEV void cc_set_inputs(const CcNodeDev& nd, const CcWarp& w, const float (&v)[CC_TINY_K], float t, int mode) {
  const float h = nd.dt;
  float* X = cc_ws(w) + nd.w_X;
#pragma unroll
  for (int i = 0; i < CC_TINY_K; ++i) {
    if (i < nd.I) {
      const float ut = cc_ut(nd.ut, nd.u_scale, v[i]);
      const int o = (nd.row_v + i) * CC_LD + w.lane;
      X[o] = ut;
      if (mode == 1) {
        if (nd.integrator == CC_INT_RK4)
          for (int s = 1; s < 4; ++s) cc_ws(w)[nd.w_ZS[s] + o] = ut;
        if (nd.w_XN >= 0) cc_ws(w)[nd.w_XN + o] = ut;
      } else if (mode == 0 && nd.integrator == CC_INT_RK4) {
        cc_ws(w)[nd.w_Z + o] = ut;
      }
    }
  }
  if (nd.has_t) {
    const int o = nd.row_t * CC_LD + w.lane;
    X[o] = t;
    if (mode == 1) {
      if (nd.integrator == CC_INT_RK4) {
        cc_ws(w)[nd.w_ZS[1] + o] = t + h / 2.f;
        cc_ws(w)[nd.w_ZS[2] + o] = t + h / 2.f;
        cc_ws(w)[nd.w_ZS[3] + o] = t + h;
      }
      if (nd.w_XN >= 0) cc_ws(w)[nd.w_XN + o] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// "tiny" nodes (single-layer f and g, S <= 8, rows <= 12, O <= 4, I <= 4): one trajectory per LANE.
// All per-trajectory vectors live in lane-private columns of per-warp shared memory (buf[row][lane],
// conflict free); the current stage input is gathered into 12 registers and every output is one
// 12-term dot product against a weight row read as three warp-uniform LDS.128.  Compact code (the
// loops over outputs are not unrolled) - the step of the compact feedback controller
// (neural_ode_controller_compact_example.py:106-129) is ~1% of the closed loop's arithmetic.
// Weights of tiny nodes are staged row-major: W[n][CC_TINY_K] (zero padded).
// ------------------------------------------------------------------------------------------------
CC_DEV void cc_tiny_load_in(const float* IN, int K, float (&in)[CC_TINY_K]) {
#pragma unroll
  for (int k = 0; k < CC_TINY_K; ++k) in[k] = (k < K) ? IN[k * CC_LD] : 0.f;
}
CC_DEV void cc_tiny_load_row(const float* Wrow, float (&wr)[CC_TINY_K]) {
#pragma unroll
  for (int q = 0; q < CC_TINY_K / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(Wrow + 4 * q);
    wr[4 * q + 0] = v.x; wr[4 * q + 1] = v.y; wr[4 * q + 2] = v.z; wr[4 * q + 3] = v.w;
  }
}
CC_DEV float cc_tiny_dot(const float* Wrow, const float (&in)[CC_TINY_K]) {
  float wr[CC_TINY_K];
  cc_tiny_load_row(Wrow, wr);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < CC_TINY_K; ++k) acc = fmaf(in[k], wr[k], acc);
  return acc;
}
// t / v~ / 1 rows of a lane-private input column
CC_DEV void cc_tiny_fill_tv(const CcNodeDev& nd, float* IN, float ts, const float (&vt)[CC_TINY_O]) {
  if (nd.has_t) IN[nd.row_t * CC_LD] = ts;
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i)
    if (i < nd.I) IN[(nd.row_v + i) * CC_LD] = vt[i];
  IN[nd.row_one * CC_LD] = 1.f;
}
CC_DEV void cc_tiny_readout(const CcNodeDev& nd, const float* WG, const float* GIN, float (&out)[CC_TINY_O]) {
  float in[CC_TINY_K];
  cc_tiny_load_in(GIN, nd.K0, in);
  const int act = nd.g.layer[0].act;
#pragma unroll
  for (int o = 0; o < CC_TINY_O; ++o) out[o] = (o < nd.O) ? nd.out_scale * cc_act(act, cc_tiny_dot(WG + o * CC_TINY_K, in)) : 0.f;
}

// One step of a tiny node for this lane's trajectory.  The state lives in TX = ws[w_X][lane].
// v: raw input.  out: readout.  KEEP (backward): stage inputs stay in ZS[0..3], stage outputs in
// KS[0..3], the post-step readout input in XN, and TX keeps x.
template <bool KEEP>
CC_DEV void cc_tiny_step(const CcNodeDev& nd, const CcWarp& w, float t, const float (&v)[CC_TINY_O], float (&out)[CC_TINY_O]) {
  float* ws = cc_ws(w) + w.lane;
  const float* cs = cc_cs(w);
  const CcLayerDev& F = nd.f.layer[0];
  const float* WF = cs + F.s_wt;
  const float* WG = cs + nd.g.layer[0].s_wt;
  float* TX = ws + nd.w_X;
  float* TKS = ws + nd.w_DB;
  float* INA = ws + nd.w_Z;
  float* INB = ws + nd.w_XN;
  const float h = nd.dt;
  float vt[CC_TINY_O];
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i) vt[i] = (i < nd.I) ? cc_ut(nd.ut, nd.u_scale, v[i]) : 0.f;
  float* IN0 = KEEP ? ws + nd.w_ZS[0] : INA;
  for (int s = 0; s < nd.S; ++s) IN0[s * CC_LD] = TX[s * CC_LD];
  cc_tiny_fill_tv(nd, IN0, t, vt);
  if (nd.pre) cc_tiny_readout(nd, WG, IN0, out);
  const int nst = nd.integrator == CC_INT_RK4 ? 4 : 1;
  float* XNEXT = KEEP ? INB : TX;  // where x_next goes (keep mode: state rows of XN)
#pragma unroll 1
  for (int st = 0; st < nst; ++st) {
    float* IN = KEEP ? ws + nd.w_ZS[st] : ((st & 1) ? INB : INA);
    float* INn = (st + 1 < nst) ? (KEEP ? ws + nd.w_ZS[st + 1] : ((st & 1) ? INA : INB)) : nullptr;
    float in[CC_TINY_K];
    cc_tiny_load_in(IN, nd.K0, in);
    if (INn) cc_tiny_fill_tv(nd, INn, st == 2 ? t + h : t + h / 2.f, vt);
    // all S outputs as independent 12-term chains (ILP across outputs hides the FMA latency)
    float kk[CC_TINY_S];
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) kk[n] = (n < nd.S) ? cc_tiny_dot(WF + n * CC_TINY_K, in) : 0.f;
#pragma unroll
    for (int n = 0; n < CC_TINY_S; ++n) {
      if (n < nd.S) {
        const float k = cc_act(F.act, kk[n]);
        if (KEEP) (ws + nd.w_KS[st])[n * CC_LD] = k;
        const float x = TX[n * CC_LD];
        if (nd.integrator == CC_INT_RK4) {
          if (st == 0) { TKS[n * CC_LD] = k; INn[n * CC_LD] = x + h * k / 2.f; }
          else if (st == 1) { TKS[n * CC_LD] = TKS[n * CC_LD] + 2.f * k; INn[n * CC_LD] = x + h * k / 2.f; }
          else if (st == 2) { TKS[n * CC_LD] = TKS[n * CC_LD] + 2.f * k; INn[n * CC_LD] = x + h * k; }
          else XNEXT[n * CC_LD] = x + h * ((1.f / 6.f) * (TKS[n * CC_LD] + k));
        } else {
          XNEXT[n * CC_LD] = nd.integrator == CC_INT_EE ? x + h * k : k;
        }
      }
    }
  }
  if (!nd.pre) {
    // post-step readout g([x_next ; t ; v ; 1]) with the pre-step time
    float* GIN = KEEP ? INB : INA;
    if (!KEEP)
      for (int s = 0; s < nd.S; ++s) GIN[s * CC_LD] = TX[s * CC_LD];
    cc_tiny_fill_tv(nd, GIN, t, vt);
    cc_tiny_readout(nd, WG, GIN, out);
  }
}

// The launch plan is copied once into the head of dynamic shared memory; all device code reads it
// from there (node-level functions are not inlined; a by-reference kernel parameter would otherwise
// be spilled to local memory or read through generic loads).
#define CC_PARAM_FLOATS ((int)((sizeof(CcKParams) + 15) / 16 * 4))
#define CC_KERNEL_PROLOGUE(gp)                                                      \
  float* cc_smem = CC_SMEM_PTR;                                                     \
  {                                                                                 \
    const int* src_ = reinterpret_cast<const int*>(&(gp));                          \
    int* dst_ = reinterpret_cast<int*>(cc_smem);                                    \
    for (int i_ = threadIdx.x; i_ < (int)(sizeof(CcKParams) / 4); i_ += blockDim.x) dst_[i_] = src_[i_]; \
  }                                                                                 \
  __syncthreads();                                                                  \
  const CcKParams& p = *reinterpret_cast<const CcKParams*>(cc_smem);                \
  float* cta_smem = cc_smem + CC_PARAM_FLOATS;

CC_DEV void cc_make_warp(const CcKParams& p, float* cta_smem, CcWarp& w) {
  const int tid = threadIdx.x;
  const int wid = tid / 32;
  w.lane = tid % 32;
  w.cg = w.lane / 8;
  w.rg = w.lane % 8;
  w.wt = (long long)blockIdx.x * p.wpc + wid;
  w.b0 = w.wt * 32;
  const long long rem = p.B - w.b0;
  w.nact = rem >= 32 ? 32 : (rem > 0 ? (int)rem : 0);
  (void)cta_smem;
  w.cs_off = CC_PARAM_FLOATS;
  w.ws_off = CC_PARAM_FLOATS + p.cta_floats + wid * p.warp_floats;
}

// ------------------------------------------------------------------------------------------------
// K1 + K3: forward rollout (open loop: n_nodes == 1; closed loop: nd[0] controller, nd[1] model)
// TN: register-tile width of the model; CM: controller mode (none / tiny / tile with TN = 8)
// ------------------------------------------------------------------------------------------------
template <int TN, int CM>
__global__ void __launch_bounds__(256, 1) cc_rollout_fwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  CC_KERNEL_PROLOGUE(gp)
  const bool closed = CM != CC_CTRL_NONE;
  const CcNodeDev& M = p.nd[closed ? 1 : 0];
  const CcNodeDev& C = p.nd[0];
  const bool m4 = cc_node_fits4(M), c4 = CM == CC_CTRL_TILE && cc_node_fits4(C);  // narrow [4][4] register tiles
  for (int n = 0; n < p.n_nodes; ++n) {
    cc_stage_mlp(p.nd[n], p.nd[n].f, p.theta[n], cta_smem, threadIdx.x, blockDim.x);
    cc_stage_mlp(p.nd[n], p.nd[n].g, p.theta[n], cta_smem, threadIdx.x, blockDim.x);
  }
  __syncthreads();  // the only CTA-level barrier: weights are staged
  CcWarp w;
  cc_make_warp(p, cta_smem, w);
  if (w.nact == 0) return;
  const bool act = w.lane < w.nact;
  const long long b = w.b0 + w.lane;  // this lane's trajectory in the per-lane I/O phases
  const bool tape = p.tape != nullptr;
  for (int n = 0; n < p.n_nodes; ++n) cc_init_node_buffers(p.nd[n], w, false);
  __syncwarp();

  const int k_beg = p.seg_k0, k_end = p.seg_k1;  // (the whole horizon unless the host replays one checkpoint segment)
  float tc = closed ? C.t0 : 0.f, tm = M.t0;
  if (p.carry_in) {
    // resume at step k_beg from a tape record: module states, fed-back output, step times
    for (int n = 0; n < p.n_nodes; ++n) {
      const CcNodeDev& nd = p.nd[n];
      for (int s = 0; s < nd.S; ++s) cc_ws(w)[nd.w_X + s * CC_LD + w.lane] = act ? p.carry_in[(long long)(nd.tape_row + s) * p.Bpad + b] : 0.f;
    }
    if (closed)
      for (int o = 0; o < M.O; ++o) cc_ws(w)[p.w_yprev + o * CC_LD + w.lane] = act ? p.carry_in[(long long)(p.tape_yrow + o) * p.Bpad + b] : 0.f;
    if (closed) { tc = p.ttab[k_beg]; tm = p.ttab[p.T + k_beg]; } else { tm = p.ttab[k_beg]; }
    __syncwarp();
  } else {
    if (!closed && p.x0) {
      for (int s = 0; s < M.S; ++s) cc_ws(w)[M.w_X + s * CC_LD + w.lane] = act ? __ldg(p.x0 + b * M.S + s) : 0.f;
      __syncwarp();
    }
    // y0 = out_scale * g([x0 ; t0])     neural_ode_model.py:160-175 (models have no feed-through)
    if (M.has_t) {
      cc_ws(w)[M.w_X + M.row_t * CC_LD + w.lane] = M.t0;
      __syncwarp();
    }
    cc_readout<TN>(closed ? 1 : 0, M.w_X, false, w);
    for (int o = 0; o < M.O; ++o) {
      const float y0 = cc_ws(w)[M.w_OUT + o * CC_LD + w.lane];
      if (closed) cc_ws(w)[p.w_yprev + o * CC_LD + w.lane] = y0;
      else if (act && p.out) p.out[b * (long long)(p.T + 1) * M.O + o] = y0;
    }
    __syncwarp();
  }

  const int Win = closed ? p.R : M.I;
  float vin[CC_TINY_K];  // this step's external input (u or ref), prefetched one step ahead
#pragma unroll
  for (int i = 0; i < CC_TINY_K; ++i) vin[i] = (i < Win && act && k_beg < k_end) ? __ldg(p.in + b * (long long)p.T * Win + (long long)k_beg * Win + i) : 0.f;

  for (int k = k_beg; k < k_end; ++k) {
    const bool wt = tape && ((k - k_beg) % p.ckpt_every == 0);
    const int ck = (k - k_beg) / p.ckpt_every;
    if (tape && w.wt == 0 && w.lane == 0) {  // step times for the reverse pass (fp32 accumulation is not invertible)
      float* ttab = p.ttab;
      if (closed) { ttab[k] = tc; ttab[p.T + k] = tm; } else { ttab[k] = tm; }
    }
    float vcur[CC_TINY_K];
#pragma unroll
    for (int i = 0; i < CC_TINY_K; ++i) vcur[i] = vin[i];
    if (k + 1 < k_end) {
#pragma unroll
      for (int i = 0; i < CC_TINY_K; ++i)
        if (i < Win && act) vin[i] = __ldg(p.in + b * (long long)p.T * Win + (long long)(k + 1) * Win + i);
    }
    if (closed) {
      // controller input = merge_x_y(ref, y_prev) = [ref ; obs]     step_fn.py:187-192
      float vc[CC_TINY_K];
#pragma unroll
      for (int i = 0; i < CC_TINY_K; ++i) vc[i] = vcur[i];
      for (int o = 0; o < M.O; ++o) {
        const float yp = cc_ws(w)[p.w_yprev + o * CC_LD + w.lane];
#pragma unroll
        for (int i = 0; i < CC_TINY_K; ++i)
          if (i == p.R + o) vc[i] = yp;
        if (wt) p.tape[((long long)ck * p.tape_rows + p.tape_yrow + o) * p.Bpad + b] = yp;
      }
      float a[CC_TINY_O];
      if (CM == CC_CTRL_TINY) {
        if (wt)
          for (int s = 0; s < C.S; ++s) p.tape[((long long)ck * p.tape_rows + C.tape_row + s) * p.Bpad + b] = cc_ws(w)[C.w_X + s * CC_LD + w.lane];
        float vcc[CC_TINY_O];
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i) vcc[i] = vc[i];
        cc_tiny_step<false>(C, w, tc, vcc, a);
      } else {
        cc_set_inputs(C, w, vc, tc, 0);
        __syncwarp();
        CC_TILE_DISPATCH(8, c4, cc_node_step, 0, w, tc, false, true, wt, ck);
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i) a[i] = (i < M.I) ? cc_ws(w)[C.w_OUT + i * CC_LD + w.lane] : 0.f;
      }
      if (p.u_out && act)
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i)
          if (i < M.I) p.u_out[b * (long long)p.T * M.I + (long long)k * M.I + i] = a[i];
      float am[CC_TINY_K];
#pragma unroll
      for (int i = 0; i < CC_TINY_K; ++i) am[i] = (i < CC_TINY_O) ? a[i < CC_TINY_O ? i : 0] : 0.f;
      cc_set_inputs(M, w, am, tm, 0);
      __syncwarp();
      CC_TILE_DISPATCH(TN, m4, cc_node_step, closed ? 1 : 0, w, tm, false, true, wt, ck);
      for (int o = 0; o < M.O; ++o) {
        float y = cc_ws(w)[M.w_OUT + o * CC_LD + w.lane];
        if (p.noise && act) y += __ldg(p.noise + b * (long long)p.T * M.O + (long long)k * M.O + o);
        cc_ws(w)[p.w_yprev + o * CC_LD + w.lane] = y;
        if (act && p.out) p.out[b * (long long)p.T * M.O + (long long)k * M.O + o] = y;
      }
      if (C.has_t) tc = tc + C.dt;
    } else {
      cc_set_inputs(M, w, vcur, tm, 0);
      __syncwarp();
      CC_TILE_DISPATCH(TN, m4, cc_node_step, closed ? 1 : 0, w, tm, false, true, wt, ck);
      if (act && p.out)
        for (int o = 0; o < M.O; ++o) p.out[b * (long long)(p.T + 1) * M.O + (long long)(k + 1) * M.O + o] = cc_ws(w)[M.w_OUT + o * CC_LD + w.lane];
    }
    if (M.has_t) tm = tm + M.dt;
  }
  if (!closed && p.x_final && act && k_end == p.T)
    for (int s = 0; s < M.S; ++s) p.x_final[b * M.S + s] = cc_ws(w)[M.w_X + s * CC_LD + w.lane];
}
