// cc_tc_kernels.cuh -- tcgen05 (5th-gen tensor core) rollout kernels for sm_100a.
//
// Used when the MODEL node's right-hand side f is a single Linear layer (depth 0: the compact model of the
// reference's notebooks 3/4 and end_to_end_learning, neural_ode_model_compact_example.py:85-124, and the
// full model with f_depth=0, neural_ode_model.py:106-149).  Restates the same reference functions as
// cc_kernels.cuh (abstract.py:54-70,97-127; integrate.py:1-28); only the execution model differs:
//
//   * one CTA = one tile of 128 trajectories = 4 warps (thread = trajectory = TMEM lane; lane 0 of warp 0 also
//     issues the MMAs); two CTAs per SM (255 registers per thread) interleave so that one tile's epilogue
//     overlaps the other tile's MMAs.
//   * every stage evaluation k = f([z ; v~ ; 1]) of all 128 trajectories is ONE batch of tcgen05.mma
//     (M = 128, N = NP, K = 8 per instruction, kind::tf32, A from TENSOR MEMORY, B = weights in shared memory,
//     accumulator in TMEM).  fp32 accuracy comes from the 3xTF32 split  z.W = z_lo.W_hi + z_hi.W_lo + z_hi.W_hi
//     (hi = upper 19 bits, lo = exact remainder; small terms accumulated first).
//   * a worker thread reads its row of D with tcgen05.ld, does the RK4 bookkeeping for its trajectory in
//     registers (state x, stage sum), splits the next stage input and writes it back with tcgen05.st;
//     mbarriers (a_ready: 128 arrivals, d_ready: tcgen05.commit) are the only synchronisation in the loop.
//   * readout g(x) (O <= 4 outputs), the tiny feedback controller (cc_tc_ctrl.cuh, register resident) and all
//     I/O are FFMA work of the same thread, placed behind the publish of a stage so that it overlaps the MMAs.
//
// TMEM map of a CTA (256 columns allocated): D [0, NP), A_hi [NP, NP+KP), A_lo [NP+KP, NP+2KP).
// A column layout (forward): [0, 8*NCH) state (zero beyond S); v~_i and the constant 1 sit at the fixed
// in-chunk positions 3+i / 7 of the LAST chunk of the state range if S <= 8*NCH-5, else of one extra chunk.
#pragma once
#include "cc_kernels.cuh"
#include "cc_tc_ctrl.cuh"
#include "cc_tmem.cuh"

#define CC_TC_THREADS 128
#define CC_TC_TMEM_COLS 256

// column of v~_i (i < 4) / of the constant 1 inside the A operand
CC_DEV int cc_tc_tail_base(int S, int SR) { return (S + 5 <= SR) ? SR - 5 : SR + 3; }

template <int NC>
CC_DEV void cc_tc_ld(uint32_t addr, uint32_t (&r)[NC]) {
  if constexpr (NC == 32) cc_tmem_ld32(addr, r);
  else if constexpr (NC == 16) cc_tmem_ld16(addr, r);
  else if constexpr (NC == 8) cc_tmem_ld8(addr, r);
  else cc_tmem_ld4(addr, r);
}

// split NC values (a multiple of 8) into tf32 hi / exact lo and store them at columns c0.. of A_hi / A_lo
template <int NC>
CC_DEV void cc_tc_store_split(uint32_t aHi, uint32_t aLo, int c0, const float (&z)[NC]) {
#pragma unroll
  for (int j = 0; j < NC / 8; ++j) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {  // packed fp32x2 subtract (FADD2): lo = z - hi, exact
      const float v0 = z[8 * j + i], v1 = z[8 * j + i + 1];
      hi[i] = __float_as_uint(v0) & 0xFFFFE000u;
      hi[i + 1] = __float_as_uint(v1) & 0xFFFFE000u;
      const float2 l = __fadd2_rn(make_float2(v0, v1), make_float2(-__uint_as_float(hi[i]), -__uint_as_float(hi[i + 1])));
      lo[i] = __float_as_uint(l.x);
      lo[i + 1] = __float_as_uint(l.y);
    }
    cc_tmem_st8(aHi + c0 + 8 * j, hi);
    cc_tmem_st8(aLo + c0 + 8 * j, lo);
  }
}

// forward epilogue of one RK4 stage for columns [C0, C0+NC) of this trajectory, in the reference's operation
// order (integrate.py:6-13: x + h*k1/2, x + h*k2/2, x + h*k3, x + dt*(1/6*(k1+2k2+2k3+k4))).
//   generic (stages 1..3 of 4): acc += wk * k (acc starts at 0) ; next stage input z = x + (h*k)*cz -> A
//   final   (last stage, or the single Euler stage): x <- x + h*(1/6*(acc + k))  |  x + h*k ; acc <- 0
// ONE generic body serves three stages with run-time (wk, cz): the hot loop has to fit the instruction cache.
template <int SR, int C0, int NC, bool FINAL, bool FACT>
CC_DEV void cc_tc_fwd_group(uint32_t aD, uint32_t aHi, uint32_t aLo, float (&x)[SR], float (&acc)[SR], int fact, float h, float wk, float cz, bool rk4) {
  uint32_t r[NC];
  cc_tc_ld<NC>(aD + C0, r);
  cc_tmem_wait_ld();
  float z[NC];
  const float2 h2 = make_float2(h, h), wk2 = make_float2(wk, wk), cz2 = make_float2(cz, cz);
#pragma unroll
  for (int i = 0; i < NC; i += 2) {  // two columns per instruction: packed fp32x2 FMUL2 / FFMA2 (same rounding as the scalar forms)
    float2 k = make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
    if (FACT) { k.x = cc_act(fact, k.x); k.y = cc_act(fact, k.y); }
    if (!FINAL) {
      const float2 a = __ffma2_rn(wk2, k, make_float2(acc[C0 + i], acc[C0 + i + 1]));
      acc[C0 + i] = a.x; acc[C0 + i + 1] = a.y;
      const float2 zz = __ffma2_rn(__fmul2_rn(h2, k), cz2, make_float2(x[C0 + i], x[C0 + i + 1]));
      z[i] = zz.x; z[i + 1] = zz.y;
    } else {
      float2 s = k;
      if (rk4) s = __fmul2_rn(make_float2(1.f / 6.f, 1.f / 6.f), __fadd2_rn(make_float2(acc[C0 + i], acc[C0 + i + 1]), k));
      const float2 xn = __ffma2_rn(h2, s, make_float2(x[C0 + i], x[C0 + i + 1]));
      x[C0 + i] = xn.x; x[C0 + i + 1] = xn.y;
      acc[C0 + i] = 0.f; acc[C0 + i + 1] = 0.f;
    }
  }
  if (!FINAL) cc_tc_store_split<NC>(aHi, aLo, C0, z);
}
template <int SR, bool FINAL, bool FACT>
CC_DEV void cc_tc_fwd_stage(uint32_t aD, uint32_t aHi, uint32_t aLo, float (&x)[SR], float (&acc)[SR], int fact, float h, float wk, float cz, bool rk4) {
  cc_tc_fwd_group<SR, 0, 32, FINAL, FACT>(aD, aHi, aLo, x, acc, fact, h, wk, cz, rk4);
  if constexpr (SR == 56) {
    cc_tc_fwd_group<SR, 32, 16, FINAL, FACT>(aD, aHi, aLo, x, acc, fact, h, wk, cz, rk4);
    cc_tc_fwd_group<SR, 48, 8, FINAL, FACT>(aD, aHi, aLo, x, acc, fact, h, wk, cz, rk4);
  }
}

// write the whole stage input (state range + optional tail chunk) for the first stage of a step
template <int SR>
CC_DEV void cc_tc_write_state(uint32_t aHi, uint32_t aLo, const float (&x)[SR], bool extra_chunk, const float (&vt)[CC_TINY_O]) {
#pragma unroll
  for (int c = 0; c < SR / 8; ++c) {
    float z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = x[8 * c + i];
    cc_tc_store_split<8>(aHi, aLo, 8 * c, z);
  }
  if (extra_chunk) {
    const float z[8] = {0.f, 0.f, 0.f, vt[0], vt[1], vt[2], vt[3], 1.f};
    cc_tc_store_split<8>(aHi, aLo, SR, z);
  }
}

// y = out_scale * act(g . x + b): per-thread dot products against broadcast weight rows gw[o][SR+4]
// (zero beyond S, bias at index SR)
template <int SR>
CC_DEV void cc_tc_readout(const float* gw, const float (&x)[SR], int O, int gact, float out_scale, float (&y)[CC_TINY_O]) {
#pragma unroll
  for (int o = 0; o < CC_TINY_O; ++o) {
    y[o] = 0.f;
    if (o < O) {
      const float* row = gw + o * (SR + 4);
      float s0 = row[SR], s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < SR; i += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(row + i);
        s0 = fmaf(wv.x, x[i], s0); s1 = fmaf(wv.y, x[i + 1], s1); s2 = fmaf(wv.z, x[i + 2], s2); s3 = fmaf(wv.w, x[i + 3], s3);
      }
      y[o] = out_scale * cc_act(gact, (s0 + s1) + (s2 + s3));
    }
  }
}

struct CcTcShared {
  float* bhi; float* blo; float* gw; float* wfc; float* wgc; uint64_t* a_ready; uint64_t* d_ready; uint32_t* tmem_base; float* gp;
};
template <int SR>
CC_DEV CcTcShared cc_tc_carve(float* smem, const CcKParams& p) {
  CcTcShared s;
  s.bhi = smem + p.tc_off;
  s.blo = s.bhi + p.tc_KP * p.tc_NP;
  s.gw = s.blo + p.tc_KP * p.tc_NP;
  s.wfc = s.gw + CC_TINY_O * (SR + 4);          // controller f weights [8][16]
  s.wgc = s.wfc + CC_TINY_S * CC_CK;            // controller g weights [4][16]
  s.a_ready = reinterpret_cast<uint64_t*>(s.wgc + CC_TINY_O * CC_CK);
  s.d_ready = s.a_ready + 1;
  s.tmem_base = reinterpret_cast<uint32_t*>(s.d_ready + 1);
  s.gp = s.wgc + CC_TINY_O * CC_CK + 8;         // (reverse pass) lane-private gradient accumulators
  return s;
}
// floats of the region; the reverse pass appends (S_c + O_c) * CC_CKU * 128 accumulator floats
#define CC_TC_REGION_FLOATS(SR, KP, NP) (2 * (KP) * (NP) + CC_TINY_O * ((SR) + 4) + (CC_TINY_S + CC_TINY_O) * CC_CK + 8)

// stage the MMA B operand (hi/lo split) and the readout weights.  B is K-major without swizzle:
// element (n, k) at ((k/4)*NP + n)*4 + k%4  ->  8x16-byte core matrices, SBO = 128 B, LBO = NP*16 B.
//   forward : B[n][k] = weight of f output n w.r.t. A column k ([x ; v ; 1] through the tail positions)
//   backward: B[n'][k] = weight of f output k w.r.t. input n' (n' < S state, n' = SR + i input i): the transpose
// gfold (reverse pass of the trainable model, cc_tcw_kernels.cuh): contraction index k = S + o carries g's weight of output o, so that a
// readout cotangent d_o placed in column S + o of kbar adds g^T d to the state adjoint inside the same MMA batch
template <int SR>
CC_DEV void cc_tc_stage_b(const CcNodeDev& M, const float* theta, int KP, int NP, float* bhi, float* blo, bool bwd, int tid, int nthreads, bool gfold = false) {
  const CcLayerDev& F = M.f.layer[0];
  const int tb = cc_tc_tail_base(M.S, SR);
  for (int i = tid; i < KP * NP; i += nthreads) {
    const int n = i / KP, k = i % KP;
    float w = 0.f;
    if (!bwd) {
      int lr = -1;
      if (k < M.S) lr = k;
      else if (k >= tb && k < tb + M.I) lr = M.row_v + (k - tb);
      else if (k == tb + 4) lr = M.row_one;
      if (lr >= 0 && n < M.S) w = cc_weight_at(M, F, theta, n, lr);
    } else {
      int lr = -1;
      if (n < M.S) lr = n;
      else if (n >= SR && n < SR + M.I) lr = M.row_v + (n - SR);
      if (lr >= 0 && k < M.S) w = cc_weight_at(M, F, theta, k, lr);
      if (gfold && n < M.S && k >= M.S && k < M.S + M.O) w = cc_weight_at(M, M.g.layer[0], theta, k - M.S, n);
    }
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    const int off = ((k >> 2) * NP + n) * 4 + (k & 3);
    bhi[off] = hi;
    blo[off] = w - hi;
  }
}
// readout weights gw[o][SR + 4]: zero beyond S, bias at index SR
template <int SR>
CC_DEV void cc_tc_stage_gw(const CcNodeDev& M, const float* theta, float* gw, int tid, int nthreads) {
  const CcLayerDev& G = M.g.layer[0];
  for (int i = tid; i < CC_TINY_O * (SR + 4); i += nthreads) {
    const int o = i / (SR + 4), s = i % (SR + 4);
    float w = 0.f;
    if (o < M.O) {
      if (s < M.S) w = cc_weight_at(M, G, theta, o, s);
      else if (s == SR) w = cc_weight_at(M, G, theta, o, M.row_one);
    }
    gw[i] = w;
  }
}
template <int SR>
CC_DEV void cc_tc_stage_weights(const CcKParams& p, const CcNodeDev& M, const float* theta, const CcTcShared& sh, bool bwd, int tid, int nthreads) {
  cc_tc_stage_b<SR>(M, theta, p.tc_KP, p.tc_NP, sh.bhi, sh.blo, bwd, tid, nthreads);
  cc_tc_stage_gw<SR>(M, theta, sh.gw, tid, nthreads);
}

// one batch of 3 * NA MMAs: D = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi   (issued by ONE thread; fully unrolled with
// compile-time shapes so that an MMA costs a handful of instructions -- the issue rate of this thread bounds the batch)
template <int NA, int NP>
CC_DEV void cc_tc_issue_batch(uint32_t tbase, uint32_t bhi, uint32_t blo) {
  constexpr int KP = 8 * NA;
  constexpr uint32_t kstep = 2u * NP * 16u, lbo = NP * 16u;
  constexpr uint32_t idesc = cc_umma_idesc_tf32(128, NP);
  const uint32_t aHi = tbase + NP, aLo = tbase + NP + KP;
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(tbase, aLo + 8 * js, cc_umma_desc_kmajor(bhi + js * kstep, lbo, 128), idesc, js ? 1u : 0u);
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(tbase, aHi + 8 * js, cc_umma_desc_kmajor(blo + js * kstep, lbo, 128), idesc, 1u);
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(tbase, aHi + 8 * js, cc_umma_desc_kmajor(bhi + js * kstep, lbo, 128), idesc, 1u);
}

// prologue shared by both kernels: parameter copy, weight staging, barriers, TMEM allocation
#define CC_TC_PROLOGUE(gp, BWD)                                                                       \
  CC_DIRECT_PARAMS(gp)                                                                                \
  constexpr int SR = 8 * NCH;                                                                         \
  const bool closed = CM != CC_CTRL_NONE;                                                             \
  const CcNodeDev& M = p.nd[closed ? 1 : 0];                                                          \
  const CcNodeDev& C = p.nd[0];                                                                       \
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;                                      \
  const CcTcShared sh = cc_tc_carve<SR>(cc_smem, p);                                                  \
  (void)cta_smem;                                                                                     \
  if (closed) cc_ctrl_stage_weights(C, p.theta[0], sh.wfc, sh.wgc, tid, blockDim.x);                  \
  cc_tc_stage_weights<SR>(p, M, p.theta[closed ? 1 : 0], sh, BWD, tid, blockDim.x);                   \
  if (warp == 0) {                                                                                    \
    if (lane == 0) {                                                                                  \
      cc_mbar_init(sh.a_ready, 128);                                                                  \
      cc_mbar_init(sh.d_ready, 1);                                                                    \
      cc_mbar_fence_init();                                                                           \
    }                                                                                                 \
    __syncwarp();                                                                                     \
    cc_tmem_alloc(sh.tmem_base, CC_TC_TMEM_COLS);                                                     \
  }                                                                                                   \
  cc_fence_proxy_async();                                                                             \
  cc_tc_fence_before();                                                                               \
  __syncthreads();                                                                                    \
  cc_tc_fence_after();                                                                                \
  const uint32_t tbase = *sh.tmem_base;                                                               \
  const int KP = p.tc_KP, NP = p.tc_NP;                                                               \
  CcTcIssue q;                                                                                        \
  q.tbase = tbase; q.bhi = cc_smem_u32(sh.bhi); q.blo = cc_smem_u32(sh.blo);                          \
  q.idesc = cc_umma_idesc_tf32(128, NP); q.batch = 0; q.KP = KP; q.NP = NP; q.warp = warp; q.lane = lane;

#define CC_TC_EPILOGUE()                                          \
  cc_tc_fence_before();                                           \
  __syncthreads();                                                \
  if (warp == 0) cc_tmem_dealloc(tbase, CC_TC_TMEM_COLS);

CC_DEV void cc_tc_make_warp(const CcKParams& p, int warp, int lane, CcWarp& w) {
  w.lane = lane; w.cg = lane / 8; w.rg = lane % 8;
  w.wt = (long long)blockIdx.x * 4 + warp;
  w.b0 = w.wt * 32;
  const long long rem = p.B - w.b0;
  w.nact = rem >= 32 ? 32 : (rem > 0 ? (int)rem : 0);
  w.cs_off = CC_PARAM_FLOATS;
  w.ws_off = CC_PARAM_FLOATS + p.cta_floats + warp * p.warp_floats;
}
CC_DEV void cc_tc_stagger(const CcKParams& p) {
  if (p.tc_stagger > 0 && ((blockIdx.x / p.tc_sms) & 1)) {
    const long long t0 = clock64();
    while (clock64() - t0 < p.tc_stagger) {}
  }
}
// per-thread MMA issue context (only lane 0 of warp 0 issues; the whole warp 0 waits for the 128 arrivals)
struct CcTcIssue {
  uint32_t tbase, bhi, blo, idesc;
  uint32_t batch;  // batches published so far: batch & 3 = the warp that issues the next one, batch & 1 = a_ready parity
  int KP, NP, warp, lane;
};
// publish the stage input just written to TMEM: all 128 threads arrive; the issuing warp waits for the last arrival and
// its lane 0 issues the MMA batch + commit (every thread then sleeps on d_ready until the batch has completed)
template <int NCH, bool BWD>
CC_DEV void cc_tc_publish(const CcTcShared& sh, CcTcIssue& q) {
  cc_tmem_wait_st();
  cc_tc_fence_before();
  cc_mbar_arrive(sh.a_ready);
  const uint32_t batch = q.batch++;
  if (q.warp == (int)(batch & 3u)) {  // the issuing role rotates: the issuer is busy for the whole batch (the tensor pipe
                                      // back-pressures tcgen05.mma), the other three warps use that time for side work
    cc_mbar_wait(sh.a_ready, batch & 1u);
    cc_tc_fence_after();
    if (q.lane == 0) {
      // shapes fixed by the host plan (cc_abi.cu): forward KP = 8*NCH (+8 with the extra tail chunk), NP = ceil16(8*NCH);
      // backward KP = 8*NCH, NP = ceil16(8*NCH + 4)
      constexpr int NP = BWD ? (8 * NCH + CC_TINY_O + 15) / 16 * 16 : (8 * NCH + 15) / 16 * 16;
      if (q.KP == 8 * NCH) cc_tc_issue_batch<NCH, NP>(q.tbase, q.bhi, q.blo);
      else cc_tc_issue_batch<NCH + 1, NP>(q.tbase, q.bhi, q.blo);
      cc_tc_commit(sh.d_ready);
    }
    __syncwarp();
  }
}
CC_DEV void cc_tc_await(const CcTcShared& sh, uint32_t& par) {
  cc_mbar_wait(sh.d_ready, par);
  par ^= 1;
  cc_tc_fence_after();
}

// ------------------------------------------------------------------------------------------------
// K1 + K3 on tcgen05: forward rollout.  CM: CC_CTRL_NONE (open loop) or CC_CTRL_TINY (closed loop).
// ------------------------------------------------------------------------------------------------
template <int NCH, int CM, bool FACT, int CFG>
__global__ void __launch_bounds__(CC_TC_THREADS, 2) cc_tc_fwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  CC_TC_PROLOGUE(gp, false)
  const CcCtrlDims CD = cc_ctrl_dims<CFG>(C);
  const int nst = M.integrator == CC_INT_RK4 ? 4 : 1;
  {
    CcWarp w;
    cc_tc_make_warp(p, warp, lane, w);
    const bool act = lane < w.nact;
    const long long b = w.b0 + lane;
    const bool tape = p.tape != nullptr;
    const uint32_t aD = tbase + ((uint32_t)(warp * 32) << 16), aHi = aD + NP, aLo = aD + NP + KP;
    const CcLayerDev& F = M.f.layer[0];
    const CcLayerDev& G = M.g.layer[0];
    const float h = M.dt;
    const int tb = cc_tc_tail_base(M.S, SR);
    const bool extra = tb >= SR;
    float xc[CC_TINY_S];  // controller state (registers)
#pragma unroll
    for (int i = 0; i < CC_TINY_S; ++i) xc[i] = 0.f;
    float x[SR], acc[SR];
#pragma unroll
    for (int i = 0; i < SR; ++i) {
      x[i] = (!closed && p.x0 && act && i < M.S) ? __ldg(p.x0 + b * M.S + i) : 0.f;
      acc[i] = 0.f;
    }
    if (!extra) x[SR - 1] = 1.f;
    // y0 = g(x0)      neural_ode_model.py:160-175
    float yprev[CC_TINY_O];
    cc_tc_readout<SR>(sh.gw, x, M.O, G.act, M.out_scale, yprev);
    if (!closed && act)
      for (int o = 0; o < M.O; ++o) p.out[b * (long long)(p.T + 1) * M.O + o] = yprev[o];
    float tc = closed ? C.t0 : 0.f;
    const int Win = closed ? p.R : M.I;
    float vin[CC_TINY_O];
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) vin[i] = (i < Win && act && p.T > 0) ? __ldg(p.in + b * (long long)p.T * Win + i) : 0.f;
    uint32_t par = 0;
    cc_tc_stagger(p);
    for (int k = 0; k < p.T; ++k) {
      const bool wt = tape && (k % p.ckpt_every == 0);
      const int ck = k / p.ckpt_every;
      if (tape && w.wt == 0 && lane == 0) {
        float* ttab = p.tape + (long long)p.n_ckpt * p.tape_rows * p.Bpad;
        if (closed) { ttab[k] = tc; ttab[p.T + k] = 0.f; } else { ttab[k] = 0.f; }
      }
      float vcur[CC_TINY_O];
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) vcur[i] = vin[i];
      if (k + 1 < p.T) {
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i)
          if (i < Win && act) vin[i] = __ldg(p.in + b * (long long)p.T * Win + (long long)(k + 1) * Win + i);
      }
      float a[CC_TINY_O];  // raw model input of this step
      float vc[CC_TINY_O];
      if (closed) {
        // controller input = merge_x_y(ref, y_prev) = [ref ; obs]     step_fn.py:187-192
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i) {
          vc[i] = vcur[i];
#pragma unroll
          for (int o = 0; o < CC_TINY_O; ++o)
            if (o < M.O && i == p.R + o) vc[i] = yprev[o];
        }
        if (wt && act) {
          for (int o = 0; o < M.O; ++o) p.tape[((long long)ck * p.tape_rows + p.tape_yrow + o) * p.Bpad + b] = yprev[o];
#pragma unroll
          for (int s = 0; s < CC_TINY_S; ++s)
            if (s < C.S) p.tape[((long long)ck * p.tape_rows + C.tape_row + s) * p.Bpad + b] = xc[s];
        }
        if (wt && act && M.tape_row >= 0)  // non-linear frozen model: its state is part of the carry the FFMA reverse pass reads
#pragma unroll
          for (int s = 0; s < SR; ++s)
            if (s < M.S) p.tape[((long long)ck * p.tape_rows + M.tape_row + s) * p.Bpad + b] = x[s];
        // compact controller: u = out_scale * g_c(x_c) needs no state update -> the update overlaps the first MMA batch
        if (C.pre) cc_ctrl_readout_pre(CD, sh.wgc, xc, tc, vc, a);
        else cc_ctrl_update<false>(CD, sh.wfc, sh.wgc, xc, tc, vc, a, nullptr);
        if (p.u_out && act)
#pragma unroll
          for (int i = 0; i < CC_TINY_O; ++i)
            if (i < M.I) p.u_out[b * (long long)p.T * M.I + (long long)k * M.I + i] = a[i];
      } else {
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i) a[i] = vcur[i];
        if (wt && act)  // open loop: the model state of every checkpointed step (pre-step) is the tape
#pragma unroll
          for (int s = 0; s < SR; ++s)
            if (s < M.S) p.tape[((long long)ck * p.tape_rows + s) * p.Bpad + b] = x[s];
      }
      float vt[CC_TINY_O];
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) vt[i] = (i < M.I) ? cc_ut(M.ut, M.u_scale, a[i]) : 0.f;
      if (!extra) {
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i) x[SR - 5 + i] = vt[i];
      }
      cc_tc_write_state<SR>(aHi, aLo, x, extra, vt);
      cc_tc_publish<NCH, false>(sh, q);
      // ---- work that overlaps the first MMA batch ----
      float y[CC_TINY_O];
      if (M.pre) cc_tc_readout<SR>(sh.gw, x, M.O, G.act, M.out_scale, y);  // compact: y = g(x) BEFORE the update
      // compact controller: its state update runs one stage per MMA window (the readout above needed only x_c)
      CcCtrlRun crun;
      const int nstc = closed ? cc_ctrl_nstages(CD) : 0;
      const bool cpipe = closed && C.pre;
      int cst = 0;
      // window j (after the j-th publish of the step): a warp that is not issuing runs up to two controller stages,
      // staying at most one stage ahead (cst < j + 2) -> every warp is done before the step's last window
      int win = 0;
      if (cpipe) {
        cc_ctrl_begin(CD, xc, tc, vc, crun);
        if (warp != (int)((q.batch - 1u) & 3u))
          for (int c2 = 0; c2 < 2 && cst < nstc && cst < win + 2; ++c2) cc_ctrl_stage<false>(CD, sh.wfc, cst++, xc, tc, crun, nullptr);
      }
#pragma unroll 1
      for (int st = 0; st < nst - 1; ++st) {
        cc_tc_await(sh, par);
        cc_tc_fwd_stage<SR, false, FACT>(aD, aHi, aLo, x, acc, F.act, h, st == 0 ? 1.f : 2.f, st == 2 ? 1.f : 0.5f, true);
        cc_tc_publish<NCH, false>(sh, q);
        ++win;
        if (cpipe && warp != (int)((q.batch - 1u) & 3u))
          for (int c2 = 0; c2 < 2 && cst < nstc && cst < win + 2; ++c2) cc_ctrl_stage<false>(CD, sh.wfc, cst++, xc, tc, crun, nullptr);
      }
      if (cpipe) {
#pragma unroll 1
        for (; cst < nstc; ++cst) cc_ctrl_stage<false>(CD, sh.wfc, cst, xc, tc, crun, nullptr);
        float dummy[CC_TINY_O];
        cc_ctrl_finish<false>(CD, sh.wgc, xc, tc, crun, dummy, nullptr);
      }
      if (closed && C.has_t) tc = tc + C.dt;
      cc_tc_await(sh, par);
      cc_tc_fwd_stage<SR, true, FACT>(aD, aHi, aLo, x, acc, F.act, h, 0.f, 0.f, nst == 4);
      if (!M.pre) cc_tc_readout<SR>(sh.gw, x, M.O, G.act, M.out_scale, y);  // full model: y = g(x_next)
      if (closed) {
#pragma unroll
        for (int o = 0; o < CC_TINY_O; ++o)
          if (o < M.O) {
            if (p.noise && act) y[o] += __ldg(p.noise + b * (long long)p.T * M.O + (long long)k * M.O + o);
            yprev[o] = y[o];
            if (act) p.out[b * (long long)p.T * M.O + (long long)k * M.O + o] = y[o];
          }
      } else if (act) {
#pragma unroll
        for (int o = 0; o < CC_TINY_O; ++o)
          if (o < M.O) p.out[b * (long long)(p.T + 1) * M.O + (long long)(k + 1) * M.O + o] = y[o];
      }
    }
    if (!closed && p.x_final && act)
#pragma unroll
      for (int s = 0; s < SR; ++s)
        if (s < M.S) p.x_final[b * M.S + s] = x[s];
  }
  CC_TC_EPILOGUE()
}

// ------------------------------------------------------------------------------------------------
// K4 on tcgen05: reverse pass of the closed loop with a FROZEN LINEAR model (f, g single Linear layers with
// identity final activation) and the tiny trainable controller.  The model's adjoint needs no recomputation:
// per RK4 stage one MMA batch  [zbar ; vbar] = kbar . [W_x ; W_v]   (SURVEY 9.4), kbar built from the state
// adjoint lam (registers) and the previous stage's zbar.
// ------------------------------------------------------------------------------------------------
// epilogue of one backward stage for columns [C0, C0+NC): zb = D ;
//   generic (stages 4, 3, 2 of RK4): zsum += zb (zsum starts at 0) ; next kbar = ca*lam + cb*zb -> A
//   final   (stage 1, or the single Euler stage): lam <- lam + (zsum + zb) ; zsum <- 0
template <int SR, int C0, int NC, bool FINAL>
CC_DEV void cc_tc_bwd_group(uint32_t aD, uint32_t aHi, uint32_t aLo, float (&lam)[SR], float (&zsum)[SR], float ca, float cb) {
  uint32_t r[NC];
  cc_tc_ld<NC>(aD + C0, r);
  cc_tmem_wait_ld();
  float kb[NC];
  const float2 ca2 = make_float2(ca, ca), cb2 = make_float2(cb, cb);
#pragma unroll
  for (int i = 0; i < NC; i += 2) {
    const float2 zb = make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
    if (!FINAL) {
      const float2 zs = __fadd2_rn(make_float2(zsum[C0 + i], zsum[C0 + i + 1]), zb);
      zsum[C0 + i] = zs.x; zsum[C0 + i + 1] = zs.y;
      const float2 kk = __ffma2_rn(ca2, make_float2(lam[C0 + i], lam[C0 + i + 1]), __fmul2_rn(cb2, zb));
      kb[i] = kk.x; kb[i + 1] = kk.y;
    } else {
      const float2 l = __fadd2_rn(make_float2(lam[C0 + i], lam[C0 + i + 1]), __fadd2_rn(make_float2(zsum[C0 + i], zsum[C0 + i + 1]), zb));
      lam[C0 + i] = l.x; lam[C0 + i + 1] = l.y;
      zsum[C0 + i] = 0.f; zsum[C0 + i + 1] = 0.f;
    }
  }
  if (!FINAL) cc_tc_store_split<NC>(aHi, aLo, C0, kb);
}
template <int SR, bool FINAL>
CC_DEV void cc_tc_bwd_stage(uint32_t aD, uint32_t aHi, uint32_t aLo, float (&lam)[SR], float (&zsum)[SR], float ca, float cb, float (&utb)[CC_TINY_O]) {
  cc_tc_bwd_group<SR, 0, 32, FINAL>(aD, aHi, aLo, lam, zsum, ca, cb);
  if constexpr (SR == 56) {
    cc_tc_bwd_group<SR, 32, 16, FINAL>(aD, aHi, aLo, lam, zsum, ca, cb);
    cc_tc_bwd_group<SR, 48, 8, FINAL>(aD, aHi, aLo, lam, zsum, ca, cb);
  }
  uint32_t r[4];
  cc_tmem_ld4(aD + SR, r);  // vbar columns
  cc_tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < CC_TINY_O; ++i) utb[i] += __uint_as_float(r[i]);
}

template <int NCH, int CM, int CFG>
__global__ void __launch_bounds__(CC_TC_THREADS, 2) cc_tc_bwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  CC_TC_PROLOGUE(gp, true)
  const CcCtrlDims CD = cc_ctrl_dims<CFG>(C);
  const int nst = M.integrator == CC_INT_RK4 ? 4 : 1;
  {
    CcWarp w;
    cc_tc_make_warp(p, warp, lane, w);
    const bool act = lane < w.nact;
    const long long b = w.b0 + lane;
    const uint32_t aD = tbase + ((uint32_t)(warp * 32) << 16), aHi = aD + NP, aLo = aD + NP + KP;
    const float h = M.dt;
    const float* ttab = p.tape + (long long)p.n_ckpt * p.tape_rows * p.Bpad;
    float* GP = sh.gp + tid;  // lane-private columns: conflict free
    float* keep = sh.gp + (C.S + C.O) * CC_CKU * 128 + tid;  // controller stage record of the current step
    for (int e = 0; e < (C.S + C.O) * CC_CKU; ++e) GP[e * 128] = 0.f;
    float lamc[CC_TINY_S];  // adjoint of the controller state
#pragma unroll
    for (int i = 0; i < CC_TINY_S; ++i) lamc[i] = 0.f;
    float lam[SR], zsum[SR];
#pragma unroll
    for (int i = 0; i < SR; ++i) { lam[i] = 0.f; zsum[i] = 0.f; }
    float yfb[CC_TINY_O];
#pragma unroll
    for (int o = 0; o < CC_TINY_O; ++o) yfb[o] = 0.f;
    const long long yrow = (long long)p.T * M.O;
    float n_vin[CC_TINY_O], n_yb[CC_TINY_O], n_yp[CC_TINY_O], n_xc[CC_TINY_S];
    auto fetch = [&](int k) {
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) n_vin[i] = (i < p.R && act) ? __ldg(p.in + b * (long long)p.T * p.R + (long long)k * p.R + i) : 0.f;
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o) {
        n_yb[o] = (o < M.O && act) ? __ldg(p.ybar + b * yrow + (long long)k * M.O + o) : 0.f;
        n_yp[o] = (o < M.O && act) ? p.tape[((long long)k * p.tape_rows + p.tape_yrow + o) * p.Bpad + b] : 0.f;
      }
#pragma unroll
      for (int s = 0; s < CC_TINY_S; ++s) n_xc[s] = (s < C.S && act) ? p.tape[((long long)k * p.tape_rows + C.tape_row + s) * p.Bpad + b] : 0.f;
    };
    if (p.T > 0) fetch(p.T - 1);
    uint32_t par = 0;
    cc_tc_stagger(p);
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
      // first stage input of the model's reverse pass depends only on lam (and, post-step readout, on yb)
      const float* gw = sh.gw;
      const float os = M.out_scale;
      if (!M.pre) {  // y_k = g(x_{k+1}): its cotangent joins lam before the integrator's reverse pass
#pragma unroll
        for (int o = 0; o < CC_TINY_O; ++o)
          if (o < M.O) {
            const float d = os * yb[o];
#pragma unroll
            for (int i = 0; i < SR; ++i) lam[i] = fmaf(gw[o * (SR + 4) + i], d, lam[i]);
          }
      }
      {
        const float c0 = nst == 4 ? h / 6.f : h;
        float kb[SR];
#pragma unroll
        for (int i = 0; i < SR; ++i) kb[i] = c0 * lam[i];
#pragma unroll
        for (int c = 0; c < SR / 8; ++c) {
          float z[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) z[i] = kb[8 * c + i];
          cc_tc_store_split<8>(aHi, aLo, 8 * c, z);
        }
      }
      cc_tc_publish<NCH, true>(sh, q);
      // while the first MMA batch runs: recompute the controller step (its stage intermediates are kept)
      // recompute the controller step, its RK4 stages spread over the windows in which this warp does not issue (at most
      // two per window, staying at most one stage ahead), like the forward kernel does
      float a[CC_TINY_O];
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) a[i] = 0.f;
      if (C.pre) cc_ctrl_readout_pre(CD, sh.wgc, xcv, tc, vc, a);
      CcCtrlRun crun;
      const int nstc = cc_ctrl_nstages(CD);
      int cst = 0, win = 0;
      cc_ctrl_begin(CD, xcv, tc, vc, crun);
      if (warp != (int)((q.batch - 1u) & 3u))
        for (int c2 = 0; c2 < 2 && cst < nstc && cst < win + 2; ++c2) cc_ctrl_stage<true>(CD, sh.wfc, cst++, xcv, tc, crun, keep);
      float utb[CC_TINY_O];
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) utb[i] = 0.f;
#pragma unroll 1
      for (int st = 3; st >= 1 && nst == 4; --st) {  // after z4bar: kbar3 = h/3 lam + h z4bar ; z3bar: kbar2 = h/3 lam + h/2 z3bar ; z2bar: kbar1 = h/6 lam + h/2 z2bar
        cc_tc_await(sh, par);
        cc_tc_bwd_stage<SR, false>(aD, aHi, aLo, lam, zsum, st == 1 ? h / 6.f : h / 3.f, st == 3 ? h : h / 2.f, utb);
        cc_tc_publish<NCH, true>(sh, q);
        ++win;
        if (warp != (int)((q.batch - 1u) & 3u))
          for (int c2 = 0; c2 < 2 && cst < nstc && cst < win + 2; ++c2) cc_ctrl_stage<true>(CD, sh.wfc, cst++, xcv, tc, crun, keep);
      }
#pragma unroll 1
      for (; cst < nstc; ++cst) cc_ctrl_stage<true>(CD, sh.wfc, cst, xcv, tc, crun, keep);
      cc_ctrl_finish<true>(CD, sh.wgc, xcv, tc, crun, a, keep);
      cc_tc_await(sh, par);
      cc_tc_bwd_stage<SR, true>(aD, aHi, aLo, lam, zsum, 0.f, 0.f, utb);
      if (M.pre) {  // y_k = g(x_k): cotangent joins the adjoint of x_k
#pragma unroll
        for (int o = 0; o < CC_TINY_O; ++o)
          if (o < M.O) {
            const float d = os * yb[o];
#pragma unroll
            for (int i = 0; i < SR; ++i) lam[i] = fmaf(gw[o * (SR + 4) + i], d, lam[i]);
          }
      }
      // cotangent of the controller output = cotangent of the raw model input
      float ob[CC_TINY_O], vb[CC_TINY_O];
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) ob[i] = (i < M.I) ? utb[i] * cc_ut_grad(M.ut, M.u_scale, a[i]) : 0.f;
      cc_ctrl_bwd(CD, sh.wfc, sh.wgc, GP, tc, vc, keep, a, ob, lamc, vb);
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o) {
        yfb[o] = 0.f;
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i)
          if (o < M.O && i == p.R + o) yfb[o] = vb[i];
      }
      __syncwarp();
    }
    // gradient record of this warp tile (32 trajectories) -> workspace, reduced in fixed order afterwards
    // (32 lanes reduced with a fixed xor tree -> deterministic)
    {
      float* rec = p.gpart + w.wt * p.g_rec;
      for (int row = 0; row < C.S + C.O; ++row) {
        const int goff = row < C.S ? C.f.layer[0].g_off + row * C.K0 : C.g.layer[0].g_off + (row - C.S) * C.K0;
        for (int r = 0; r < C.K0; ++r) {
          float v = GP[(row * CC_CKU + cc_ctrl_col_of_row(C, r)) * 128];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(CC_FULL, v, off);
          if (lane == 0 && w.wt < p.n_wtiles) rec[C.g_base + goff + r] = v;
        }
      }
    }
  }
  CC_TC_EPILOGUE()
}
