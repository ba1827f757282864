// cc_mlp_kernels.cuh -- deep-MLP forward rollout on tcgen05 (kernel family 5; sm_100a).
//
// Open-loop unroll (AbstractModel.unroll, cc/core/abstract.py:54-70) of a neural-ODE model whose right-hand side f is a
// DEEP MLP (mlp_network.py:5-35, f_depth 1..3: the reference's default model neural_ode_model.py:27-36 and the width
// sweep of BASELINE.json configs[2]) -- the architectures the FFMA family serves at <= 0.45x of the FP32-FMA peak and not
// at all beyond width 64.  Same reference semantics as cc_kernels.cuh (neural_ode_model.py:106-149, compact :94-110,
// integrate.py:1-28); only the execution model differs:
//
//   * one CTA = one tile of 128 trajectories; thread = trajectory = TMEM lane.  The state x and the RK4 stage sum live in
//     the thread's registers for the whole horizon.
//   * EVERY LAYER of every stage evaluation is one batch of tcgen05.mma (cta_group::1, kind::tf32, M = 128, N = PAD,
//     K = 8 per instruction): D[128 x PAD] = A[128 x (PAD+8)] . B_l[PAD x (PAD+8)]^T with A = the layer input in TENSOR
//     MEMORY (written by the threads with tcgen05.st), B_l = layer l's weights resident in shared memory (K-major, no
//     swizzle, staged once per CTA from the flat theta), D in tensor memory.  fp32 accuracy: 3xTF32 split of both
//     operands, z.W = z_lo.W_hi + z_hi.W_lo + z_hi.W_hi (hi, lo rounded to nearest tf32; small terms first).
//   * A columns: [0, PAD) the layer input (state / hidden activations, zero beyond the logical width) and one tail chunk
//     [PAD, PAD+8) = [0 0 0 u~_0..u~_3 1] written once per step: layer 0's B carries the input weights in rows PAD+3+i,
//     every layer's B carries its bias in row PAD+7 (bias add and input injection are part of the MMA).
//   * the epilogue of a hidden layer is tcgen05.ld -> activation -> hi/lo split -> tcgen05.st of the next layer's input;
//     the epilogue of the last layer does the integrator bookkeeping in the reference's operation order
//     (integrate.py:6-13).  Two mbarriers (a_ready: 128 arrivals, d_ready: tcgen05.commit) are the only synchronisation;
//     the MMA-issuing role rotates over the four warps.  Two or three CTAs per SM overlap one tile's epilogue with another
//     tile's MMAs.
//   * readout g (single Linear layer, O <= 4) = FFMA dot products of the thread against broadcast weight rows.
//
// TMEM map (power-of-two allocation >= 3 PAD + 16): D [0, PAD), A_hi [PAD, 2 PAD + 8), A_lo [2 PAD + 8, 3 PAD + 16).
// The thread-level code is the same source for the CPU emulation of tests/emu (TMEM = an array in emulated shared
// memory, the MMA batch = a loop run by thread 0 between two barriers with the same tf32 operand truncation).
#pragma once
#include "cc_kernels.cuh"
#include "cc_mlp.h"
#ifndef CC_EMULATE
#include "cc_tmem.cuh"
#endif

#define CC_MLP_THREADS 128

CC_DEV float cc_mlp_tf32(float v) {  // truncation (what the tensor core does with a 32-bit operand)
  unsigned u;
  memcpy(&u, &v, 4);
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}
CC_DEV float cc_mlp_tf32_rn(float v) {  // round to nearest tf32
  unsigned u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

template <int PAD>
struct CcMlpTm {
  static constexpr int NP = PAD, KP = PAD + 8, COL_D = 0, COL_AHI = PAD, COL_ALO = 2 * PAD + 8, COLS = 3 * PAD + 16;
  static constexpr int TMEM_ALLOC = COLS <= 64 ? 64 : (COLS <= 128 ? 128 : (COLS <= 256 ? 256 : 512));
  static constexpr int LAYER_FLOATS = 2 * KP * NP;  // B_hi then B_lo of one layer
#ifdef CC_EMULATE
  float* tm;       // [COLS][128]
  const float* b;  // layer 0's B_hi
  int tid;
#else
  uint32_t lane_base, tbase, b_u32;
  uint64_t* a_ready;
  uint64_t* d_ready;
  uint32_t batch, par;
  int warp, lane;
#endif
};

#ifdef CC_EMULATE
template <int PAD, int NC>
CC_DEV void cc_mlp_ld(const CcMlpTm<PAD>& t, int col, float (&r)[NC]) {
  for (int i = 0; i < NC; ++i) r[i] = t.tm[(col + i) * 128 + t.tid];
}
template <int PAD, int NC>
CC_DEV void cc_mlp_st(const CcMlpTm<PAD>& t, int col, const float (&r)[NC]) {
  for (int i = 0; i < NC; ++i) t.tm[(col + i) * 128 + t.tid] = r[i];
}
CC_DEV void cc_mlp_wait_ld() {}
template <int PAD>
CC_DEV void cc_mlp_round(CcMlpTm<PAD>& t, int layer) {
  using TM = CcMlpTm<PAD>;
  __syncthreads();
  if (t.tid == 0) {
    const float* bhi = t.b + (size_t)layer * TM::LAYER_FLOATS;
    const float* blo = bhi + TM::KP * TM::NP;
    for (int l = 0; l < 128; ++l)
      for (int n = 0; n < TM::NP; ++n) {
        float acc = 0.f;
        for (int pass = 0; pass < 3; ++pass)
          for (int k = 0; k < TM::KP; ++k) {
            const int off = ((k >> 2) * TM::NP + n) * 4 + (k & 3);
            const float ahi = cc_mlp_tf32(t.tm[(TM::COL_AHI + k) * 128 + l]), alo = cc_mlp_tf32(t.tm[(TM::COL_ALO + k) * 128 + l]);
            const float bh = cc_mlp_tf32(bhi[off]), bl = cc_mlp_tf32(blo[off]);
            acc += pass == 0 ? alo * bh : (pass == 1 ? ahi * bl : ahi * bh);
          }
        t.tm[(TM::COL_D + n) * 128 + l] = acc;
      }
  }
  __syncthreads();
}
template <int PAD>
CC_DEV void cc_mlp_await(CcMlpTm<PAD>&) {}
#else
template <int PAD, int NC>
CC_DEV void cc_mlp_ld(const CcMlpTm<PAD>& t, int col, float (&r)[NC]) {
  uint32_t u[NC];
  if constexpr (NC == 32) cc_tmem_ld32(t.lane_base + col, u);
  else if constexpr (NC == 16) cc_tmem_ld16(t.lane_base + col, u);
  else if constexpr (NC == 4) cc_tmem_ld4(t.lane_base + col, u);
  else cc_tmem_ld8(t.lane_base + col, u);
#pragma unroll
  for (int i = 0; i < NC; ++i) r[i] = __uint_as_float(u[i]);
}
template <int PAD, int NC>
CC_DEV void cc_mlp_st(const CcMlpTm<PAD>& t, int col, const float (&r)[NC]) {
  uint32_t u[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) u[i] = __float_as_uint(r[i]);
  if constexpr (NC == 32) cc_tmem_st32(t.lane_base + col, u);
  else if constexpr (NC == 16) cc_tmem_st16(t.lane_base + col, u);
  else cc_tmem_st8(t.lane_base + col, u);
}
CC_DEV void cc_mlp_wait_ld() { cc_tmem_wait_ld(); }
// D = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi over all (PAD+8)/8 contraction chunks; issued by ONE thread, compile-time shapes
template <int PAD>
CC_DEV void cc_mlp_issue(uint32_t tbase, uint32_t bhi, uint32_t blo) {
  using TM = CcMlpTm<PAD>;
  constexpr int NA = TM::KP / 8;
  constexpr uint32_t kstep = 2u * TM::NP * 16u, lbo = TM::NP * 16u;
  constexpr uint32_t idesc = cc_umma_idesc_tf32(128, TM::NP);
  const uint32_t aHi = tbase + TM::COL_AHI, aLo = tbase + TM::COL_ALO;
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(tbase, aLo + 8 * js, cc_umma_desc_kmajor(bhi + js * kstep, lbo, 128), idesc, js ? 1u : 0u);
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(tbase, aHi + 8 * js, cc_umma_desc_kmajor(blo + js * kstep, lbo, 128), idesc, 1u);
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(tbase, aHi + 8 * js, cc_umma_desc_kmajor(bhi + js * kstep, lbo, 128), idesc, 1u);
}
// publish the A row just written and start layer `layer`'s MMA batch: all 128 threads arrive; the issuing warp (rotating)
// waits for the last arrival, its lane 0 issues the batch and commits onto d_ready
template <int PAD>
CC_DEV void cc_mlp_round(CcMlpTm<PAD>& t, int layer) {
  using TM = CcMlpTm<PAD>;
  cc_tmem_wait_st();
  cc_tc_fence_before();
  cc_mbar_arrive(t.a_ready);
  const uint32_t batch = t.batch++;
  if (t.warp == (int)(batch & 3u)) {
    cc_mbar_wait(t.a_ready, batch & 1u);
    cc_tc_fence_after();
    if (t.lane == 0) {
      const uint32_t bhi = t.b_u32 + (uint32_t)layer * (uint32_t)(TM::LAYER_FLOATS * 4);
      cc_mlp_issue<PAD>(t.tbase, bhi, bhi + (uint32_t)(TM::KP * TM::NP * 4));
      cc_tc_commit(t.d_ready);
    }
    __syncwarp();
  }
}
template <int PAD>
CC_DEV void cc_mlp_await(CcMlpTm<PAD>& t) {
  cc_mbar_wait(t.d_ready, t.par);
  t.par ^= 1;
  cc_tc_fence_after();
}
#endif

// split NC values into tf32 hi / lo and store them at columns c0.. of A_hi / A_lo
template <int PAD, int NC>
CC_DEV void cc_mlp_store_split(const CcMlpTm<PAD>& t, int c0, const float (&z)[NC]) {
  float hi[NC], lo[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    // hi rounded to nearest tf32 (unbiased main term); lo = z - hi is exact (<= 13 significant bits, |lo| <= 2^-12 |z|) and the
    // tensor core truncates it to tf32 itself: <= 2^-22 |z| of one-sided error, two instructions per element less than rounding it
    hi[i] = cc_mlp_tf32_rn(z[i]);
    lo[i] = z[i] - hi[i];
  }
  cc_mlp_st<PAD, NC>(t, CcMlpTm<PAD>::COL_AHI + c0, hi);
  cc_mlp_st<PAD, NC>(t, CcMlpTm<PAD>::COL_ALO + c0, lo);
}

// stage one layer's B operand (hi / lo) : element (n, k) at ((k/4)*NP + n)*4 + k%4 (8 x 16-byte core matrices, SBO = 128 B,
// LBO = NP*16 B).  k < in_log: the weight W[n][k]; layer 0: the u rows sit at k = PAD+3+i; every layer: bias at k = PAD+7.
template <int PAD>
CC_DEV void cc_mlp_stage_layer(const CcMlpArgs& a, int l, float* bhi, float* blo, int tid, int nthreads) {
  using TM = CcMlpTm<PAD>;
  const CcMlpLayerArg& L = a.f[l];
  const int kin = l == 0 ? a.S : L.in_log;  // layer 0: the state columns; its input columns live in the tail chunk
  for (int i = tid; i < TM::KP * TM::NP; i += nthreads) {
    const int n = i / TM::KP, k = i % TM::KP;
    float w = 0.f;
    if (n < L.N) {
      if (k < kin) w = __ldg(a.theta + L.w_off + n * L.in_log + k);
      else if (l == 0 && k >= PAD + 3 && k < PAD + 3 + a.I) w = __ldg(a.theta + L.w_off + n * L.in_log + a.S + (k - PAD - 3));
      else if (k == PAD + 7 && L.b_off >= 0) w = __ldg(a.theta + L.b_off + n);
    }
    const float hi = cc_mlp_tf32_rn(w);
    const int off = ((k >> 2) * TM::NP + n) * 4 + (k & 3);
    bhi[off] = hi;
    blo[off] = cc_mlp_tf32_rn(w - hi);
  }
}

// y = out_scale * act(g . x + b): dot products against broadcast weight rows gw[o][PAD + 4] (zero beyond S, bias at PAD)
template <int PAD>
CC_DEV void cc_mlp_readout(const float* gw, const float (&x)[PAD], int O, int gact, float out_scale, float (&y)[CC_TINY_O]) {
#pragma unroll
  for (int o = 0; o < CC_TINY_O; ++o) {
    y[o] = 0.f;
    if (o < O) {
      const float* row = gw + o * (PAD + 4);
      float s0 = row[PAD], s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < PAD; i += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(row + i);
        s0 = fmaf(wv.x, x[i], s0); s1 = fmaf(wv.y, x[i + 1], s1); s2 = fmaf(wv.z, x[i + 2], s2); s3 = fmaf(wv.w, x[i + 3], s3);
      }
      y[o] = out_scale * cc_act(gact, (s0 + s1) + (s2 + s3));
    }
  }
}

// activation of a whole column group: the kind is uniform, so the dispatch happens once per group, not per element
template <int NC>
CC_DEV void cc_mlp_act_group(int act, float (&r)[NC]) {
  if (act == CC_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < NC; ++i) r[i] = fmaxf(r[i], 0.f);
  } else if (act == CC_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < NC; ++i) r[i] = tanhf(r[i]);
  }
}

// hidden layer: h = act(D) -> next layer's input
template <int PAD>
CC_DEV void cc_mlp_hidden(const CcMlpTm<PAD>& t, int act) {
  constexpr int G = PAD < 32 ? PAD : 32;
#pragma unroll
  for (int c0 = 0; c0 < PAD; c0 += G) {
    float r[G];
    cc_mlp_ld<PAD, G>(t, CcMlpTm<PAD>::COL_D + c0, r);
    cc_mlp_wait_ld();
    cc_mlp_act_group<G>(act, r);
    cc_mlp_store_split<PAD, G>(t, c0, r);
  }
}
// last layer of a stage: k = act_final(D), then the integrator bookkeeping (integrate.py:6-13)
//   not FINAL (RK4 stages 1..3): acc += wk * k ; next stage input z = x + (h*k)*cz -> A
//   FINAL: RK4  x <- x + h*(1/6*(acc + k)) ;  Euler x <- x + h*k ; no-integrate x <- k
template <int PAD, bool FINAL>
CC_DEV void cc_mlp_last(const CcMlpTm<PAD>& t, int act, float (&x)[PAD], float (&acc)[PAD], float h, float wk, float cz, int integrator) {
  constexpr int G = PAD < 32 ? PAD : (PAD >= 64 ? 16 : 32);  // 64 columns of x and acc are live: small groups keep the temporaries in registers
#pragma unroll
  for (int c0 = 0; c0 < PAD; c0 += G) {
    float r[G];
    cc_mlp_ld<PAD, G>(t, CcMlpTm<PAD>::COL_D + c0, r);
    cc_mlp_wait_ld();
    cc_mlp_act_group<G>(act, r);
    if (!FINAL) {
      float z[G];
#pragma unroll
      for (int i = 0; i < G; ++i) {
        acc[c0 + i] = fmaf(wk, r[i], acc[c0 + i]);
        z[i] = fmaf(h * r[i], cz, x[c0 + i]);
      }
      cc_mlp_store_split<PAD, G>(t, c0, z);
    } else if (integrator == CC_INT_RK4) {
#pragma unroll
      for (int i = 0; i < G; ++i) {
        x[c0 + i] = fmaf(h, (1.f / 6.f) * (acc[c0 + i] + r[i]), x[c0 + i]);
        acc[c0 + i] = 0.f;
      }
    } else if (integrator == CC_INT_EE) {
#pragma unroll
      for (int i = 0; i < G; ++i) x[c0 + i] = fmaf(h, r[i], x[c0 + i]);
    } else {
#pragma unroll
      for (int i = 0; i < G; ++i) x[c0 + i] = r[i];
    }
  }
}
template <int PAD>
CC_DEV void cc_mlp_write_state(const CcMlpTm<PAD>& t, const float (&x)[PAD]) {
  constexpr int G = PAD < 32 ? PAD : 32;
#pragma unroll
  for (int c0 = 0; c0 < PAD; c0 += G) {
    float z[G];
#pragma unroll
    for (int i = 0; i < G; ++i) z[i] = x[c0 + i];
    cc_mlp_store_split<PAD, G>(t, c0, z);
  }
}

#define CC_MLP_SMEM_FLOATS(PAD, L) ((L) * CcMlpTm<PAD>::LAYER_FLOATS + CC_TINY_O * ((PAD) + 4) + 8)

// SEG = segment replay of the checkpointed reverse pass (steps [k0, k1) from the checkpoint carry_in, segment-local tape, no
// outputs); a template parameter so that the plain forward pass keeps its register allocation
template <int PAD, bool SEG>
__global__ void __launch_bounds__(CC_MLP_THREADS, PAD >= 64 ? 2 : 3) cc_mlp_fwd_kernel(const CC_GRID_CONSTANT CcMlpArgs a) {
  using TM = CcMlpTm<PAD>;
  float* smem = CC_SMEM_PTR;
  const int tid = threadIdx.x;
  float* bmat = smem;
  float* gw = smem + a.L * TM::LAYER_FLOATS;
  for (int l = 0; l < a.L; ++l) cc_mlp_stage_layer<PAD>(a, l, bmat + l * TM::LAYER_FLOATS, bmat + l * TM::LAYER_FLOATS + TM::KP * TM::NP, tid, CC_MLP_THREADS);
  for (int i = tid; i < CC_TINY_O * (PAD + 4); i += CC_MLP_THREADS) {
    const int o = i / (PAD + 4), s = i % (PAD + 4);
    float w = 0.f;
    if (o < a.O) {
      if (s < a.S) w = __ldg(a.theta + a.g.w_off + o * a.g.in_log + s);
      else if (s == PAD && a.g.b_off >= 0) w = __ldg(a.theta + a.g.b_off + o);
    }
    gw[i] = w;
  }
  TM t;
#ifdef CC_EMULATE
  t.tm = gw + CC_TINY_O * (PAD + 4) + 8;
  t.b = bmat;
  t.tid = tid;
  __syncthreads();
#else
  const int warp = tid >> 5, lane = tid & 31;
  uint64_t* a_ready = reinterpret_cast<uint64_t*>(gw + CC_TINY_O * (PAD + 4));
  uint64_t* d_ready = a_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_ready + 1);
  if (warp == 0) {
    if (lane == 0) {
      cc_mbar_init(a_ready, CC_MLP_THREADS);
      cc_mbar_init(d_ready, 1);
      cc_mbar_fence_init();
    }
    __syncwarp();
    cc_tmem_alloc(tmem_slot, TM::TMEM_ALLOC);
  }
  cc_fence_proxy_async();  // the staged weights (generic proxy) become visible to the tensor core (async proxy)
  cc_tc_fence_before();
  __syncthreads();
  cc_tc_fence_after();
  t.tbase = *tmem_slot;
  t.lane_base = t.tbase + ((uint32_t)(warp * 32) << 16);
  t.b_u32 = cc_smem_u32(bmat);
  t.a_ready = a_ready; t.d_ready = d_ready;
  t.batch = 0; t.par = 0; t.warp = warp; t.lane = lane;
#endif
  {
    const long long b = (long long)blockIdx.x * 128 + tid;
    const bool act = b < a.B;
    const float h = a.dt;
    const int nst = a.integrator == CC_INT_RK4 ? 4 : 1;
    float x[PAD], acc[PAD];
#pragma unroll
    for (int i = 0; i < PAD; ++i) {
      if (SEG && a.carry_in) x[i] = (act && i < a.S) ? a.carry_in[(size_t)i * a.Bpad + b] : 0.f;
      else x[i] = (a.x0 && act && i < a.S) ? __ldg(a.x0 + b * a.S + i) : 0.f;
      acc[i] = 0.f;
    }
    const int kbeg = SEG ? a.k0 : 0, kend = SEG ? a.k1 : a.T;
    const bool wy = !SEG || a.y != nullptr;
    // y0 = g(x0)      neural_ode_model.py:160-175
    float y[CC_TINY_O];
    cc_mlp_readout<PAD>(gw, x, a.O, a.g.act, a.out_scale, y);
    if (act && wy && kbeg == 0)
      for (int o = 0; o < a.O; ++o) a.y[b * (long long)(a.T + 1) * a.O + o] = y[o];
    float vin[CC_TINY_O];
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) vin[i] = (i < a.I && act && kend > kbeg) ? __ldg(a.u + b * (long long)a.T * a.I + (long long)kbeg * a.I + i) : 0.f;
    for (int k = kbeg; k < kend; ++k) {
      if (!SEG && a.tape && blockIdx.x == 0 && tid == 0) a.ttab[k] = 0.f;
      float vt[CC_TINY_O];
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) vt[i] = (i < a.I) ? cc_ut(a.ut, a.u_scale, vin[i]) : 0.f;
      if (k + 1 < kend) {
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i)
          if (i < a.I && act) vin[i] = __ldg(a.u + b * (long long)a.T * a.I + (long long)(k + 1) * a.I + i);
      }
      if (a.tape && act && (SEG || k % a.ckpt_every == 0)) {  // open loop: the pre-step state of every checkpointed step is the tape
        const long long ck = SEG ? k - kbeg : k / a.ckpt_every;
#pragma unroll
        for (int s = 0; s < PAD; ++s)
          if (s < a.S) a.tape[(ck * a.S + s) * a.Bpad + b] = x[s];
      }
      cc_mlp_write_state<PAD>(t, x);
      {
        float tail[8];
        tail[0] = 0.f; tail[1] = 0.f; tail[2] = 0.f; tail[3] = vt[0]; tail[4] = vt[1]; tail[5] = vt[2]; tail[6] = vt[3]; tail[7] = 1.f;
        cc_mlp_store_split<PAD, 8>(t, PAD, tail);
      }
      cc_mlp_round<PAD>(t, 0);
      if (a.pre) {  // compact model: y = g(x) BEFORE the update (neural_ode_model_compact_example.py:94-110); overlaps the MMAs
        cc_mlp_readout<PAD>(gw, x, a.O, a.g.act, a.out_scale, y);
        if (act && wy)
          for (int o = 0; o < a.O; ++o) a.y[b * (long long)(a.T + 1) * a.O + (long long)(k + 1) * a.O + o] = y[o];
      }
#pragma unroll 1
      for (int st = 0; st < nst; ++st) {
#pragma unroll 1
        for (int l = 0; l < a.L; ++l) {
          cc_mlp_await<PAD>(t);
          if (l < a.L - 1) {
            cc_mlp_hidden<PAD>(t, a.f[l].act);
            cc_mlp_round<PAD>(t, l + 1);
          } else if (st < nst - 1) {
            cc_mlp_last<PAD, false>(t, a.f[l].act, x, acc, h, st == 0 ? 1.f : 2.f, st == 2 ? 1.f : 0.5f, a.integrator);
            cc_mlp_round<PAD>(t, 0);
          } else {
            cc_mlp_last<PAD, true>(t, a.f[l].act, x, acc, h, 0.f, 0.f, a.integrator);
          }
        }
      }
      if (!a.pre) {  // full model: y = g(x_next)      neural_ode_model.py:131-137
        cc_mlp_readout<PAD>(gw, x, a.O, a.g.act, a.out_scale, y);
        if (act && wy)
          for (int o = 0; o < a.O; ++o) a.y[b * (long long)(a.T + 1) * a.O + (long long)(k + 1) * a.O + o] = y[o];
      }
    }
    if (a.x_final && act)
#pragma unroll
      for (int s = 0; s < PAD; ++s)
        if (s < a.S) a.x_final[b * a.S + s] = x[s];
    if (a.tape_xT && act)  // final state for the tcgen05 reverse pass (cc_mlpb_kernels.cuh)
#pragma unroll
      for (int s = 0; s < PAD; ++s)
        if (s < a.S) a.tape_xT[(size_t)s * a.Bpad + b] = x[s];
  }
#ifndef CC_EMULATE
  cc_tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == 0) cc_tmem_dealloc(t.tbase, TM::TMEM_ALLOC);
#endif
}
