// cc_mlp2_kernels.cuh -- the deep-MLP tcgen05 forward kernel (cc_mlp_kernels.cuh) with TWO THREADS PER TRAJECTORY.
//
// Same algorithm, TMEM map and shared-memory layout as cc_mlp_fwd_kernel<PAD> (PAD = 32 / 64, weights resident); what
// changes is the worker side of a round.  ncu on the one-thread-per-trajectory kernel (profiles/r02b_mlp_ncu_full.md):
// tensor pipe 35 % active, 8 warps per SM, ~1000 instructions per warp and round -- the epilogue of a tile (tcgen05.ld,
// activation, hi/lo split, tcgen05.st of 64 columns per thread) takes twice as long as the tile's MMA batch, so two
// resident tiles cannot keep the pipe busy.  Here a CTA has 8 warps: warp w and w + 4 share the TMEM lanes 32 (w % 4) ..
// (a warp can only access its own lane quarter, whichever of the two it is) and thread half hh = w / 4 owns columns
// [hh PAD/2, (hh+1) PAD/2) of the state, the stage sum and every epilogue: half the instructions per warp and round, 16
// warps per SM.  The readout's two partial dot products meet through 8 spare TMEM columns of the lane (no shared memory:
// two CTAs of the 64-wide variant fill the SM's 227 KB): half 1 stores its partial before it publishes a round, half 0
// loads it behind that round's d_ready and writes y -- outputs trail by one round, the last one by a CTA barrier.
#pragma once
#include "cc_mlp_kernels.cuh"

#define CC_MLP2_THREADS 256

template <int PAD>
struct CcMlp2 {
  static constexpr int PC = PAD / 2;                   // columns per thread
  static constexpr int G = PC <= 16 ? PC : 8;          // epilogue group (32 columns of x and acc live in 128 registers: small groups)
  static constexpr int COL_Y = CcMlpTm<PAD>::COLS;     // 2 slots x 4 spare columns for the readout partials
  static_assert(CcMlpTm<PAD>::COLS + 8 <= CcMlpTm<PAD>::TMEM_ALLOC, "no spare TMEM columns");
};

#ifndef CC_EMULATE
// publish with 256 arrivals; the issuing role rotates over the eight warps
template <int PAD>
CC_DEV void cc_mlp2_round(CcMlpTm<PAD>& t, int layer) {
  using TM = CcMlpTm<PAD>;
  cc_tmem_wait_st();
  cc_tc_fence_before();
  cc_mbar_arrive(t.a_ready);
  const uint32_t batch = t.batch++;
  if (t.warp == (int)(batch & 7u)) {
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
#else
template <int PAD>
CC_DEV void cc_mlp2_round(CcMlpTm<PAD>& t, int layer) {
  using TM = CcMlpTm<PAD>;
  __syncthreads();
  if (threadIdx.x == 0) {
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
#endif

template <int PAD>
CC_DEV void cc_mlp2_hidden(const CcMlpTm<PAD>& t, int act, int cb) {
  constexpr int PC = CcMlp2<PAD>::PC, G = CcMlp2<PAD>::G;
#pragma unroll
  for (int c0 = 0; c0 < PC; c0 += G) {
    float r[G];
    cc_mlp_ld<PAD, G>(t, CcMlpTm<PAD>::COL_D + cb + c0, r);
    cc_mlp_wait_ld();
    cc_mlp_act_group<G>(act, r);
    cc_mlp_store_split<PAD, G>(t, cb + c0, r);
  }
}
template <int PAD, bool FINAL>
CC_DEV void cc_mlp2_last(const CcMlpTm<PAD>& t, int act, int cb, float (&x)[PAD / 2], float (&acc)[PAD / 2], float h, float wk, float cz,
                         int integrator) {
  constexpr int PC = CcMlp2<PAD>::PC, G = CcMlp2<PAD>::G;
#pragma unroll
  for (int c0 = 0; c0 < PC; c0 += G) {
    float r[G];
    cc_mlp_ld<PAD, G>(t, CcMlpTm<PAD>::COL_D + cb + c0, r);
    cc_mlp_wait_ld();
    cc_mlp_act_group<G>(act, r);
    if (!FINAL) {
      float z[G];
#pragma unroll
      for (int i = 0; i < G; ++i) {
        acc[c0 + i] = fmaf(wk, r[i], acc[c0 + i]);
        z[i] = fmaf(h * r[i], cz, x[c0 + i]);
      }
      cc_mlp_store_split<PAD, G>(t, cb + c0, z);
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
// partial readout over this half's columns: gw[o][PAD + 4] (zero beyond S, bias at index PAD: added by half 0)
template <int PAD>
CC_DEV void cc_mlp2_readout_part(const float* gw, const float (&x)[PAD / 2], int hh, int O, float (&y)[CC_TINY_O]) {
  constexpr int PC = PAD / 2;
#pragma unroll
  for (int o = 0; o < CC_TINY_O; ++o) {
    y[o] = 0.f;
    if (o < O) {
      const float* row = gw + o * (PAD + 4) + hh * PC;
      float s0 = hh == 0 ? gw[o * (PAD + 4) + PAD] : 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < PC; i += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(row + i);
        s0 = fmaf(wv.x, x[i], s0); s1 = fmaf(wv.y, x[i + 1], s1); s2 = fmaf(wv.z, x[i + 2], s2); s3 = fmaf(wv.w, x[i + 3], s3);
      }
      y[o] = (s0 + s1) + (s2 + s3);
    }
  }
}

// SEG = segment replay of the checkpointed reverse pass (steps [k0, k1) from the checkpoint carry_in, segment-local tape, no
// outputs); a template parameter so that the plain forward pass keeps its register allocation (128 registers, two CTAs per SM)
template <int PAD, bool SEG>
__global__ void __launch_bounds__(CC_MLP2_THREADS, 2) cc_mlp2_fwd_kernel(const CC_GRID_CONSTANT CcMlpArgs a) {
  using TM = CcMlpTm<PAD>;
  constexpr int PC = CcMlp2<PAD>::PC, G = CcMlp2<PAD>::G, COL_Y = CcMlp2<PAD>::COL_Y;
  float* smem = CC_SMEM_PTR;
  const int tid = threadIdx.x;
  float* bmat = smem;
  float* gw = smem + a.L * TM::LAYER_FLOATS;
  for (int l = 0; l < a.L; ++l) cc_mlp_stage_layer<PAD>(a, l, bmat + l * TM::LAYER_FLOATS, bmat + l * TM::LAYER_FLOATS + TM::KP * TM::NP, tid, CC_MLP2_THREADS);
  for (int i = tid; i < CC_TINY_O * (PAD + 4); i += CC_MLP2_THREADS) {
    const int o = i / (PAD + 4), s = i % (PAD + 4);
    float w = 0.f;
    if (o < a.O) {
      if (s < a.S) w = __ldg(a.theta + a.g.w_off + o * a.g.in_log + s);
      else if (s == PAD && a.g.b_off >= 0) w = __ldg(a.theta + a.g.b_off + o);
    }
    gw[i] = w;
  }
  TM t;
  const int row = tid & 127, hh = tid >> 7, cb = hh * PC;
#ifdef CC_EMULATE
  t.tm = gw + CC_TINY_O * (PAD + 4) + 8;
  t.b = bmat;
  t.tid = row;
  __syncthreads();
#else
  const int warp = tid >> 5, lane = tid & 31;
  uint64_t* a_ready = reinterpret_cast<uint64_t*>(gw + CC_TINY_O * (PAD + 4));
  uint64_t* d_ready = a_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_ready + 1);
  if (warp == 0) {
    if (lane == 0) {
      cc_mbar_init(a_ready, CC_MLP2_THREADS);
      cc_mbar_init(d_ready, 1);
      cc_mbar_fence_init();
    }
    __syncwarp();
    cc_tmem_alloc(tmem_slot, TM::TMEM_ALLOC);
  }
  cc_fence_proxy_async();
  cc_tc_fence_before();
  __syncthreads();
  cc_tc_fence_after();
  t.tbase = *tmem_slot;
  t.lane_base = t.tbase + ((uint32_t)((warp & 3) * 32) << 16);
  t.b_u32 = cc_smem_u32(bmat);
  t.a_ready = a_ready; t.d_ready = d_ready;
  t.batch = 0; t.par = 0; t.warp = warp; t.lane = lane;
#endif
  {
    const long long b = (long long)blockIdx.x * 128 + row;
    const bool act = b < a.B;
    const float h = a.dt;
    const int nst = a.integrator == CC_INT_RK4 ? 4 : 1;
    float x[PC], acc[PC];
#pragma unroll
    for (int i = 0; i < PC; ++i) {
      if (SEG && a.carry_in) x[i] = (act && cb + i < a.S) ? a.carry_in[(size_t)(cb + i) * a.Bpad + b] : 0.f;
      else x[i] = (a.x0 && act && cb + i < a.S) ? __ldg(a.x0 + b * a.S + cb + i) : 0.f;
      acc[i] = 0.f;
    }
    const int kbeg = SEG ? a.k0 : 0, kend = SEG ? a.k1 : a.T;
    // deferred readout: stash(slot) = partial dot products of this half (half 1 -> TMEM slot, half 0 -> registers);
    // flush(slot, krow) = half 0 adds half 1's partial and writes y[krow]; a flush sits behind the d_ready of a round that was
    // published after the stash (or behind the CTA barrier at the end)
    float ymine[2][CC_TINY_O];
    auto stash = [&](int slot) {
      float yp[CC_TINY_O];
      cc_mlp2_readout_part<PAD>(gw, x, hh, a.O, yp);
      if (hh == 1) {
        float q[4];
        q[0] = yp[0]; q[1] = yp[1]; q[2] = yp[2]; q[3] = yp[3];
#ifdef CC_EMULATE
        for (int i = 0; i < 4; ++i) t.tm[(COL_Y + 4 * slot + i) * 128 + row] = q[i];
#else
        uint32_t u[4];
        for (int i = 0; i < 4; ++i) u[i] = __float_as_uint(q[i]);
        cc_tmem_st4(t.lane_base + COL_Y + 4 * slot, u);
#endif
      } else {
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i) ymine[slot][i] = yp[i];
      }
    };
    auto flush = [&](int slot, long long krow) {
      if (hh == 0) {
        float q[4];
#ifdef CC_EMULATE
        for (int i = 0; i < 4; ++i) q[i] = t.tm[(COL_Y + 4 * slot + i) * 128 + row];
#else
        uint32_t u[4];
        cc_tmem_ld4(t.lane_base + COL_Y + 4 * slot, u);
        cc_tmem_wait_ld();
        for (int i = 0; i < 4; ++i) q[i] = __uint_as_float(u[i]);
#endif
        if (act && (!SEG || a.y))
#pragma unroll
          for (int o = 0; o < CC_TINY_O; ++o)
            if (o < a.O) a.y[b * (long long)(a.T + 1) * a.O + krow * a.O + o] = a.out_scale * cc_act(a.g.act, ymine[slot][o] + q[o]);
      }
    };
    stash(0);  // y0 = g(x0)      neural_ode_model.py:160-175
    long long pend0 = kbeg == 0 ? 0 : -1;  // row waiting in slot 0 (-1: none)
    float vin[CC_TINY_O];
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) vin[i] = (hh == 0 && i < a.I && act && kend > kbeg) ? __ldg(a.u + b * (long long)a.T * a.I + (long long)kbeg * a.I + i) : 0.f;
    for (int k = kbeg; k < kend; ++k) {
      if (!SEG && a.tape && blockIdx.x == 0 && tid == 0) a.ttab[k] = 0.f;
      if (a.tape && act && (SEG || k % a.ckpt_every == 0)) {
        const long long ck = SEG ? k - kbeg : k / a.ckpt_every;
#pragma unroll
        for (int s = 0; s < PC; ++s)
          if (cb + s < a.S) a.tape[(ck * a.S + cb + s) * a.Bpad + b] = x[s];
      }
#pragma unroll
      for (int c0 = 0; c0 < PC; c0 += G) {
        float z[G];
#pragma unroll
        for (int i = 0; i < G; ++i) z[i] = x[c0 + i];
        cc_mlp_store_split<PAD, G>(t, cb + c0, z);
      }
      if (hh == 0) {
        float vt[CC_TINY_O];
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i) vt[i] = (i < a.I) ? cc_ut(a.ut, a.u_scale, vin[i]) : 0.f;
        if (k + 1 < kend) {
#pragma unroll
          for (int i = 0; i < CC_TINY_O; ++i)
            if (i < a.I && act) vin[i] = __ldg(a.u + b * (long long)a.T * a.I + (long long)(k + 1) * a.I + i);
        }
        float tail[8];
        tail[0] = 0.f; tail[1] = 0.f; tail[2] = 0.f; tail[3] = vt[0]; tail[4] = vt[1]; tail[5] = vt[2]; tail[6] = vt[3]; tail[7] = 1.f;
        cc_mlp_store_split<PAD, 8>(t, PAD, tail);
      }
      if (a.pre) stash(1);  // compact model: y_{k+1} = g(x_k), the state BEFORE the update
      cc_mlp2_round<PAD>(t, 0);
#pragma unroll 1
      for (int st = 0; st < nst; ++st) {
#pragma unroll 1
        for (int l = 0; l < a.L; ++l) {
          cc_mlp_await<PAD>(t);
          if (st == 0 && l == 0) {  // the first round of the step was published after the stashes: their partials are visible
            if (pend0 >= 0) { flush(0, pend0); pend0 = -1; }
            if (a.pre) flush(1, k + 1);
          }
          if (l < a.L - 1) {
            cc_mlp2_hidden<PAD>(t, a.f[l].act, cb);
            cc_mlp2_round<PAD>(t, l + 1);
          } else if (st < nst - 1) {
            cc_mlp2_last<PAD, false>(t, a.f[l].act, cb, x, acc, h, st == 0 ? 1.f : 2.f, st == 2 ? 1.f : 0.5f, a.integrator);
            cc_mlp2_round<PAD>(t, 0);
          } else {
            cc_mlp2_last<PAD, true>(t, a.f[l].act, cb, x, acc, h, 0.f, 0.f, a.integrator);
          }
        }
      }
      if (!a.pre) {  // full model: y_{k+1} = g(x_{k+1}); flushed behind the next step's first round
        stash(0);
        pend0 = k + 1;
      }
    }
    if (a.x_final && act)
#pragma unroll
      for (int s = 0; s < PC; ++s)
        if (cb + s < a.S) a.x_final[b * a.S + cb + s] = x[s];
    if (a.tape_xT && act)
#pragma unroll
      for (int s = 0; s < PC; ++s)
        if (cb + s < a.S) a.tape_xT[(size_t)(cb + s) * a.Bpad + b] = x[s];
    // the last stash has no round behind it: CTA barrier
#ifndef CC_EMULATE
    cc_tmem_wait_st();
    cc_tc_fence_before();
#endif
    __syncthreads();
#ifndef CC_EMULATE
    cc_tc_fence_after();
#endif
    if (pend0 >= 0) flush(0, pend0);
  }
#ifndef CC_EMULATE
  cc_tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == 0) cc_tmem_dealloc(t.tbase, TM::TMEM_ALLOC);
#endif
}
