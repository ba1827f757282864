// cc_mlp128_kernels.cuh -- the deep-MLP tcgen05 forward family (cc_mlp_kernels.cuh) for widths 65..128
// (BASELINE.json configs[2]: "MLP width 32..128"; the (128,128,2) architecture no other family runs).
//
// Same algorithm and TMEM map as cc_mlp_fwd_kernel<PAD> with PAD = 128 (D [0,128), A_hi [128,264), A_lo [264,400): the
// whole 512-column tensor memory, one CTA per SM); what changes is how the operands get there:
//
//   * WEIGHTS ARE STREAMED.  One layer's 3xTF32 operand (128 x 136 hi + lo) is 139 KB: three layers do not fit in shared
//     memory.  A preparation kernel (cc_mlp128_prep_kernel) writes every layer as two N-halves ("chunks": 64 outputs x 136
//     contraction columns, hi then lo, already in the MMA's K-major core-matrix layout, 69,632 B) into a global pack that
//     stays L2 resident; a 3-deep ring of chunk buffers in shared memory is refilled by cp.async.bulk (TMA 1D copies
//     completing on `full` mbarriers), freed by tcgen05.commit onto `empty` mbarriers.  The chunk sequence is periodic
//     (2 L chunks per stage evaluation), so the producer runs two chunks ahead of the MMAs.
//   * 12 warps (warps 9..11 only fill the issuer's warpgroup; setmaxnreg 56 / 224 moves its registers to the workers): warps 0..7 are workers, TWO THREADS PER TRAJECTORY (warp w and w+4 share TMEM lanes 32(w%4)..; thread half
//     hh = w/4 owns columns [64 hh, 64 hh + 64) of the state, the stage sum and every epilogue), warp 8 is the producer +
//     MMA issuer (its lane 0 waits for the 256 worker arrivals, issues the two N = 64 halves of the layer and their
//     commits, and refills the ring).  Each half has its own d_ready barrier: the epilogue of half 0 overlaps the MMAs of
//     half 1.
//   * readout: each half computes its partial dot products; halves meet through shared memory and a named barrier of the
//     256 workers once per step.
#pragma once
#include "cc_mlp_kernels.cuh"

#define CC_MLP128_WORKERS 256
#define CC_MLP128_THREADS 384  // 3 warpgroups: 2 of workers + the producer/issuer warp's (setmaxnreg is warpgroup-wide)
#define CC_MLP128_KP 136
#define CC_MLP128_CHUNK_FLOATS (2 * 64 * CC_MLP128_KP)  // hi then lo
#define CC_MLP128_CHUNK_BYTES (CC_MLP128_CHUNK_FLOATS * 4)
#define CC_MLP128_RING 3
// shared memory: ring | gw[4][132] | ypart[2][4][128] | barriers
#define CC_MLP128_SMEM_FLOATS (CC_MLP128_RING * CC_MLP128_CHUNK_FLOATS + CC_TINY_O * 132 + 2 * CC_TINY_O * 128 + 32)

// chunk c = 2 l + h of the pack: rows n = 64 h .. 64 h + 63 of layer l; element (n, k) at ((k/4)*64 + n%64)*4 + k%4
__global__ void cc_mlp128_prep_kernel(const CC_GRID_CONSTANT CcMlpArgs a) {
  const int c = blockIdx.x, l = c >> 1, hf = c & 1;
  const CcMlpLayerArg& L = a.f[l];
  const int kin = l == 0 ? a.S : L.in_log;
  float* hi_out = a.pack + (size_t)c * CC_MLP128_CHUNK_FLOATS;
  float* lo_out = hi_out + 64 * CC_MLP128_KP;
  for (int i = threadIdx.x; i < 64 * CC_MLP128_KP; i += blockDim.x) {
    const int nl = i / CC_MLP128_KP, k = i % CC_MLP128_KP, n = 64 * hf + nl;
    float w = 0.f;
    if (n < L.N) {
      if (k < kin) w = __ldg(a.theta + L.w_off + n * L.in_log + k);
      else if (l == 0 && k >= 128 + 3 && k < 128 + 3 + a.I) w = __ldg(a.theta + L.w_off + n * L.in_log + a.S + (k - 128 - 3));
      else if (k == 128 + 7 && L.b_off >= 0) w = __ldg(a.theta + L.b_off + n);
    }
    const float hi = cc_mlp_tf32_rn(w);
    const int off = ((k >> 2) * 64 + nl) * 4 + (k & 3);
    hi_out[off] = hi;
    lo_out[off] = cc_mlp_tf32_rn(w - hi);
  }
}

struct CcMlp128Ctx {
  CcMlpTm<128> t;
  int hh;  // column half of this worker thread
#ifndef CC_EMULATE
  uint64_t* d_ready2[2];
  uint32_t dpar, opar;
#endif
};

#ifdef CC_EMULATE
CC_DEV void cc_mlp128_worker_sync() { __syncthreads(); }
// the MMA batch of one layer (both halves), from the global pack
CC_DEV void cc_mlp128_round(CcMlp128Ctx& c, const float* pack, int layer) {
  using TM = CcMlpTm<128>;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int l = 0; l < 128; ++l)
      for (int n = 0; n < 128; ++n) {
        const float* bhi = pack + (size_t)(2 * layer + n / 64) * CC_MLP128_CHUNK_FLOATS;
        const float* blo = bhi + 64 * CC_MLP128_KP;
        float acc = 0.f;
        for (int pass = 0; pass < 3; ++pass)
          for (int k = 0; k < CC_MLP128_KP; ++k) {
            const int off = ((k >> 2) * 64 + (n & 63)) * 4 + (k & 3);
            const float ahi = cc_mlp_tf32(c.t.tm[(TM::COL_AHI + k) * 128 + l]), alo = cc_mlp_tf32(c.t.tm[(TM::COL_ALO + k) * 128 + l]);
            const float bh = cc_mlp_tf32(bhi[off]), bl = cc_mlp_tf32(blo[off]);
            acc += pass == 0 ? alo * bh : (pass == 1 ? ahi * bl : ahi * bh);
          }
        c.t.tm[(TM::COL_D + n) * 128 + l] = acc;
      }
  }
  __syncthreads();
}
CC_DEV void cc_mlp128_await(CcMlp128Ctx&) {}
CC_DEV void cc_mlp128_await_other(CcMlp128Ctx&) {}
#else
CC_DEV void cc_mlp128_worker_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
CC_DEV void cc_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cc_smem_u32(b)), "r"(bytes) : "memory");
}
CC_DEV void cc_bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes),
               "r"(cc_smem_u32(b))
               : "memory");
}
// one N = 64 half of a layer: D[:, 64 hf ..] = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi over the 17 contraction chunks
CC_DEV void cc_mlp128_issue_half(uint32_t tbase, int hf, uint32_t bhi) {
  using TM = CcMlpTm<128>;
  constexpr int NA = CC_MLP128_KP / 8;
  constexpr uint32_t kstep = 2u * 64u * 16u, lbo = 64u * 16u;
  constexpr uint32_t idesc = cc_umma_idesc_tf32(128, 64);
  const uint32_t blo = bhi + 64u * CC_MLP128_KP * 4u;
  const uint32_t d = tbase + 64u * (uint32_t)hf, aHi = tbase + TM::COL_AHI, aLo = tbase + TM::COL_ALO;
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(d, aLo + 8 * js, cc_umma_desc_kmajor(bhi + js * kstep, lbo, 128), idesc, js ? 1u : 0u);
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(d, aHi + 8 * js, cc_umma_desc_kmajor(blo + js * kstep, lbo, 128), idesc, 1u);
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(d, aHi + 8 * js, cc_umma_desc_kmajor(bhi + js * kstep, lbo, 128), idesc, 1u);
}
// worker side of a round: publish the A columns just written (the issuer warp picks the layer itself)
CC_DEV void cc_mlp128_round(CcMlp128Ctx& c, const float*, int) {
  cc_tmem_wait_st();
  cc_tc_fence_before();
  cc_mbar_arrive(c.t.a_ready);
}
CC_DEV void cc_mlp128_await(CcMlp128Ctx& c) {
  cc_mbar_wait(c.d_ready2[c.hh], c.dpar);
  c.dpar ^= 1;
  cc_tc_fence_after();
}
// half 0's epilogue starts while half 1's MMAs still READ the A operand: before a half-0 thread overwrites its A columns
// with the next layer's input it waits for half 1's commit too (every round, so that its parity bit stays in step)
CC_DEV void cc_mlp128_await_other(CcMlp128Ctx& c) {
  if (c.hh == 0) {
    cc_mbar_wait(c.d_ready2[1], c.opar);
    c.opar ^= 1;
  }
}
#endif

// epilogues on this thread's 64 columns [cb, cb + 64), in groups of 16 (x[64] + acc[64] leave ~90 registers for temporaries)
CC_DEV void cc_mlp128_hidden(CcMlp128Ctx& c, int act) {
  const int cb = 64 * c.hh;
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float r[16];
    cc_mlp_ld<128, 16>(c.t, CcMlpTm<128>::COL_D + cb + c0, r);
    cc_mlp_wait_ld();
    cc_mlp_act_group<16>(act, r);
    if (c0 == 0) cc_mlp128_await_other(c);
    cc_mlp_store_split<128, 16>(c.t, cb + c0, r);
  }
}
template <bool FINAL>
CC_DEV void cc_mlp128_last(CcMlp128Ctx& c, int act, float (&x)[64], float (&acc)[64], float h, float wk, float cz, int integrator) {
  const int cb = 64 * c.hh;
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float r[16];
    cc_mlp_ld<128, 16>(c.t, CcMlpTm<128>::COL_D + cb + c0, r);
    cc_mlp_wait_ld();
    cc_mlp_act_group<16>(act, r);
    if (c0 == 0) cc_mlp128_await_other(c);
    if (!FINAL) {
      float z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        acc[c0 + i] = fmaf(wk, r[i], acc[c0 + i]);
        z[i] = fmaf(h * r[i], cz, x[c0 + i]);
      }
      cc_mlp_store_split<128, 16>(c.t, cb + c0, z);
    } else if (integrator == CC_INT_RK4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        x[c0 + i] = fmaf(h, (1.f / 6.f) * (acc[c0 + i] + r[i]), x[c0 + i]);
        acc[c0 + i] = 0.f;
      }
    } else if (integrator == CC_INT_EE) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[c0 + i] = fmaf(h, r[i], x[c0 + i]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[c0 + i] = r[i];
    }
  }
}
// partial readout over this half's columns: gw[o][132] (zero beyond S, bias at index 128: added by half 0 only)
CC_DEV void cc_mlp128_readout_part(const float* gw, const float (&x)[64], int hh, int O, float (&y)[CC_TINY_O]) {
#pragma unroll
  for (int o = 0; o < CC_TINY_O; ++o) {
    y[o] = 0.f;
    if (o < O) {
      const float* row = gw + o * 132 + 64 * hh;
      float s0 = hh == 0 ? gw[o * 132 + 128] : 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(row + i);
        s0 = fmaf(wv.x, x[i], s0); s1 = fmaf(wv.y, x[i + 1], s1); s2 = fmaf(wv.z, x[i + 2], s2); s3 = fmaf(wv.w, x[i + 3], s3);
      }
      y[o] = (s0 + s1) + (s2 + s3);
    }
  }
}

__global__ void __launch_bounds__(CC_MLP128_THREADS, 1) cc_mlp128_fwd_kernel(const CC_GRID_CONSTANT CcMlpArgs a) {
  using TM = CcMlpTm<128>;
  float* smem = CC_SMEM_PTR;
  const int tid = threadIdx.x;
  float* ring = smem;
  float* gw = smem + CC_MLP128_RING * CC_MLP128_CHUNK_FLOATS;
  float* ypart = gw + CC_TINY_O * 132;
  for (int i = tid; i < CC_TINY_O * 132; i += blockDim.x) {
    const int o = i / 132, s = i % 132;
    float w = 0.f;
    if (o < a.O) {
      if (s < a.S) w = __ldg(a.theta + a.g.w_off + o * a.g.in_log + s);
      else if (s == 128 && a.g.b_off >= 0) w = __ldg(a.theta + a.g.b_off + o);
    }
    gw[i] = w;
  }
  const int nst = a.integrator == CC_INT_RK4 ? 4 : 1;
  CcMlp128Ctx c;
#ifdef CC_EMULATE
  (void)ring;
  c.t.tm = ypart + 2 * CC_TINY_O * 128 + 32;
  c.t.b = a.pack;
  c.t.tid = tid & 127;
  c.hh = tid >> 7;
  __syncthreads();
  const bool worker = true;
#else
  const int warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ypart + 2 * CC_TINY_O * 128);
  uint64_t* a_ready = bars;         // 256 worker arrivals per round
  uint64_t* d_ready = bars + 1;     // [2] one per N half (tcgen05.commit)
  uint64_t* full = bars + 3;        // [3] chunk landed (TMA complete_tx)
  uint64_t* empty = bars + 6;       // [3] chunk consumed (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  if (warp == 0) {
    if (lane == 0) {
      cc_mbar_init(a_ready, CC_MLP128_WORKERS);
      cc_mbar_init(d_ready, 1);
      cc_mbar_init(d_ready + 1, 1);
      for (int i = 0; i < CC_MLP128_RING; ++i) { cc_mbar_init(full + i, 1); cc_mbar_init(empty + i, 1); }
      cc_mbar_fence_init();
    }
    __syncwarp();
    cc_tmem_alloc(tmem_slot, 512);
  }
  cc_tc_fence_before();
  __syncthreads();
  cc_tc_fence_after();
  c.t.tbase = *tmem_slot;
  c.t.lane_base = c.t.tbase + ((uint32_t)((warp & 3) * 32) << 16);
  c.t.a_ready = a_ready;
  c.hh = (warp >> 2) & 1;
  c.d_ready2[0] = d_ready; c.d_ready2[1] = d_ready + 1;
  c.dpar = 0; c.opar = 0;
  const bool worker = warp < 8;
  if (!worker) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // ---- producer + MMA issuer (one thread) ----
    if (warp == 8 && lane == 0) {
      const int nchunk = 2 * a.L;
      const long long total = (long long)a.T * nst * nchunk;
      const uint32_t ring_u32 = cc_smem_u32(ring);
      uint32_t fpar = 0, epar = 0, apar = 0;  // bit i = parity of full[i] / empty[i]
      for (int i = 0; i < CC_MLP128_RING && i < total; ++i) {
        cc_mbar_expect_tx(full + i, CC_MLP128_CHUNK_BYTES);
        cc_bulk_g2s(ring_u32 + i * CC_MLP128_CHUNK_BYTES, a.pack + (size_t)(i % nchunk) * CC_MLP128_CHUNK_FLOATS, CC_MLP128_CHUNK_BYTES, full + i);
      }
      long long cidx = 0;
      int buf = 0, nxt = CC_MLP128_RING % nchunk;  // nxt = pack chunk of sequence position cidx + 2 (advanced below)
      const long long rounds = (long long)a.T * nst * a.L;
      for (long long r = 0; r < rounds; ++r) {
        cc_mbar_wait(a_ready, apar);
        apar ^= 1;
        cc_tc_fence_after();
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          cc_mbar_wait(full + buf, (fpar >> buf) & 1u);
          fpar ^= 1u << buf;
          cc_mlp128_issue_half(c.t.tbase, hf, ring_u32 + buf * CC_MLP128_CHUNK_BYTES);
          cc_tc_commit(d_ready + hf);
          cc_tc_commit(empty + buf);
          if (cidx >= 1 && cidx + 2 < total) {  // refill the buffer chunk cidx-1 used with chunk cidx+2
            const int pb = buf == 0 ? CC_MLP128_RING - 1 : buf - 1;
            cc_mbar_wait(empty + pb, (epar >> pb) & 1u);
            epar ^= 1u << pb;
            cc_mbar_expect_tx(full + pb, CC_MLP128_CHUNK_BYTES);
            cc_bulk_g2s(ring_u32 + pb * CC_MLP128_CHUNK_BYTES, a.pack + (size_t)nxt * CC_MLP128_CHUNK_FLOATS, CC_MLP128_CHUNK_BYTES, full + pb);
            nxt = nxt + 1 == nchunk ? 0 : nxt + 1;
          }
          ++cidx;
          buf = buf + 1 == CC_MLP128_RING ? 0 : buf + 1;
        }
      }
    }
    __syncwarp();
  }
#endif
  if (worker) {
#ifndef CC_EMULATE
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
#endif
    const int row = (tid & 127);
    const long long b = (long long)blockIdx.x * 128 + row;
    const bool act = b < a.B;
    const int hh = c.hh, cb = 64 * hh;
    const float h = a.dt;
    float x[64], acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      x[i] = (a.x0 && act && cb + i < a.S) ? __ldg(a.x0 + b * a.S + cb + i) : 0.f;
      acc[i] = 0.f;
    }
    float yp[CC_TINY_O];
    // y = out_scale * act(g . x + b) from the two halves' partial sums (half 1 -> shared memory -> half 0 writes the output)
    // (ypart is double buffered by krow parity: the NEXT emit's barrier orders half 0's read against half 1's rewrite)
    auto emit = [&](long long krow) {
      float* yb = ypart + (krow & 1) * (CC_TINY_O * 128);
      cc_mlp128_readout_part(gw, x, hh, a.O, yp);
      if (hh == 1)
        for (int o = 0; o < a.O; ++o) yb[o * 128 + row] = yp[o];
      cc_mlp128_worker_sync();
      if (hh == 0 && act)
        for (int o = 0; o < a.O; ++o)
          a.y[b * (long long)(a.T + 1) * a.O + krow * a.O + o] = a.out_scale * cc_act(a.g.act, yp[o] + yb[o * 128 + row]);
    };
    emit(0);  // y0 = g(x0)      neural_ode_model.py:160-175
    float vin[CC_TINY_O];
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) vin[i] = (hh == 0 && i < a.I && act && a.T > 0) ? __ldg(a.u + b * (long long)a.T * a.I + i) : 0.f;
    for (int k = 0; k < a.T; ++k) {
      if (a.tape && blockIdx.x == 0 && tid == 0) a.ttab[k] = 0.f;
      if (a.tape && act && (k % a.ckpt_every == 0)) {
        const long long ck = k / a.ckpt_every;
#pragma unroll
        for (int s = 0; s < 64; ++s)
          if (cb + s < a.S) a.tape[(ck * a.S + cb + s) * a.Bpad + b] = x[s];
      }
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        float z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = x[c0 + i];
        cc_mlp_store_split<128, 16>(c.t, cb + c0, z);
      }
      if (hh == 0) {
        float vt[CC_TINY_O];
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i) vt[i] = (i < a.I) ? cc_ut(a.ut, a.u_scale, vin[i]) : 0.f;
        if (k + 1 < a.T) {
#pragma unroll
          for (int i = 0; i < CC_TINY_O; ++i)
            if (i < a.I && act) vin[i] = __ldg(a.u + b * (long long)a.T * a.I + (long long)(k + 1) * a.I + i);
        }
        float tail[8];
        tail[0] = 0.f; tail[1] = 0.f; tail[2] = 0.f; tail[3] = vt[0]; tail[4] = vt[1]; tail[5] = vt[2]; tail[6] = vt[3]; tail[7] = 1.f;
        cc_mlp_store_split<128, 8>(c.t, 128, tail);
      }
      cc_mlp128_round(c, a.pack, 0);
      if (a.pre) emit(k + 1);  // compact model: y = g(x) BEFORE the update; overlaps the MMAs
#pragma unroll 1
      for (int st = 0; st < nst; ++st) {
#pragma unroll 1
        for (int l = 0; l < a.L; ++l) {
          cc_mlp128_await(c);
          if (l < a.L - 1) {
            cc_mlp128_hidden(c, a.f[l].act);
            cc_mlp128_round(c, a.pack, l + 1);
          } else if (st < nst - 1) {
            cc_mlp128_last<false>(c, a.f[l].act, x, acc, h, st == 0 ? 1.f : 2.f, st == 2 ? 1.f : 0.5f, a.integrator);
            cc_mlp128_round(c, a.pack, 0);
          } else {
            cc_mlp128_last<true>(c, a.f[l].act, x, acc, h, 0.f, 0.f, a.integrator);
          }
        }
      }
      if (!a.pre) emit(k + 1);  // full model: y = g(x_next)
    }
    if (a.x_final && act)
#pragma unroll
      for (int s = 0; s < 64; ++s)
        if (cb + s < a.S) a.x_final[b * a.S + cb + s] = x[s];
    if (a.tape_xT && act)
#pragma unroll
      for (int s = 0; s < 64; ++s)
        if (cb + s < a.S) a.tape_xT[(size_t)(cb + s) * a.Bpad + b] = x[s];
  }
#ifndef CC_EMULATE
  cc_tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == 0) cc_tmem_dealloc(c.t.tbase, 512);
#endif
}
