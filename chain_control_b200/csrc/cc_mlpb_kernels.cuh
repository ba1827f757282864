// cc_mlpb_kernels.cuh -- deep-MLP REVERSE pass on tcgen05 (kernel family 6; sm_100a): BPTT of the open-loop unroll with a
// trainable deep-MLP model (ModelTrainer, cc/train/step_fn.py:133-146; model neural_ode_model.py:27-36 / 106-175, MLP
// mlp_network.py:5-35, integrator integrate.py:1-28) for widths 17..64 -- the architectures of BASELINE.json configs[2]
// whose reverse pass the FFMA family serves with one or two warps per SM (width <= 48) or not at all (width 64).
//
// What eqx.filter_value_and_grad derives for loss_fn_model: per step, last first, the RK4 adjoint
//   kbar_s = a_s lam + b_s zbar_{s+1},  delta_{L-1} = kbar_s . act_f'(k_s),  delta_{l-1} = (delta_l W_l) . act'(a_l),
//   zbar_s = delta_0 W_0[:, :S],  lam <- lam + sum_s zbar_s + G^T d,   Wbar_l += delta_l^T [a_l ; u~ ; 1],  Gbar += d^T [x ; 1].
//
// Execution model (same tile / thread mapping as cc_mlp128_kernels.cuh):
//   * one CTA = 128 trajectories; 8 worker warps, TWO THREADS PER TRAJECTORY (thread half hh owns columns [32 hh, 32 hh + 32)
//     of every vector: lam, the stage sum, the hidden activations kept for act'), warp 8 = producer + MMA issuer.
//   * EVERY matrix product is a tcgen05.mma batch, 3xTF32 for the chains:
//       forward recompute  D[128 x 64] = A[128 x 72] . F_l^T      (A = layer input hi/lo in TMEM, tail chunk [0 0 0 u~ 1])
//       input gradient     D[128 x 64] = A[128 x 64] . T_l^T      (A = delta_l hi/lo in TMEM, T_l = W_l transposed)
//       weight gradient    DW_l[128 x 80] += opA^T . opB          (SS mode, K = the 128 TRAJECTORIES: opA rows = delta_l hi (0..63)
//                                                                   and lo (64..127), opB rows = [a_l ; tail] rounded to tf32;
//                                                                   both written K-major by the thread that owns the trajectory)
//       readout gradient   DG[128 x 64]   = opA2^T . opB          (rows 0..3 = d hi, 4..7 = d lo; rows 8.. alias opA, ignored)
//     MN-major (transposed) tf32 operands return zeros on this hardware (scratch/tc_probe6.cu / tc_probe7.cu), so W_l and
//     W_l^T are separate operand images and do not fit in shared memory next to the weight-gradient operands: ALL weight
//     operands are STREAMED -- a preparation kernel writes the 2 L slabs (hi + lo, MMA core-matrix layout, 36,864 B each) into
//     a global pack that stays L2 resident, and a 2-deep shared-memory ring is refilled by cp.async.bulk one round ahead.
//   * per step (RK4, L layers): a first pass recomputes z_2..z_4 (3 L rounds; they wait in a per-CTA global scratch, written and
//     read by the same thread, L2 resident), then for s = 4..1 the hidden activations of stage s are recomputed (L-1 rounds)
//     and the reverse chain runs (L rounds): 29 rounds for L = 3.  One mbarrier pair per round (256 arrivals -> issue ->
//     tcgen05.commit); the round sequence is a table in shared memory that the issuer and the workers walk in lockstep.
//   * the tensor core accumulates with truncation (scratch/tc_probe5.cu): the DW accumulators are flushed into per-CTA global
//     sums every `flush_every` steps; a reduction kernel adds the CTAs' sums in fixed order (deterministic).
//   * the tape is the forward family's ([k][S][Bpad] pre-step states) plus the final state x_T (post-step readout only).
//
// Limits (checked by the host, cc_abi.cu: mlpb_ok): f with 2 or 3 Linear layers of width <= 64, g one Linear layer without
// activation, time-invariant, no feed-through, 1..4 inputs, <= 4 outputs, no ubar.
// ckpt_every > 1 (cc_abi_bwd.inc): the host replays each segment [k0, k1) with the forward family into a segment-local tape and
// launches this kernel per segment, last first: the adjoint enters / leaves through lam_io, the per-CTA sums accumulate
// (seg_accum), the terminal readout term belongs to the last segment only (seg_last).
#pragma once
#include "cc_mlp128_kernels.cuh"

#define CC_MLPB_SLAB_FLOATS 9216  // hi [0, 4608) then lo [4608, 9216); F_l uses 72 x 64, T_l 64 x 64 of each half
#define CC_MLPB_SLAB_BYTES (CC_MLPB_SLAB_FLOATS * 4)
#define CC_MLPB_LO 4608
#define CC_MLPB_COL_DW 208  // + 80 l
#define CC_MLPB_COL_DG 448
#define CC_MLPB_OPG 1152    // floats of one 8-row group of a K = trajectory operand (32 core matrices of 128 B padded to 144 B)
#define CC_MLPB_REC_FLOATS(L) ((L) * 80 * 128 + 8 * 64 + 16)  // per-CTA sums: DW [L][80][128], DG [8][64], g's bias [4]
#define CC_MLPB_CTA_FLOATS(L) (CC_MLPB_REC_FLOATS(L) + 4 * 64 * 128)  // + the z_2..z_4 slots and the step's incoming adjoint
#define CC_MLPB_DGONLY 0x10
// shared memory (floats): ring[2] | opA2 | opA | opB | gw[4][68] | gacc[8][64] | gbw[4][4] | tab[64] | barriers
#define CC_MLPB_SMEM_FLOATS (2 * CC_MLPB_SLAB_FLOATS + 27 * CC_MLPB_OPG + 4 * 68 + 8 * 64 + 16 + 64 + 16)

CC_DEV constexpr int cc_mlpb_row(int r) { return (r >> 3) * CC_MLPB_OPG + (r & 7) * 4; }

// slab c < L: F_c (forward operand of layer c: element (n, k) = W[n][k] / input weights at k = 67 + i / bias at k = 71);
// slab L + l: T_l (element (j, n) = W_l[n][j], j < in-state width): both at ((k/4)*64 + row)*4 + k%4, hi then lo
__global__ void cc_mlpb_prep_kernel(const CC_GRID_CONSTANT CcMlpArgs a) {
  const int c = blockIdx.x;
  const bool tr = c >= a.L;
  const int l = tr ? c - a.L : c;
  const CcMlpLayerArg& Ly = a.f[l];
  const int kin = l == 0 ? a.S : Ly.in_log;
  float* hi_out = a.pack + (size_t)c * CC_MLPB_SLAB_FLOATS;
  float* lo_out = hi_out + CC_MLPB_LO;
  for (int i = threadIdx.x; i < 64 * 72; i += blockDim.x) {
    const int r = i / 72, k = i % 72;
    float w = 0.f;
    if (!tr) {
      if (r < Ly.N) {
        if (k < kin) w = __ldg(a.theta + Ly.w_off + r * Ly.in_log + k);
        else if (l == 0 && k >= 67 && k < 67 + a.I) w = __ldg(a.theta + Ly.w_off + r * Ly.in_log + a.S + (k - 67));
        else if (k == 71 && Ly.b_off >= 0) w = __ldg(a.theta + Ly.b_off + r);
      }
    } else if (r < kin && k < Ly.N) {
      w = __ldg(a.theta + Ly.w_off + k * Ly.in_log + r);
    }
    const float hi = cc_mlp_tf32_rn(w);
    const int off = ((k >> 2) * 64 + r) * 4 + (k & 3);
    hi_out[off] = hi;
    lo_out[off] = cc_mlp_tf32_rn(w - hi);
  }
}

// round table of one step: bits 0-1 layer, bit 2 reverse round (T_l + weight gradient against the hi part of the layer input),
// bit 6 weight-gradient round against the lo part (no slab), bit 3 + the readout gradient, bit 5 first weight-gradient
// round of this layer in the step (overwrites the accumulator after a flush)
CC_DEV int cc_mlpb_table(int* tab, int L, int rk4, int fa) {
  int n = 0;
  if (rk4)
    for (int st = 0; st < 3; ++st)
      for (int l = 0; l < L; ++l) tab[n++] = l;
  const int nst = rk4 ? 4 : 1;
  for (int s = nst - 1; s >= 0; --s) {
    for (int l = 0; l < (fa ? L : L - 1); ++l) tab[n++] = l;
    for (int l = L - 1; l >= 0; --l) {
      tab[n++] = l | 4 | ((l == 0 && s == 0) ? 8 : 0) | (s == nst - 1 ? 32 : 0);
      tab[n++] = l | 64 | ((l == 0 && s == 0) ? 8 : 0);
    }
  }
  return n;
}

struct CcMlpbCtx {
  CcMlpTm<64> t;  // TMEM access through cc_mlp_ld / cc_mlp_st / cc_mlp_store_split (D = 0, A_hi = 64, A_lo = 136)
  int hh, row;
  float* opa2;  // this thread's trajectory column of the three K = trajectory operands
  float* opa;
  float* opb;
  const int* tab;
  int ri, R;
#ifdef CC_EMULATE
  const float* pack;
  float* opa2_base;
  float* opb_base;
  int L, fresh;
#else
  uint32_t dpar;
#endif
};

#ifdef CC_EMULATE
CC_DEV void cc_mlpb_emu_mma(CcMlpbCtx& c, int code) {
  float* tm = c.t.tm;
  if (code & 64) {  // lo part of the layer input: rows 0..63 of opB only (the tail rows keep their hi-pass content)
    if (!(code & CC_MLPB_DGONLY)) {
      const int l = code & 3;
      const float* opa = c.opa2_base + CC_MLPB_OPG;
      for (int r = 0; r < 128; ++r)
        for (int j = 0; j < 64; ++j) {
          float s = 0.f;
          for (int tr = 0; tr < 128; ++tr)
            s += cc_mlp_tf32(opa[cc_mlpb_row(r) + (tr >> 2) * 36 + (tr & 3)]) * cc_mlp_tf32(c.opb_base[cc_mlpb_row(j) + (tr >> 2) * 36 + (tr & 3)]);
          tm[(CC_MLPB_COL_DW + 80 * l + j) * 128 + r] += s;
        }
    }
    if (code & (8 | CC_MLPB_DGONLY))
      for (int r = 0; r < 128; ++r)
        for (int j = 0; j < 64; ++j) {
          float s = 0.f;
          for (int tr = 0; tr < 128; ++tr)
            s += cc_mlp_tf32(c.opa2_base[cc_mlpb_row(r) + (tr >> 2) * 36 + (tr & 3)]) * cc_mlp_tf32(c.opb_base[cc_mlpb_row(j) + (tr >> 2) * 36 + (tr & 3)]);
          tm[(CC_MLPB_COL_DG + j) * 128 + r] += s;
        }
    return;
  }
  if (code != CC_MLPB_DGONLY) {
    const int l = code & 3;
    const bool bwd = (code & 4) != 0;
    const float* slab = c.pack + (size_t)((bwd ? c.L : 0) + l) * CC_MLPB_SLAB_FLOATS;
    const int K = bwd ? 64 : 72;
    for (int ln = 0; ln < 128; ++ln)
      for (int n = 0; n < 64; ++n) {
        float acc = 0.f;
        for (int pass = 0; pass < 3; ++pass)
          for (int k = 0; k < K; ++k) {
            const int off = ((k >> 2) * 64 + n) * 4 + (k & 3);
            const float ahi = cc_mlp_tf32(tm[(64 + k) * 128 + ln]), alo = cc_mlp_tf32(tm[(136 + k) * 128 + ln]);
            const float bh = cc_mlp_tf32(slab[off]), bl = cc_mlp_tf32(slab[CC_MLPB_LO + off]);
            acc += pass == 0 ? alo * bh : (pass == 1 ? ahi * bl : ahi * bh);
          }
        tm[n * 128 + ln] = acc;
      }
    if (bwd) {
      const bool over = (code & 32) && c.fresh;
      const float* opa = c.opa2_base + CC_MLPB_OPG;
      for (int r = 0; r < 128; ++r)
        for (int j = 0; j < 80; ++j) {
          float s = 0.f;
          for (int tr = 0; tr < 128; ++tr)
            s += cc_mlp_tf32(opa[cc_mlpb_row(r) + (tr >> 2) * 36 + (tr & 3)]) * cc_mlp_tf32(c.opb_base[cc_mlpb_row(j) + (tr >> 2) * 36 + (tr & 3)]);
          float& d = tm[(CC_MLPB_COL_DW + 80 * l + j) * 128 + r];
          d = over ? s : d + s;
        }
    }
  }
  if (code == CC_MLPB_DGONLY || (code & 8)) {
    for (int r = 0; r < 128; ++r)
      for (int j = 0; j < 64; ++j) {
        float s = 0.f;
        for (int tr = 0; tr < 128; ++tr)
          s += cc_mlp_tf32(c.opa2_base[cc_mlpb_row(r) + (tr >> 2) * 36 + (tr & 3)]) * cc_mlp_tf32(c.opb_base[cc_mlpb_row(j) + (tr >> 2) * 36 + (tr & 3)]);
        tm[(CC_MLPB_COL_DG + j) * 128 + r] = s;
      }
  }
}
// publish this thread's operand writes and run the round `expect` (the table must agree: the issuer of the GPU build walks it alone)
CC_DEV void cc_mlpb_round(CcMlpbCtx& c, int expect) {
  int code = expect;
  if (!(expect & CC_MLPB_DGONLY)) {
    code = c.tab[c.ri];
    c.ri = c.ri + 1 == c.R ? 0 : c.ri + 1;
    if ((code & 79) != (expect & 79)) {
      fprintf(stderr, "ccemu: cc_mlpb round table out of step (table %d, worker %d)\n", code, expect);
      abort();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cc_mlpb_emu_mma(c, code);
  __syncthreads();
}
CC_DEV void cc_mlpb_await(CcMlpbCtx&) {}
#else
CC_DEV void cc_mlpb_round(CcMlpbCtx& c, int) {
  cc_tmem_wait_st();
  cc_fence_proxy_async();  // the operand rows written with plain stores become visible to the tensor core's (async-proxy) reads
  cc_tc_fence_before();
  cc_mbar_arrive(c.t.a_ready);
}
CC_DEV void cc_mlpb_await(CcMlpbCtx& c) {
  cc_mbar_wait(c.t.d_ready, c.dpar);
  c.dpar ^= 1;
  cc_tc_fence_after();
}
// Descriptors are derived from pre-encoded bases by one add each (the start-address field holds address >> 4 and never
// overflows); the bases are made opaque per batch so that the compiler does not hoist ~100 loop-invariant descriptors out of the
// round loop into the issuing thread's 56 registers (ncu on the first version: the issuer spent the whole step on their spill loads)
CC_DEV uint64_t cc_mlpb_desc(uint64_t base, uint32_t byte_off) { return base + (uint64_t)(byte_off >> 4); }
CC_DEV void cc_mlpb_opaque(uint64_t& d) { asm volatile("" : "+l"(d)); }
// forward / input-gradient batch: D = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi over NA contraction chunks of one slab
template <int NA>
CC_DEV void cc_mlpb_issue_ts(uint32_t tbase, uint32_t bhi_addr) {
  constexpr uint32_t kstep = 2u * 64u * 16u, lbo = 64u * 16u;
  constexpr uint32_t idesc = cc_umma_idesc_tf32(128, 64);
  uint64_t bhi = cc_umma_desc_kmajor(bhi_addr, lbo, 128);
  cc_mlpb_opaque(bhi);
  const uint64_t blo = cc_mlpb_desc(bhi, CC_MLPB_LO * 4u);
  const uint32_t aHi = tbase + 64, aLo = tbase + 136;
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(tbase, aLo + 8 * js, cc_mlpb_desc(bhi, js * kstep), idesc, js ? 1u : 0u);
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(tbase, aHi + 8 * js, cc_mlpb_desc(blo, js * kstep), idesc, 1u);
#pragma unroll
  for (int js = 0; js < NA; ++js) cc_umma_tf32_ts(tbase, aHi + 8 * js, cc_mlpb_desc(bhi, js * kstep), idesc, 1u);
}
// K = trajectory product: 16 instructions of 8 trajectories (two padded core matrices, LBO 144 B; 8-row groups SBO 4608 B)
template <int N>
CC_DEV void cc_mlpb_issue_ss(uint32_t d, uint64_t opa, uint64_t opb, uint32_t acc0) {
  constexpr uint32_t idesc = cc_umma_idesc_tf32(128, N);
  cc_mlpb_opaque(opa);
  cc_mlpb_opaque(opb);
#pragma unroll
  for (int js = 0; js < 16; ++js) cc_umma_tf32_ss(d, cc_mlpb_desc(opa, js * 288u), cc_mlpb_desc(opb, js * 288u), idesc, js ? 1u : acc0);
}
#endif

template <int NWORK>
CC_DEV void cc_mlpb_worker_sync() {
#ifdef CC_EMULATE
  __syncthreads();
#else
  asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory");
#endif
}
// NC values -> rows c0.. of a K = trajectory operand, rounded to nearest tf32
template <int NC>
CC_DEV void cc_mlpb_op_rn(float* op, int c0, const float (&v)[NC]) {
#pragma unroll
  for (int i = 0; i < NC; ++i) op[cc_mlpb_row(c0 + i)] = cc_mlp_tf32_rn(v[i]);
}
// ... and the exact remainders (the tensor core truncates them to tf32 itself)
template <int NC>
CC_DEV void cc_mlpb_op_lo(float* op, int c0, const float (&v)[NC]) {
#pragma unroll
  for (int i = 0; i < NC; ++i) op[cc_mlpb_row(c0 + i)] = v[i] - cc_mlp_tf32_rn(v[i]);
}
// delta: hi / lo split -> A (TMEM columns c0..) and the operand rows c0.. (hi) / 64 + c0.. (lo)
template <int NC>
CC_DEV void cc_mlpb_store_delta(const CcMlpbCtx& c, int cb, int c0, const float (&v)[NC]) {
  float hi[NC], lo[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    hi[i] = cc_mlp_tf32_rn(v[i]);
    lo[i] = v[i] - hi[i];
  }
  cc_mlp_st<64, NC>(c.t, 64 + cb + c0, hi);
  cc_mlp_st<64, NC>(c.t, 136 + cb + c0, lo);
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    c.opa[cc_mlpb_row(c0 + i)] = hi[i];
    c.opa[cc_mlpb_row(64 + c0 + i)] = lo[i];
  }
}
CC_DEV void cc_mlpb_act_grad_group(int act, float (&d)[16], const float* y) {
  if (act == CC_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = y[i] > 0.f ? d[i] : 0.f;
  } else if (act == CC_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] *= 1.f - y[i] * y[i];
  }
}

template <int L, int NH>
__global__ void __launch_bounds__(128 * NH + 128, 1) cc_mlpb_bwd_kernel(const CC_GRID_CONSTANT CcMlpArgs a) {
  constexpr int PC = 64 / NH;        // columns per worker thread
  constexpr int NWORK = 128 * NH;    // worker threads: NH per trajectory (warp w, w + 4, ... share a TMEM lane quarter)
  float* smem = CC_SMEM_PTR;
  const int tid = threadIdx.x;
  float* ring = smem;
  float* opa2 = ring + 2 * CC_MLPB_SLAB_FLOATS;
  float* opb = opa2 + 17 * CC_MLPB_OPG;
  float* gw = opb + 10 * CC_MLPB_OPG;
  float* gacc = gw + 4 * 68;
  float* gbw = gacc + 8 * 64;
  int* tab = reinterpret_cast<int*>(gbw + 16);
  const bool rk4 = a.integrator == CC_INT_RK4;
  const int nst = rk4 ? 4 : 1;
  const int fa = a.f[L - 1].act != CC_ACT_IDENTITY;
  const bool post = !a.pre;
  for (int i = tid; i < 4 * 68; i += blockDim.x) {
    const int o = i / 68, s = i % 68;
    gw[i] = (o < a.O && s < a.S) ? __ldg(a.theta + a.g.w_off + o * a.g.in_log + s) : 0.f;
  }
  for (int i = tid; i < 8 * 64; i += blockDim.x) gacc[i] = 0.f;
  if (tid == 0) tab[63] = cc_mlpb_table(tab, L, rk4, fa);
  float* rec = a.recs + (size_t)blockIdx.x * CC_MLPB_CTA_FLOATS(L);
  float* zs = rec + CC_MLPB_REC_FLOATS(L);
  CcMlpbCtx c;
#ifdef CC_EMULATE
  (void)ring;
  c.t.tm = reinterpret_cast<float*>(tab + 64) + 16;
  c.t.tid = tid & 127;
  c.pack = a.pack; c.opa2_base = opa2; c.opb_base = opb; c.L = L; c.fresh = 1;
  c.hh = tid >> 7;
  __syncthreads();
  const bool worker = true;
#else
  const int warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab + 64);
  uint64_t* a_ready = bars;      // 256 worker arrivals per round
  uint64_t* d_ready = bars + 1;  // tcgen05.commit of the round's batch
  uint64_t* full = bars + 2;     // [2] slab landed (TMA complete_tx)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  if (warp == 0) {
    if (lane == 0) {
      cc_mbar_init(a_ready, NWORK);
      cc_mbar_init(d_ready, 1);
      cc_mbar_init(full, 1);
      cc_mbar_init(full + 1, 1);
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
  c.t.a_ready = a_ready; c.t.d_ready = d_ready;
  c.hh = warp >> 2;
  c.dpar = 0;
  const bool worker = warp < 4 * NH;
  if (!worker) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // ---- producer + MMA issuer (one thread) ----
    if (warp == 4 * NH && lane == 0) {
      const int R = tab[63];
      const int kend = a.k1 > 0 ? a.k1 : a.T;
      const long long rounds = (long long)(kend - a.k0) * R;
      const uint32_t ring_u32 = cc_smem_u32(ring);
      const uint64_t opa2_u32 = cc_umma_desc_kmajor(cc_smem_u32(opa2), 144, 4608), opa_u32 = cc_mlpb_desc(opa2_u32, CC_MLPB_OPG * 4u), opb_u32 = cc_umma_desc_kmajor(cc_smem_u32(opb), 144, 4608);
      const uint32_t tb = c.t.tbase;
      auto slab_of = [&](int code) { return ((code & 4) ? L : 0) + (code & 3); };
      int nx = 0;  // table position of the next slab to request (rounds with bit 6 use none)
      auto request = [&](int buf) {
        while (tab[nx] & 64) nx = nx + 1 == R ? 0 : nx + 1;
        cc_mbar_expect_tx(full + buf, CC_MLPB_SLAB_BYTES);
        cc_bulk_g2s(ring_u32 + buf * CC_MLPB_SLAB_BYTES, a.pack + (size_t)slab_of(tab[nx]) * CC_MLPB_SLAB_FLOATS, CC_MLPB_SLAB_BYTES, full + buf);
        nx = nx + 1 == R ? 0 : nx + 1;
      };
      if (rounds > 0) { request(0); request(1); }  // (requests past the last round are never waited for: harmless reads of the pack)
      uint32_t apar = 0, dpar = 0, fpar = 0;
      if (post && a.seg_last) {  // terminal state: readout gradient only (hi then lo part of x_T)
        for (int part = 0; part < 2; ++part) {
          cc_mbar_wait(a_ready, apar); apar ^= 1;
          cc_tc_fence_after();
          cc_mlpb_issue_ss<64>(tb + CC_MLPB_COL_DG, opa2_u32, opb_u32, (uint32_t)part);
          cc_tc_commit(d_ready);
          cc_mbar_wait(d_ready, dpar); dpar ^= 1;
        }
      }
      int ri = 0, step = 0, buf = 0;
      for (long long r = 0; r < rounds; ++r) {
        const int code = tab[ri];
        cc_mbar_wait(a_ready, apar); apar ^= 1;
        cc_tc_fence_after();
        if (code & 64) {
          cc_mlpb_issue_ss<64>(tb + CC_MLPB_COL_DW + 80 * (code & 3), opa_u32, opb_u32, 1u);
          if (code & 8) cc_mlpb_issue_ss<64>(tb + CC_MLPB_COL_DG, opa2_u32, opb_u32, 1u);
          cc_tc_commit(d_ready);
          cc_mbar_wait(d_ready, dpar); dpar ^= 1;  // (every phase is waited for: a skipped one would alias the parity test)
        } else {
          cc_mbar_wait(full + buf, (fpar >> buf) & 1u);
          fpar ^= 1u << buf;
          const uint32_t sl = ring_u32 + buf * CC_MLPB_SLAB_BYTES;
          if (code & 4) {
            cc_mlpb_issue_ts<8>(tb, sl);
            const uint32_t keep = ((code & 32) && (step % a.flush_every == 0)) ? 0u : 1u;
            cc_mlpb_issue_ss<80>(tb + CC_MLPB_COL_DW + 80 * (code & 3), opa_u32, opb_u32, keep);
            if (code & 8) cc_mlpb_issue_ss<64>(tb + CC_MLPB_COL_DG, opa2_u32, opb_u32, 0u);
          } else {
            cc_mlpb_issue_ts<9>(tb, sl);
          }
          cc_tc_commit(d_ready);
          // the batch must have finished reading this buffer before the slab two ahead lands in it
          cc_mbar_wait(d_ready, dpar); dpar ^= 1;
          request(buf);
          buf ^= 1;
        }
        ri = ri + 1;
        if (ri == R) { ri = 0; ++step; }
      }
    }
    __syncwarp();
  }
#endif
  if (worker) {
#ifndef CC_EMULATE
    if (NH == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");  // (640 threads x 96 registers at launch: 128 x 56 + 512 x 104 fit the pool)
#endif
    const int row = tid & 127, hh = c.hh, cb = PC * hh;
    c.row = row;
    // (cb is a multiple of 8: row(cb + j) = row(cb) + row(j), so every operand store below has a compile-time offset)
    c.opa2 = opa2 + (row >> 2) * 36 + (row & 3);
    c.opa = c.opa2 + CC_MLPB_OPG + cc_mlpb_row(cb);
    c.opb = opb + (row >> 2) * 36 + (row & 3) + cc_mlpb_row(cb);
    float* const opb_tail = opb + (row >> 2) * 36 + (row & 3) + cc_mlpb_row(64);
    c.tab = tab; c.ri = 0; c.R = tab[63];
    const long long b = (long long)blockIdx.x * 128 + row;
    const bool act = b < a.B;
    const float h = a.dt;
    const float c6 = h / 6.f, c3 = h / 3.f, c2 = h / 2.f;
    const long long yrow = (long long)(a.T + 1) * a.O;
    // per-CTA sums start at zero; opB's unused rows 72..79 stay zero
    if (!a.seg_accum)
      for (int l = 0; l < L; ++l)
        for (int j = 0; j < 80 / NH; ++j) rec[(size_t)(l * 80 + (80 / NH) * hh + j) * 128 + row] = 0.f;
    if (hh == 0)
      for (int r = 8; r < 16; ++r) opb_tail[cc_mlpb_row(r)] = 0.f;
    float lam[PC], dl[PC], hid[L - 1][PC];  // (the stage sum accumulates in lam; the incoming adjoint waits in slot 3 of the scratch)
    float* const zs_c = zs + (size_t)cb * 128 + row;  // this thread's column block of the scratch slots
    float* lam0 = zs_c + (size_t)(3 * 64) * 128;
    float gb[CC_TINY_O];
    // (segment replay: the adjoint of the later segment comes in through lam_io)
#pragma unroll
    for (int i = 0; i < PC; ++i) lam[i] = (a.lam_io && !a.seg_last && act && cb + i < a.S) ? a.lam_io[(size_t)(cb + i) * a.Bpad + b] : 0.f;
#pragma unroll
    for (int o = 0; o < CC_TINY_O; ++o) gb[o] = 0.f;
    const int kbeg = a.k0, kend = a.k1 > 0 ? a.k1 : a.T;

    // readout cotangent rows of opA2 (hi rows 0..3, lo rows 4..7) -- written by half 0
    auto put_d = [&](const float (&d)[CC_TINY_O]) {
      if (hh == 0) {
#pragma unroll
        for (int o = 0; o < CC_TINY_O; ++o) {
          const float hi = cc_mlp_tf32_rn(d[o]);
          c.opa2[cc_mlpb_row(o)] = hi;
          c.opa2[cc_mlpb_row(4 + o)] = d[o] - hi;
        }
      }
    };
    // DG (rows 0..7 of the TMEM accumulator) -> gacc; the lane quarter 0 warps of both halves
    auto flush_dg = [&]() {
      if ((row >> 5) == 0) {
#pragma unroll
        for (int c0 = 0; c0 < PC; c0 += 16) {
          float r[16];
          cc_mlp_ld<64, 16>(c.t, CC_MLPB_COL_DG + cb + c0, r);
          cc_mlp_wait_ld();
          if (row < 8) {
#pragma unroll
            for (int i = 0; i < 16; ++i) gacc[row * 64 + cb + c0 + i] += r[i];
          }
        }
      }
    };
    auto flush_dw = [&]() {
      constexpr int CW = 80 / NH;  // accumulator columns of this thread: 40 = 16 + 16 + 8 or 20 = 16 + 4
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        float* dst = rec + (size_t)(l * 80 + CW * hh) * 128 + row;
        const int col = CC_MLPB_COL_DW + 80 * l + CW * hh;
#pragma unroll
        for (int c0 = 0; c0 + 16 <= CW; c0 += 16) {
          float r[16];
          cc_mlp_ld<64, 16>(c.t, col + c0, r);
          cc_mlp_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[(size_t)(c0 + i) * 128] += r[i];
        }
        constexpr int REM = CW % 16, R0 = CW - REM;
        float rr[REM];
        cc_mlp_ld<64, REM>(c.t, col + R0, rr);
        cc_mlp_wait_ld();
#pragma unroll
        for (int i = 0; i < REM; ++i) dst[(size_t)(R0 + i) * 128] += rr[i];
      }
    };
    // lam += G^T d over this half's columns
    auto add_readout = [&](const float (&d)[CC_TINY_O]) {
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o)
        if (o < a.O) {
          const float* grow = gw + o * 68 + cb;
#pragma unroll
          for (int i = 0; i < PC; ++i) lam[i] = fmaf(grow[i], d[o], lam[i]);
        }
    };

    if (post && a.seg_last) {  // y_T = g(x_T): the final state sits behind the step times on the tape
      float d[CC_TINY_O];
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o) d[o] = (o < a.O && act) ? a.out_scale * __ldg(a.ybar + b * yrow + (long long)a.T * a.O + o) : 0.f;
#pragma unroll
      for (int c0 = 0; c0 < PC; c0 += 16) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (act && cb + c0 + i < a.S) ? a.tape_xT[(size_t)(cb + c0 + i) * a.Bpad + b] : 0.f;
        cc_mlpb_op_rn<16>(c.opb, c0, v);
      }
      put_d(d);
      add_readout(d);
      if (hh == 0)
#pragma unroll
        for (int o = 0; o < CC_TINY_O; ++o) gb[o] += d[o];
      cc_mlpb_round(c, CC_MLPB_DGONLY);
      cc_mlpb_await(c);
#pragma unroll
      for (int c0 = 0; c0 < PC; c0 += 16) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (act && cb + c0 + i < a.S) ? a.tape_xT[(size_t)(cb + c0 + i) * a.Bpad + b] : 0.f;
        cc_mlpb_op_lo<16>(c.opb, c0, v);
      }
      cc_mlpb_round(c, CC_MLPB_DGONLY | 64);
      cc_mlpb_await(c);
      flush_dg();
    }

    int since_flush = 0;
    for (int k = kend - 1; k >= kbeg; --k) {
      const float* xk = a.tape + (size_t)(k - kbeg) * a.S * a.Bpad + b;
      float d[CC_TINY_O], dg[CC_TINY_O];
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o) {
        d[o] = (o < a.O && act) ? a.out_scale * __ldg(a.ybar + b * yrow + (long long)(a.pre ? k + 1 : k) * a.O + o) : 0.f;
        dg[o] = d[o];
        if (a.pre && k == 0 && o < a.O && act) dg[o] += a.out_scale * __ldg(a.ybar + b * yrow + o);  // y0 = g(x0) feeds g only (x0 is not trained)
      }
      if (hh == 0) {  // tail chunk [0 0 0 u~ 1] of this step: A columns 64..71 and operand rows 64..71
        float tail[8];
        tail[0] = 0.f; tail[1] = 0.f; tail[2] = 0.f;
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i) tail[3 + i] = (i < a.I && act) ? cc_ut(a.ut, a.u_scale, __ldg(a.u + b * (long long)a.T * a.I + (long long)k * a.I + i)) : 0.f;
        tail[7] = 1.f;
        cc_mlp_store_split<64, 8>(c.t, 64, tail);
        // operand rows 64..71 = [lo(u~_0..2) hi(u~_0..3) 1]: the lo-part rounds leave the tail rows alone, so the remainders of
        // the inputs ride in the three spare rows (their products are added to the input-weight columns by the reduction)
        float trow[8];
#pragma unroll
        for (int i = 0; i < 3; ++i) trow[i] = tail[3 + i] - cc_mlp_tf32_rn(tail[3 + i]);
#pragma unroll
        for (int i = 3; i < 8; ++i) trow[i] = cc_mlp_tf32_rn(tail[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) opb_tail[cc_mlpb_row(i)] = trow[i];
      }
      // ---- first pass: z_2..z_4 (RK4) ----
      {
        float x[PC];
#pragma unroll
        for (int i = 0; i < PC; ++i) x[i] = (act && cb + i < a.S) ? xk[(size_t)(cb + i) * a.Bpad] : 0.f;
#pragma unroll
        for (int c0 = 0; c0 < PC; c0 += 16) {
          float z[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) z[i] = x[c0 + i];
          cc_mlp_store_split<64, 16>(c.t, cb + c0, z);
        }
        cc_mlpb_round(c, 0);
        if (rk4) {
#pragma unroll 1
          for (int st = 0; st < 3; ++st) {
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
              cc_mlpb_await(c);
              const int actl = a.f[l].act;
              const float cz = st == 2 ? 1.f : 0.5f;
#pragma unroll
              for (int c0 = 0; c0 < PC; c0 += 16) {
                float r[16];
                cc_mlp_ld<64, 16>(c.t, cb + c0, r);
                cc_mlp_wait_ld();
                cc_mlp_act_group<16>(actl, r);
                if (l == L - 1) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    r[i] = fmaf(h * r[i], cz, x[c0 + i]);
                    zs_c[(size_t)(st * 64 + c0 + i) * 128] = r[i];
                  }
                }
                cc_mlp_store_split<64, 16>(c.t, cb + c0, r);
              }
              cc_mlpb_round(c, l == L - 1 ? 0 : l + 1);
            }
          }
        }
      }
      // ---- reverse stages ----
      const float cfirst = rk4 ? c6 : (a.integrator == CC_INT_EE ? h : 1.f);
#pragma unroll
      for (int i = 0; i < PC; ++i) {
        dl[i] = cfirst * lam[i];
        lam0[(size_t)i * 128] = lam[i];
        if (a.integrator == CC_INT_NONE) lam[i] = 0.f;  // (no integrator: x_{k+1} = f(x_k), the adjoint is replaced)
      }
#pragma unroll 1
      for (int s = nst - 1; s >= 0; --s) {
        const float* a0 = s == 0 ? xk + (size_t)cb * a.Bpad : zs_c + (size_t)((s - 1) * 64) * 128;  // stage input: x_k or z_{s+1} (slot s-1), this thread's columns
        const size_t a0s = s == 0 ? (size_t)a.Bpad : 128;
        const bool a0ok = s == 0 ? act : true;
        if (s != nst - 1) {  // (the first pass left z_4 in A and launched its round)
#pragma unroll
          for (int c0 = 0; c0 < PC; c0 += 16) {
            float z[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = (a0ok && (s != 0 || cb + c0 + i < a.S)) ? a0[(size_t)(c0 + i) * a0s] : 0.f;
            cc_mlp_store_split<64, 16>(c.t, cb + c0, z);
          }
          cc_mlpb_round(c, 0);
        }
        // hidden activations a_1..a_{L-1} of this stage
#pragma unroll
        for (int l = 0; l < L - 1; ++l) {
          cc_mlpb_await(c);
#pragma unroll
          for (int c0 = 0; c0 < PC; c0 += 16) {
            float r[16];
            cc_mlp_ld<64, 16>(c.t, cb + c0, r);
            cc_mlp_wait_ld();
            cc_mlp_act_group<16>(a.f[l].act, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) hid[l][c0 + i] = r[i];
            if (l < L - 2 || fa) cc_mlp_store_split<64, 16>(c.t, cb + c0, r);
            if (l == L - 2) cc_mlpb_op_rn<16>(c.opb, c0, r);  // operand of the last layer's weight gradient
          }
          if (l < L - 2 || fa) cc_mlpb_round(c, l + 1);
        }
        if (fa) {  // final activation: delta = kbar . act_f'(k_s)
          cc_mlpb_await(c);
#pragma unroll
          for (int c0 = 0; c0 < PC; c0 += 16) {
            float r[16], dd[16];
            cc_mlp_ld<64, 16>(c.t, cb + c0, r);
            cc_mlp_wait_ld();
            cc_mlp_act_group<16>(a.f[L - 1].act, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) dd[i] = dl[c0 + i];
            cc_mlpb_act_grad_group(a.f[L - 1].act, dd, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) dl[c0 + i] = dd[i];
          }
        }
        // reverse chain
#pragma unroll
        for (int l = L - 1; l >= 0; --l) {
#pragma unroll
          for (int c0 = 0; c0 < PC; c0 += 16) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = dl[c0 + i];
            cc_mlpb_store_delta<16>(c, cb, c0, v);
            if (l < L - 1) {  // operand rows = the layer's input: a hidden activation or the stage input
              float in[16];
              if (l > 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) in[i] = hid[l > 0 ? l - 1 : 0][c0 + i];
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) in[i] = (a0ok && (s != 0 || cb + c0 + i < a.S)) ? a0[(size_t)(c0 + i) * a0s] : 0.f;
              }
              cc_mlpb_op_rn<16>(c.opb, c0, in);
            }
          }
          if (l == 0 && s == 0) put_d(dg);
          cc_mlpb_round(c, l | 4 | ((l == 0 && s == 0) ? 8 : 0));
          cc_mlpb_await(c);
          // second weight-gradient round of the layer: the lo part of its input replaces the hi part in the operand rows
#pragma unroll
          for (int c0 = 0; c0 < PC; c0 += 16) {
            float in[16];
            if (l > 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) in[i] = hid[l > 0 ? l - 1 : 0][c0 + i];
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) in[i] = (a0ok && (s != 0 || cb + c0 + i < a.S)) ? a0[(size_t)(c0 + i) * a0s] : 0.f;
            }
            cc_mlpb_op_lo<16>(c.opb, c0, in);
          }
          cc_mlpb_round(c, l | 64 | ((l == 0 && s == 0) ? 8 : 0));
          const float ca = (s - 1 == 0) ? c6 : c3;
          const float cbk = (s - 1 == 2) ? h : c2;
#pragma unroll
          for (int c0 = 0; c0 < PC; c0 += 16) {
            float r[16];
            cc_mlp_ld<64, 16>(c.t, cb + c0, r);
            cc_mlp_wait_ld();
            if (l > 0) {
              cc_mlpb_act_grad_group(a.f[l - 1].act, r, &hid[l > 0 ? l - 1 : 0][c0]);
#pragma unroll
              for (int i = 0; i < 16; ++i) dl[c0 + i] = r[i];
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                lam[c0 + i] += r[i];
                if (s > 0) dl[c0 + i] = fmaf(ca, lam0[(size_t)(c0 + i) * 128], cbk * r[i]);  // kbar of the next (earlier) stage
              }
            }
          }
          cc_mlpb_await(c);  // the operand rows are free again
        }
      }
      flush_dg();
#pragma unroll
      add_readout(d);
      if (hh == 0)
#pragma unroll
        for (int o = 0; o < CC_TINY_O; ++o) gb[o] += dg[o];
      ++since_flush;
      if (since_flush == a.flush_every || k == kbeg) {
        flush_dw();
        since_flush = 0;
      }
#ifdef CC_EMULATE
      c.fresh = since_flush == 0;
#endif
    }
    // g's bias gradient: warp sums in lane order, then the four warps of half 0 in fixed order
#pragma unroll
    for (int o = 0; o < CC_TINY_O; ++o) {
      float v = gb[o];
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, m);
      if (hh == 0 && (row & 31) == 0) gbw[(row >> 5) * 4 + o] = v;
    }
    cc_mlpb_worker_sync<NWORK>();
    float* grec = rec + (size_t)L * 80 * 128;
    for (int i = tid; i < 8 * 64; i += NWORK) grec[i] = (a.seg_accum ? grec[i] : 0.f) + gacc[i];
    if (tid < CC_TINY_O) grec[8 * 64 + tid] = (a.seg_accum ? grec[8 * 64 + tid] : 0.f) + ((gbw[tid] + gbw[4 + tid]) + (gbw[8 + tid] + gbw[12 + tid]));
    if (a.lam_io && act) {  // adjoint of the state at step k0: the next (earlier) segment starts from it
#pragma unroll
      for (int i = 0; i < PC; ++i)
        if (cb + i < a.S) a.lam_io[(size_t)(cb + i) * a.Bpad + b] = lam[i];
    }
  }
#ifndef CC_EMULATE
  cc_tc_fence_before();
  __syncthreads();
  if ((tid >> 5) == 0) cc_tmem_dealloc(c.t.tbase, 512);
#endif
}

// theta_bar[p] = sum over CTAs (fixed order) of the hi-row and lo-row sums of the entry that holds parameter p
__global__ void cc_mlpb_reduce_kernel(const CC_GRID_CONSTANT CcMlpArgs a, float* theta_bar, int n_cta, int P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const size_t stride = CC_MLPB_CTA_FLOATS(a.L);
  size_t e0 = 0, e1 = 0;
  long long e2 = -1;  // input weights: + the products with the inputs' lo parts (operand rows 64..66)
  bool single = false;
  bool found = false;
  for (int l = 0; l < a.L && !found; ++l) {
    const CcMlpLayerArg& Ly = a.f[l];
    if (p >= Ly.w_off && p < Ly.w_off + Ly.N * Ly.in_log) {
      const int n = (p - Ly.w_off) / Ly.in_log, j = (p - Ly.w_off) % Ly.in_log;
      const int col = (l == 0 && j >= a.S) ? 67 + (j - a.S) : j;
      e0 = (size_t)(l * 80 + col) * 128 + n; e1 = e0 + 64; found = true;
      if (l == 0 && j >= a.S && j - a.S < 3) e2 = (long long)(64 + (j - a.S)) * 128 + n;
    } else if (Ly.b_off >= 0 && p >= Ly.b_off && p < Ly.b_off + Ly.N) {
      e0 = (size_t)(l * 80 + 71) * 128 + (p - Ly.b_off); e1 = e0 + 64; found = true;
    }
  }
  const size_t gbase = (size_t)a.L * 80 * 128;
  if (!found && p >= a.g.w_off && p < a.g.w_off + a.g.N * a.g.in_log) {
    const int o = (p - a.g.w_off) / a.g.in_log, s = (p - a.g.w_off) % a.g.in_log;
    e0 = gbase + o * 64 + s; e1 = gbase + (4 + o) * 64 + s; found = true;
  } else if (!found && a.g.b_off >= 0 && p >= a.g.b_off && p < a.g.b_off + a.g.N) {
    e0 = gbase + 8 * 64 + (p - a.g.b_off); single = true; found = true;
  }
  float s = 0.f;
  if (found)
    for (int cta = 0; cta < n_cta; ++cta) {
      const float* r = a.recs + (size_t)cta * stride;
      s += single ? r[e0] : r[e0] + r[e1];
      if (e2 >= 0) s += r[e2] + r[e2 + 64];
    }
  theta_bar[p] = s;
}
