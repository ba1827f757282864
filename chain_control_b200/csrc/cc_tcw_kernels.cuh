// cc_tcw_kernels.cuh -- tcgen05 reverse pass of the OPEN loop with a TRAINABLE linear depth-0 model at large batch
// (ModelTrainer, step_fn.py:133-146; compact model neural_ode_model_compact_example.py:85-124: f, g single Linear layers
// without final activation, readout of the state BEFORE the update, RK4).  theta_bar and, on request, ubar.
//
// The weight gradient  Wbar[n][j] = sum over trajectories, steps and stages of kbar_s[n] * a_s[j]  contracts over the
// TRAJECTORIES, so it is a tensor-core product whose K dimension is the batch tile: both operands are written K-major
// (K = trajectory) into shared memory by the thread that owns the trajectory and multiplied with SS-mode tcgen05.mma
// into a TMEM accumulator that lives for the whole step.
//
//   * one CTA = 128 trajectories, 3 warpgroups, one CTA per SM (the two operand buffers take 144 KB):
//       warps 0-3  adjoint chain   : kbar_4..kbar_1 of step k  (A from TMEM . W^T from smem, like cc_tc_bwd_kernel)
//       warps 4-7  recompute chain : stage inputs a_1..a_4 of step k-1, ONE STEP AHEAD (A from TMEM . W from smem);
//                                    z_2..z_4 wait in a per-CTA global scratch (3 rotating slots, L2 resident, written
//                                    and read by the same thread) until the adjoint chain reaches their stage
//       warp 8     issues every MMA of a window and commits (lane 0); its warpgroup gives its registers to the
//                  workers (setmaxnreg: 56 / 224 per thread)
//     The two chains are independent, so each hides the other's TMEM round trips; every window (4 per step) has one
//     arrival barrier (256 threads), one batch  [adjoint, recompute] -> d_ready ; [weight gradient] -> w_done.
//   * 3xTF32 everywhere.  Weight gradient with the split rows STACKED along M:  A = [kbar_hi ; kbar_lo] (rows n / 64+n),
//     B = a_hi (accumulator 1) and a_lo (accumulator 2): 2 x 16 MMAs (M = 128, N = 64, K = 8 trajectories each) give
//     all four partial products.  g's gradient rides along: rows 56+o / 120+o of A hold the readout cotangent d_o
//     (non-zero in the stage whose input is x_k).
//   * the tensor core accumulates with truncation (measured: -2.4e-4 relative bias after 4096 same-sign accumulations,
//     scratch/tc_probe5.cu), so the accumulators are flushed into FP32 registers (round to nearest) after EVERY step.
//   * operand layout: no swizzle, core matrix = 8 rows x 4 trajectories (128 B) padded to 144 B: thread t writes 4 bytes at
//     (t/4)*144 + (t%4)*4 + row offset -> conflict-free stores; LBO = 144 B, SBO = 32*144 B.
#pragma once
// Ablation switches used for the round-1 timeline (skip the weight-gradient MMAs / the operand flush: WRONG gradients).
// Compile-time only (-DCC_TCW_PROFILE=<bits>); a product build has them folded to 0.
#ifdef CC_TCW_PROFILE
#define CC_TCW_SKIP(bit) ((CC_TCW_PROFILE) & (bit))
#else
#define CC_TCW_SKIP(bit) 0
#endif
#include "cc_tc_kernels.cuh"

#define CC_TCW_THREADS 384  // 3 warpgroups: adjoint chain, recompute chain, MMA issue (warp 8; setmaxnreg moves its registers to the workers)
#define CC_TCW_OPER_FLOATS (16 * 1152)  // 128 rows x 128 trajectories with padded core matrices: 73,728 B
#define CC_TCW_TMEM_COLS 512
#define CC_TCW_DB 0     // D of the adjoint chain        [zbar (SR) ; vbar (4)]
#define CC_TCW_DF 64    // D of the recompute chain      k (SR)
#define CC_TCW_DW1 128  // weight gradient, B = a_hi
#define CC_TCW_DW2 192  // weight gradient, B = a_lo
#define CC_TCW_ABH 256  // A of the adjoint chain (kbar) hi / lo
#define CC_TCW_ABL 312
#define CC_TCW_AFH 368  // A of the recompute chain (stage input) hi / lo
#define CC_TCW_AFL 432
#define CC_TCW_ROW_D 56  // operand-A rows of the readout cotangents (lo part 64 rows further)
#define CC_TCW_SCRATCH_FLOATS(SR) (3 * (SR) * 128 + 64 * 128)  // stage-input slots + flushed weight-gradient sums [64 columns][128 rows]
// floats of the shared-memory region behind p.tc_off (SR = 8 * nch)
#define CC_TCW_REGION_FLOATS(SR) (2 * (SR) * (((SR) + CC_TINY_O + 15) / 16 * 16) + 2 * (SR) * (((SR) + 15) / 16 * 16) + CC_TINY_O * ((SR) + 4) + 8 + 32 + 2 * CC_TCW_OPER_FLOATS)

// float offset of operand row r (thread part added separately)
CC_DEV constexpr int cc_tcw_row(int r) { return (r >> 3) * 1152 + (r & 7) * 4; }

// descriptors are derived from six pre-encoded bases by one 64-bit add each (the start-address field holds address >> 4 and
// never overflows: shared memory is < 256 KB); the bases are made opaque once per window so that the compiler does not
// keep ~150 loop-invariant descriptors alive in the issuing thread's 56 registers
CC_DEV uint64_t cc_tcw_desc(uint64_t base, uint32_t byte_off) { return base + (uint64_t)(byte_off >> 4); }
CC_DEV void cc_tcw_opaque(uint64_t& d) { asm volatile("" : "+l"(d)); }
template <int NCH>
CC_DEV void cc_tcw_issue_af(uint32_t tbase, uint64_t wbh, uint64_t wbl, uint64_t wfh, uint64_t wfl, bool adj, bool fwd) {
  constexpr int SR = 8 * NCH, NPB = (SR + CC_TINY_O + 15) / 16 * 16, NPF = (SR + 15) / 16 * 16;
  constexpr uint32_t idb = cc_umma_idesc_tf32(128, NPB), idf = cc_umma_idesc_tf32(128, NPF);
  constexpr uint32_t ksb = 2u * NPB * 16u, ksf = 2u * NPF * 16u;
  cc_tcw_opaque(wbh); cc_tcw_opaque(wbl); cc_tcw_opaque(wfh); cc_tcw_opaque(wfl);
  // D = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi ; the two chains alternate (different accumulators back to back)
#pragma unroll
  for (int ps = 0; ps < 3; ++ps) {
#pragma unroll
    for (int js = 0; js < NCH; ++js) {
      const uint32_t acc = (ps | js) ? 1u : 0u;
      if (adj) cc_umma_tf32_ts(tbase + CC_TCW_DB, tbase + (ps == 0 ? CC_TCW_ABL : CC_TCW_ABH) + 8 * js, cc_tcw_desc(ps == 1 ? wbl : wbh, js * ksb), idb, acc);
      if (fwd) cc_umma_tf32_ts(tbase + CC_TCW_DF, tbase + (ps == 0 ? CC_TCW_AFL : CC_TCW_AFH) + 8 * js, cc_tcw_desc(ps == 1 ? wfl : wfh, js * ksf), idf, acc);
    }
  }
}
CC_DEV void cc_tcw_issue_w(uint32_t tbase, uint64_t opa, uint64_t opb, uint32_t acc0) {
  constexpr uint32_t idw = cc_umma_idesc_tf32(128, 64);
  cc_tcw_opaque(opa); cc_tcw_opaque(opb);
#pragma unroll
  for (int js = 0; js < 16; ++js) {  // 8 trajectories (two core matrices) per instruction
    const uint64_t ad = cc_tcw_desc(opa, js * 288u);
    cc_umma_tf32_ss(tbase + CC_TCW_DW1, ad, cc_tcw_desc(opb, js * 288u), idw, js ? 1u : acc0);
    cc_umma_tf32_ss(tbase + CC_TCW_DW2, ad, cc_tcw_desc(opb, 8u * 4608u + js * 288u), idw, js ? 1u : acc0);
  }
}

// split NC values into tf32 hi / exact lo and store them both to the TMEM A operand (columns c0..) and to the shared-memory
// operand of the weight gradient (rows c0.. hi, 64 + c0.. lo; `op` already points at this thread's trajectory column)
template <int NC>
CC_DEV void cc_tcw_store_split2(uint32_t aHi, uint32_t aLo, float* op, int c0_dummy, const float (&z)[NC], const int C0) {
  (void)c0_dummy;
#pragma unroll
  for (int j = 0; j < NC / 8; ++j) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const float v0 = z[8 * j + i], v1 = z[8 * j + i + 1];
      hi[i] = __float_as_uint(v0) & 0xFFFFE000u;
      hi[i + 1] = __float_as_uint(v1) & 0xFFFFE000u;
      const float2 l = __fadd2_rn(make_float2(v0, v1), make_float2(-__uint_as_float(hi[i]), -__uint_as_float(hi[i + 1])));
      lo[i] = __float_as_uint(l.x);
      lo[i + 1] = __float_as_uint(l.y);
    }
    cc_tmem_st8(aHi + C0 + 8 * j, hi);
    cc_tmem_st8(aLo + C0 + 8 * j, lo);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      op[cc_tcw_row(C0 + 8 * j + i)] = __uint_as_float(hi[i]);
      op[cc_tcw_row(64 + C0 + 8 * j + i)] = __uint_as_float(lo[i]);
    }
  }
}
// generic adjoint stage for columns [C0, C0+NC): zb = D ; zsum += zb ; kbar = ca*lam + cb*zb -> A (TMEM) and operand rows.
// `before_store` runs between the arithmetic and the first store (the operand buffer is free only behind w_done)
template <int SR, int C0, int NC, class W>
CC_DEV void cc_tcw_adj_group(uint32_t aD, uint32_t aHi, uint32_t aLo, float* opa, float (&lam)[SR], float (&zsum)[SR], float ca, float cb, W&& before_store) {
  uint32_t r[NC];
  cc_tc_ld<NC>(aD + C0, r);
  cc_tmem_wait_ld();
  float kb[NC];
  const float2 ca2 = make_float2(ca, ca), cb2 = make_float2(cb, cb);
#pragma unroll
  for (int i = 0; i < NC; i += 2) {
    const float2 zb = make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
    const float2 zs = __fadd2_rn(make_float2(zsum[C0 + i], zsum[C0 + i + 1]), zb);
    zsum[C0 + i] = zs.x; zsum[C0 + i + 1] = zs.y;
    const float2 kk = __ffma2_rn(ca2, make_float2(lam[C0 + i], lam[C0 + i + 1]), __fmul2_rn(cb2, zb));
    kb[i] = kk.x; kb[i + 1] = kk.y;
  }
  before_store();
  cc_tcw_store_split2<NC>(aHi, aLo, opa, 0, kb, C0);
}

// first window of a step (adjoint chain): finish the previous step (lam <- lam + zsum + zbar_1 + g^T d), then kbar_4 = c0*lam -> A
template <int SR, int C0, int NC, class W>
CC_DEV void cc_tcw_adj_first(uint32_t aD, uint32_t aHi, uint32_t aLo, float* opa, const float* gw, float (&lam)[SR], float (&zsum)[SR], bool have_prev,
                             const float (&dprev)[CC_TINY_O], int O, float c0, W&& before_store) {
  uint32_t r[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) r[i] = 0u;
  if (have_prev) {
    cc_tc_ld<NC>(aD + C0, r);
    cc_tmem_wait_ld();
  }
  float kb[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    float l = lam[C0 + i];
    if (have_prev) {
      l = l + (zsum[C0 + i] + __uint_as_float(r[i]));  // (zbar_1 already contains g^T d: the readout cotangent rides in kbar_1's columns S + o)
    }
    lam[C0 + i] = l;
    zsum[C0 + i] = 0.f;
    kb[i] = c0 * l;
  }
  before_store();
  cc_tcw_store_split2<NC>(aHi, aLo, opa, 0, kb, C0);
}
// recompute chain: z = x + (h*k)*cz for columns [C0, C0+NC) -> A (if another stage follows) and -> the scratch slot
template <int SR, int C0, int NC>
CC_DEV void cc_tcw_fwd_group(uint32_t aD, uint32_t aHi, uint32_t aLo, const float (&x)[SR], float h, float cz, bool to_tmem, float* slot) {
  uint32_t r[NC];
  cc_tc_ld<NC>(aD + C0, r);
  cc_tmem_wait_ld();
  float z[NC];
  const float2 h2 = make_float2(h, h), cz2 = make_float2(cz, cz);
#pragma unroll
  for (int i = 0; i < NC; i += 2) {
    const float2 k = make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1]));
    const float2 zz = __ffma2_rn(__fmul2_rn(h2, k), cz2, make_float2(x[C0 + i], x[C0 + i + 1]));
    z[i] = zz.x; z[i + 1] = zz.y;
  }
  if (to_tmem) cc_tc_store_split<NC>(aHi, aLo, C0, z);
#pragma unroll
  for (int i = 0; i < NC; ++i) slot[(C0 + i) * 128] = z[i];
}

template <int NCH>
__global__ void __launch_bounds__(CC_TCW_THREADS, 1) cc_tcw_bwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  CC_DIRECT_PARAMS(gp)
  (void)cta_smem;
  constexpr int SR = 8 * NCH, NPB = (SR + CC_TINY_O + 15) / 16 * 16, NPF = (SR + 15) / 16 * 16;
  const CcNodeDev& M = p.nd[0];
  const CcLayerDev& F = M.f.layer[0];
  const CcLayerDev& G = M.g.layer[0];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* wbh = cc_smem + p.tc_off;
  float* wbl = wbh + SR * NPB;
  float* wfh = wbl + SR * NPB;
  float* wfl = wfh + SR * NPF;
  float* gw = wfl + SR * NPF;
  uint64_t* a_ready = reinterpret_cast<uint64_t*>(gw + CC_TINY_O * (SR + 4));
  uint64_t* d_ready = a_ready + 1;
  uint64_t* w_done = a_ready + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 3);
  float* opA = gw + CC_TINY_O * (SR + 4) + 8;
  opA += ((128u - (cc_smem_u32(opA) & 127u)) & 127u) / 4;  // operand buffers on a 128-byte boundary
  float* opB = opA + CC_TCW_OPER_FLOATS;
  cc_tc_stage_b<SR>(M, p.theta[0], SR, NPB, wbh, wbl, true, tid, blockDim.x, true);
  cc_tc_stage_b<SR>(M, p.theta[0], SR, NPF, wfh, wfl, false, tid, blockDim.x);
  cc_tc_stage_gw<SR>(M, p.theta[0], gw, tid, blockDim.x);
  if (warp == 8) {
    if (lane == 0) {
      cc_mbar_init(a_ready, 256);
      cc_mbar_init(d_ready, 1);
      cc_mbar_init(w_done, 1);
      cc_mbar_fence_init();
    }
    __syncwarp();
    cc_tmem_alloc(tmem_slot, CC_TCW_TMEM_COLS);
  }
  cc_fence_proxy_async();
  cc_tc_fence_before();
  __syncthreads();
  cc_tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const int T = p.T;
  const float h = M.dt;
  const float os = M.out_scale;
  const int t = tid & 127;                     // trajectory inside the tile = TMEM lane = operand K index
  const long long b = (long long)blockIdx.x * 128 + t;
  const bool act = b < p.B;
  const uint32_t lb = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  const int tho = (t >> 2) * 36 + (t & 3);     // this thread's float offset inside an operand row group
  uint32_t pd = 0, pw = 0;
  auto publish = [&]() {
    cc_tmem_wait_st();
    cc_fence_proxy_async();  // the operand stores (generic proxy) must be visible to the MMAs (async proxy)
    cc_tc_fence_before();
    cc_mbar_arrive(a_ready);
  };
  auto await_d = [&]() { cc_mbar_wait(d_ready, pd); pd ^= 1; cc_tc_fence_after(); };
  auto await_w = [&]() { cc_mbar_wait(w_done, pw); pw ^= 1; cc_tc_fence_after(); };
  // the flushed weight-gradient sums (row = TMEM lane = this thread, FP32 round-to-nearest adds) live in the CTA's global
  // scratch [column][row]: warps 0-3 own columns 0..31, warps 4-7 columns 32..63 (registers are needed for the chains)
  float* wsum = p.scratch + (long long)blockIdx.x * CC_TCW_SCRATCH_FLOATS(SR) + 3 * SR * 128 + t;
  bool flushed = false;
  auto flush = [&](int c0) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r1[16], r2[16];
      cc_tmem_ld16(lb + CC_TCW_DW1 + c0 + 16 * c, r1);
      cc_tmem_ld16(lb + CC_TCW_DW2 + c0 + 16 * c, r2);
      float old[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) old[i] = flushed ? wsum[(c0 + 16 * c + i) * 128] : 0.f;
      cc_tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) wsum[(c0 + 16 * c + i) * 128] = old[i] + (__uint_as_float(r1[i]) + __uint_as_float(r2[i]));
    }
    flushed = true;
  };

  if (warp >= 8) {
    // warps 9-11 only fill the issuing warp's warpgroup (setmaxnreg is a warpgroup-wide instruction)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 8) {
    // ---------------- MMA issue ----------------
    const uint64_t wbh_u = cc_umma_desc_kmajor(cc_smem_u32(wbh), NPB * 16u, 128), wbl_u = cc_umma_desc_kmajor(cc_smem_u32(wbl), NPB * 16u, 128);
    const uint64_t wfh_u = cc_umma_desc_kmajor(cc_smem_u32(wfh), NPF * 16u, 128), wfl_u = cc_umma_desc_kmajor(cc_smem_u32(wfl), NPF * 16u, 128);
    const uint64_t opa_u = cc_umma_desc_kmajor(cc_smem_u32(opA), 144, 4608), opb_u = cc_umma_desc_kmajor(cc_smem_u32(opB), 144, 4608);
    uint32_t par = 0;
#pragma unroll 1
    for (int j = T; j >= 0; --j) {
#pragma unroll 1
      for (int w = 0; w < 4; ++w) {
        cc_mbar_wait(a_ready, par);
        par ^= 1;
        cc_tc_fence_after();
        if (lane == 0) {
          const bool adj = j < T, fwd = j >= 1 && w < 3;
          cc_tcw_issue_af<NCH>(tbase, wbh_u, wbl_u, wfh_u, wfl_u, adj, fwd);
          cc_tc_commit(d_ready);
          if (adj && !CC_TCW_SKIP(1)) cc_tcw_issue_w(tbase, opa_u, opb_u, w == 0 ? 0u : 1u);
          cc_tc_commit(w_done);
        }
        __syncwarp();
      }
    }
    }
  } else {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
  if (warp < 4) {
    // ---------------- adjoint chain ----------------
    const uint32_t aD = lb + CC_TCW_DB, aHi = lb + CC_TCW_ABH, aLo = lb + CC_TCW_ABL;
    float* opa = opA + tho;
    float lam[SR], zsum[SR];
#pragma unroll
    for (int i = 0; i < SR; ++i) { lam[i] = 0.f; zsum[i] = 0.f; }
    float utb[CC_TINY_O], d[CC_TINY_O], dprev[CC_TINY_O], ucur[CC_TINY_O];
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) { utb[i] = 0.f; d[i] = 0.f; dprev[i] = 0.f; ucur[i] = 0.f; }
    const long long yrow = (long long)(T + 1) * M.O;
    float nyb[CC_TINY_O], nu[CC_TINY_O];
    auto fetch_yu = [&](int k) {  // ybar of y_{k+1} = g(x_k) (y_0 = g(x_0) joins step 0) and u_k, one step ahead of their use
#pragma unroll
      for (int o = 0; o < CC_TINY_O; ++o) {
        nyb[o] = 0.f;
        if (o < M.O && act) {
          nyb[o] = __ldg(p.ybar + b * yrow + (long long)(k + 1) * M.O + o);
          if (k == 0) nyb[o] += __ldg(p.ybar + b * yrow + o);
        }
        nu[o] = (o < M.I && act) ? __ldg(p.in + b * (long long)T * M.I + (long long)k * M.I + o) : 0.f;
      }
    };
    fetch_yu(T - 1);
    auto write_ubar = [&](int k) {  // cotangent of the raw input of step k (its four vbar are summed in utb; ucur = u_k)
      if (p.ubar && act)
#pragma unroll
        for (int i = 0; i < CC_TINY_O; ++i)
          if (i < M.I) p.ubar[b * (long long)T * M.I + (long long)k * M.I + i] = utb[i] * cc_ut_grad(M.ut, M.u_scale, ucur[i]);
    };
#pragma unroll 1
    for (int j = T; j >= 0; --j) {
      const bool on = j < T;
#pragma unroll 1
      for (int w = 0; w < 4; ++w) {
        const bool first = j == T && w == 0;
        if (!first) await_d();
        bool waited = first;  // the operand buffer may be rewritten only behind w_done: waited for right before the first store
        auto before_store = [&]() { if (!waited) { await_w(); waited = true; } };
        auto nop = []() {};
        if (on) {
          if (w == 0) {
            const bool have_prev = j < T - 1;
            if (have_prev) {
              uint32_t r[4];
              cc_tmem_ld4(aD + SR, r);
              cc_tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < CC_TINY_O; ++i) utb[i] += __uint_as_float(r[i]);
              write_ubar(j + 1);
            }
#pragma unroll
            for (int i = 0; i < CC_TINY_O; ++i) { utb[i] = 0.f; dprev[i] = d[i]; }
            const float c0 = h / 6.f;
            cc_tcw_adj_first<SR, 0, 32>(aD, aHi, aLo, opa, gw, lam, zsum, have_prev, dprev, M.O, c0, before_store);
            if constexpr (SR == 56) {
              cc_tcw_adj_first<SR, 32, 16>(aD, aHi, aLo, opa, gw, lam, zsum, have_prev, dprev, M.O, c0, nop);
              cc_tcw_adj_first<SR, 48, 8>(aD, aHi, aLo, opa, gw, lam, zsum, have_prev, dprev, M.O, c0, nop);
            }
            // this step's readout cotangent and raw input (fetched three windows ago)
#pragma unroll
            for (int o = 0; o < CC_TINY_O; ++o) { d[o] = os * nyb[o]; ucur[o] = nu[o]; }
          } else {
            const float ca = w == 3 ? h / 6.f : h / 3.f, cb = w == 1 ? h : h / 2.f;
            cc_tcw_adj_group<SR, 0, 32>(aD, aHi, aLo, opa, lam, zsum, ca, cb, before_store);
            if constexpr (SR == 56) {
              cc_tcw_adj_group<SR, 32, 16>(aD, aHi, aLo, opa, lam, zsum, ca, cb, nop);
              cc_tcw_adj_group<SR, 48, 8>(aD, aHi, aLo, opa, lam, zsum, ca, cb, nop);
            }
            uint32_t r[4];
            cc_tmem_ld4(aD + SR, r);  // vbar columns
            cc_tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < CC_TINY_O; ++i) utb[i] += __uint_as_float(r[i]);
            if (w == 3) {  // stage 1: the readout cotangent joins kbar_1 in the (otherwise zero) columns S + o -> zbar_1 += g^T d
              uint32_t dh[4], dl[4];
#pragma unroll
              for (int o = 0; o < CC_TINY_O; ++o) {
                dh[o] = __float_as_uint(d[o]) & 0xFFFFE000u;
                dl[o] = __float_as_uint(d[o] - __uint_as_float(dh[o]));
              }
              cc_tmem_st4(aHi + M.S, dh);
              cc_tmem_st4(aLo + M.S, dl);
            }
          }
        }
        before_store();
        if (on && !CC_TCW_SKIP(2)) {
          if (w == 0 && j < T - 1) flush(0);
          if (w == 0 || w == 3) {
#pragma unroll
            for (int o = 0; o < CC_TINY_O; ++o) {
              const float v = w == 3 ? d[o] : 0.f;
              const float vh = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
              opa[cc_tcw_row(CC_TCW_ROW_D + o)] = vh;
              opa[cc_tcw_row(64 + CC_TCW_ROW_D + o)] = v - vh;
            }
          }
        }
        publish();
        if (on && w == 1 && j >= 1) fetch_yu(j - 1);
      }
    }
    await_d();
    await_w();
    {  // step 0: its last vbar, and the last flush
      uint32_t r[4];
      cc_tmem_ld4(aD + SR, r);
      cc_tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) utb[i] += __uint_as_float(r[i]);
      write_ubar(0);
      flush(0);
    }
  } else {
    // ---------------- recompute chain (one step ahead) + operand B of the weight gradient ----------------
    const uint32_t aD = lb + CC_TCW_DF, aHi = lb + CC_TCW_AFH, aLo = lb + CC_TCW_AFL;
    float* opb = opB + tho;
    float* scr = p.scratch + (long long)blockIdx.x * CC_TCW_SCRATCH_FLOATS(SR) + t;
    float x[SR], za[SR];
#pragma unroll
    for (int i = 0; i < SR; ++i) { x[i] = 0.f; za[i] = 0.f; }
    float vdut[CC_TINY_O], vch[CC_TINY_O], un[CC_TINY_O];
#pragma unroll
    for (int i = 0; i < CC_TINY_O; ++i) { vdut[i] = 0.f; vch[i] = 0.f; un[i] = 0.f; }
    // tape rows of one step are Bpad floats apart: walk a pointer instead of recomputing 64-bit products per row; lanes
    // beyond the batch read one valid finite dummy location (stride 0) so that no load needs the per-lane predicate
    const long long tstride = act ? p.Bpad : 0;
    auto tape_rows_of = [&](int k) -> const float* { return act ? p.tape + (long long)k * p.tape_rows * p.Bpad + b : p.theta[0]; };
    auto load_x = [&](int k) {  // x_k from the tape ([step][row][Bpad]: coalesced) and the raw input u_k
      const float* q = tape_rows_of(k);
#pragma unroll
      for (int i = 0; i < SR; ++i) {
        x[i] = 0.f;
        if (i < M.S) x[i] = *q;
        q += tstride;
      }
      if (!act) {
#pragma unroll
        for (int i = 0; i < SR; ++i) x[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < CC_TINY_O; ++i) un[i] = (i < M.I && act) ? __ldg(p.in + b * (long long)T * M.I + (long long)k * M.I + i) : 0.f;
    };
    auto load_slot = [&](int s) {  // (the slots cover all 128 lanes and are written before they are read: no predicate)
      const float* q = scr + s * SR * 128;
#pragma unroll
      for (int i = 0; i < SR; ++i) za[i] = q[i * 128];
    };
    load_x(T - 1);
#pragma unroll 1
    for (int j = T; j >= 0; --j) {
      const int c = j - 1;  // step of the recompute chain; j = step whose stage inputs go into operand B
      const bool chain = c >= 0, duty = j < T;
      // slots: z_3 always 1; z_2 / z_4 swap with the step parity (a slot is rewritten in the window after its last read)
      const int s2c = (c & 1) ? 2 : 0, s4c = (c & 1) ? 0 : 2, s2j = (j & 1) ? 2 : 0, s4j = (j & 1) ? 0 : 2;
#pragma unroll 1
      for (int w = 0; w < 4; ++w) {
        const bool first = j == T && w == 0;
        if (!first) await_d();
        if (chain) {
          if (w == 0) {
#pragma unroll
            for (int i = 0; i < CC_TINY_O; ++i) {
              vdut[i] = vch[i];
              vch[i] = (i < M.I) ? cc_ut(M.ut, M.u_scale, un[i]) : 0.f;
              x[SR - 5 + i] = vch[i];
            }
            x[SR - 1] = 1.f;
            cc_tc_write_state<SR>(aHi, aLo, x, false, vch);
          } else {
            const float cz = w == 3 ? 1.f : 0.5f;
            float* slot = scr + (w == 1 ? s2c : (w == 2 ? 1 : s4c)) * SR * 128;
            const bool tt = w < 3;
            cc_tcw_fwd_group<SR, 0, 32>(aD, aHi, aLo, x, h, cz, tt, slot);
            if constexpr (SR == 56) {
              cc_tcw_fwd_group<SR, 32, 16>(aD, aHi, aLo, x, h, cz, tt, slot);
              cc_tcw_fwd_group<SR, 48, 8>(aD, aHi, aLo, x, h, cz, tt, slot);
            }
          }
        } else if (w == 0) {
#pragma unroll
          for (int i = 0; i < CC_TINY_O; ++i) vdut[i] = vch[i];
        }
        if (!first) await_w();
        if (duty && !CC_TCW_SKIP(2)) {
          if (w == 0) {
            if (j < T - 1) flush(32);
#pragma unroll
            for (int i = 0; i < CC_TINY_O; ++i) {
              if (i < M.I) {
                const float vh = __uint_as_float(__float_as_uint(vdut[i]) & 0xFFFFE000u);
                opb[cc_tcw_row(0) + ((M.row_v + i) >> 3) * 1152 + ((M.row_v + i) & 7) * 4] = vh;
                opb[cc_tcw_row(64) + ((M.row_v + i) >> 3) * 1152 + ((M.row_v + i) & 7) * 4] = vdut[i] - vh;
              }
            }
            opb[cc_tcw_row(0) + (M.row_one >> 3) * 1152 + (M.row_one & 7) * 4] = 1.f;
            opb[cc_tcw_row(64) + (M.row_one >> 3) * 1152 + (M.row_one & 7) * 4] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < SR; ++i) {
            if (i < M.S) {
              const float vh = __uint_as_float(__float_as_uint(za[i]) & 0xFFFFE000u);
              opb[cc_tcw_row(i)] = vh;
              opb[cc_tcw_row(64 + i)] = za[i] - vh;
            }
          }
        }
        publish();
        // loads for the next window (their latency hides behind the MMAs)
        if (w < 2) { if (duty) load_slot(w == 0 ? 1 : s2j); }
        else if (w == 2) {
          if (duty) {
            const float* q = tape_rows_of(j);
#pragma unroll
            for (int i = 0; i < SR; ++i) {
              if (i < M.S) za[i] = *q;
              q += tstride;
            }
            if (!act) {
#pragma unroll
              for (int i = 0; i < SR; ++i) za[i] = 0.f;
            }
          }
        } else {
          if (c >= 1) load_x(c - 1);
          if (j >= 1) load_slot((j - 1) & 1 ? 0 : 2);  // z_4 of step j-1, written in this window
        }
      }
    }
    await_d();
    await_w();
    flush(32);
  }
  }
  // ---------------- gradient record of this tile (sums of the hi and lo rows) ----------------
  cc_tc_fence_before();
  __syncthreads();
  {
    const float* ws = p.scratch + (long long)blockIdx.x * CC_TCW_SCRATCH_FLOATS(SR) + 3 * SR * 128;
    float* rec = p.gpart + (long long)blockIdx.x * p.g_rec + M.g_base;
    const int K0 = M.K0;
    for (int e = tid; e < M.S * K0; e += blockDim.x) {
      const int n = e / K0, jj = e % K0;
      rec[F.g_off + e] = ws[jj * 128 + n] + ws[jj * 128 + 64 + n];
    }
    for (int e = tid; e < M.O * K0; e += blockDim.x) {
      const int o = e / K0, jj = e % K0;
      const bool real = jj < M.S || jj == M.row_one;
      rec[G.g_off + e] = real ? ws[jj * 128 + CC_TCW_ROW_D + o] + ws[jj * 128 + 64 + CC_TCW_ROW_D + o] : 0.f;
    }
  }
  __syncthreads();
  if (warp == 8) cc_tmem_dealloc(tbase, CC_TCW_TMEM_COLS);
}
