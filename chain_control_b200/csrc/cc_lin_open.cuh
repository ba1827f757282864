// cc_lin_open.cuh -- blocked linear kernels for the OPEN loop (family 4, cc_lin.h, "X" formulation):
//
//   cc_lin_ol_fwd_kernel : y = vmap(model.unroll)(u)  (cc/core/abstract.py:54-70, cc/train/step_fn.py:124,136) for a linear
//                          depth-0 model: per block of m steps ONE MMA batch
//                              [x_{k+m} - x_k ; y_{k+1} .. y_{k+m}] = Wf [x_k ; u~_k .. u~_{k+m-1} ; 1]
//                          (no per-step work at all: the thread only moves u in and y out); the tape is x_k at block starts.
//   cc_linw_bwd_kernel   : the reverse pass of a TRAINABLE linear model (ModelTrainer, step_fn.py:133-146):
//                              [mu_k - mu_E ; ubar~_k .. ] = Wb [mu_E ; ybar_{k+1} .. ybar_{k+m}]
//                          and Q += Lam^T X over the 128 trajectories of the tile as a second, SS-mode tcgen05.mma batch
//                          (contraction over the trajectories; operands K-major in shared memory, 3xTF32 with the hi/lo
//                          rows stacked along M like cc_tcw_kernels.cuh).  theta_bar follows from Q in cc_lin_grad_*.
//   cc_lin_qreduce_kernel, cc_lin_grad_kernel, cc_lin_grad_final_kernel: fp64 post-processing  Q -> theta_bar
//                          (tests/linblock_ref.py: rollout_bwd / matpoly_grad state the same algebra in NumPy).
#pragma once
#include "cc_lin_kernels.cuh"
#include "cc_lin_prep.cuh"

// ------------------------------------------------------------------------------------------------
// K1: open-loop forward
// ------------------------------------------------------------------------------------------------
template <int NCH, int DUMMY>
__global__ void __launch_bounds__(CC_LIN_THREADS, 2) cc_lin_ol_fwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  CC_DIRECT_PARAMS(gp)
  (void)cta_smem;
  const CcNodeDev& M = p.nd[0];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const CcLinShared sh = cc_lin_carve(cc_smem);
  CcLinTm tm;
  cc_lin_prologue(p, sh, false, tm, CC_LIN_SM_STAGE + 4 * 4 * CC_LIN_TILE);
  const int S = M.S, I = M.I, O = M.O, m = p.lin_m, T = p.T;
  const long long b0 = (long long)blockIdx.x * 128 + warp * 32;
  const long long rem = p.B - b0;
  const int nact = rem >= 32 ? 32 : (rem > 0 ? (int)rem : 0);
  const bool act = lane < nact;
  const long long b = b0 + lane;
  const bool stg = p.lin_stage != 0;  // I = O = 1: coalesced tiles for u and y
  // per warp: two input tiles and two output tiles (chunk c of 32 steps in slot c & 1): a block of m <= 13 steps touches at
  // most two chunks, and the tile copies are issued ONCE per block (they must not sit inside the unrolled per-slot code: 13
  // copies of a 32-iteration loop thrash the instruction cache).  4 tiles per warp keep two CTAs per SM.
  float* tin = sh.stage + warp * 4 * CC_LIN_TILE;
  float* tout = tin + 2 * CC_LIN_TILE;
  const float* small = p.lin_pack + CC_LIN_SMALL;
  float x[8 * NCH];
#pragma unroll
  for (int i = 0; i < 8 * NCH; ++i) x[i] = (p.x0 && act && i < S) ? __ldg(p.x0 + b * S + i) : 0.f;
  // y0 = g(x0)      neural_ode_model.py:160-175
  if (act) {
    for (int o = 0; o < O; ++o) {
      float s0 = small[CC_LIN_S_D0 + o];
#pragma unroll
      for (int i = 0; i < 8 * NCH; ++i) s0 = fmaf(small[CC_LIN_S_G0 + o * 64 + i], x[i], s0);
      p.out[b * (long long)(T + 1) * O + o] = s0;
    }
  }
  // state part of the first operand, constant 1
  if constexpr (NCH == 8) x[CC_LIN_ONE] = 1.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    float z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = x[8 * c + i];
    cc_lin_store_split<8>(tm, 8 * c, z);
  }
  if constexpr (NCH != 8) cc_lin_store_split1(tm, CC_LIN_ONE, 1.f);
  int issued = 0, waited = 0;  // input chunks whose copy has been issued / has landed
  if (stg) { cc_lin_tile_load(tin, p.in, b0, nact, T, T, 0, lane); issued = 1; }
  int blk = 0;
  // y of step k is row k+1 of the output: the output tile of chunk c covers rows [32c + 1, 32c + 33)
  for (int k0 = 0; k0 < T; k0 += m, ++blk) {
    const int nst = T - k0 < m ? T - k0 : m;
    if (p.tape && act) {  // block-start state: the tape of the trainable reverse pass  [block][s][Bpad]
      float* tp = p.tape + ((long long)blk * S) * p.Bpad + b;
#pragma unroll
      for (int i = 0; i < 8 * NCH; ++i)
        if (i < S) tp[(long long)i * p.Bpad] = x[i];
    }
    // u~ slots of this block (slot t = step * I + component sits at column 62 - t): gathered, then ONE 16-column store of
    // columns 48..63 (state columns below S keep their value, column 63 is the constant 1)
    float uval[CC_LIN_MAXM];
    const int cB = (k0 + nst - 1) >> 5;
    if (stg) {  // the chunks of this block must have landed
#pragma unroll 1
      for (; issued <= cB; ++issued) cc_lin_tile_load(tin + (issued & 1) * CC_LIN_TILE, p.in, b0, nact, T, T, issued, lane);
      if (waited <= cB) {
        cc_lin_cp_wait_all();
        __syncwarp();
        waited = issued;
      }
    }
#pragma unroll
    for (int t = 0; t < CC_LIN_MAXM; ++t) {
      uval[t] = 0.f;
      if (t < nst * I) {
        const int k = k0 + t;  // (staged path: I = 1, slot t = step t)
        if (stg) {
          uval[t] = cc_ut(M.ut, M.u_scale, tin[((k >> 5) & 1) * CC_LIN_TILE + lane * 33 + (k & 31)]);
        } else if (act) {
          uval[t] = cc_ut(M.ut, M.u_scale, __ldg(p.in + b * (long long)T * I + (long long)k0 * I + t));
        }
      }
    }
    if (stg && issued == cB + 1 && 32 * (cB + 1) < T && ((k0 + nst) >> 5) == cB) {
      // keep one chunk of copies in flight: chunk cB + 1 goes into the slot of chunk cB - 1, which every lane has finished reading
      // (only once the block lies inside chunk cB; a block that ends on the boundary loads the next chunk on demand)
      __syncwarp();
      cc_lin_tile_load(tin + (issued & 1) * CC_LIN_TILE, p.in, b0, nact, T, T, issued, lane);
      ++issued;
    }
    {
      float arr[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = 48 + i, t = 62 - c;
        float v = c == CC_LIN_ONE ? 1.f : 0.f;
        if (t >= 0 && t < CC_LIN_MAXM) v = uval[t < 0 ? 0 : (t < CC_LIN_MAXM ? t : 0)];
        if constexpr (NCH >= 7) {
          if (c < 8 * NCH) v = c < S ? x[c < 8 * NCH ? c : 0] : v;
        }
        arr[i] = v;
      }
      cc_lin_store_split<16>(tm, 48, arr);
    }
    cc_lin_round(tm);
    cc_lin_await(tm);
    {
      float d[16];
      cc_lin_ld<16>(tm, CC_LIN_COL_D + 48, d);
      cc_lin_wait_ld();
#pragma unroll
      for (int t = 0; t < CC_LIN_MAXM; ++t) {
        if (t < nst * O) {
          const float y = d[14 - t];
          if (stg) {
            const int k = k0 + t;  // O = 1: slot t = step t
            tout[((k >> 5) & 1) * CC_LIN_TILE + lane * 33 + (k & 31)] = y;
          } else if (act) {
            p.out[b * (long long)(T + 1) * O + (long long)(k0 + 1) * O + t] = y;
          }
        }
      }
    }
    if (stg) {  // an output chunk is complete when the block passes its last step (or the end): one coalesced tile store
      const int kl = k0 + nst - 1, c0 = k0 >> 5, c1 = kl >> 5;
      const bool end1 = (kl & 31) == 31 || kl == T - 1;
      const bool flush0 = c1 != c0 || end1, flush1 = c1 != c0 && end1;
      if (flush0) {
        __syncwarp();
        cc_lin_tile_store(tout + (c0 & 1) * CC_LIN_TILE, p.out + 1, b0, nact, T + 1, T, c0, lane);  // (+1: row k+1; row stride T+1)
        if (flush1) cc_lin_tile_store(tout + (c1 & 1) * CC_LIN_TILE, p.out + 1, b0, nact, T + 1, T, c1, lane);
        __syncwarp();
      }
    }
    if (k0 + m >= T && p.x_final && p.lin_mt > 0) {
      // ragged last block and the final state is requested: the round above advanced the chain by m steps (the phantom
      // steps saw u~ = 0), which is right for the outputs but ends BEHIND step T.  Same operand, the block matrices of
      // length T mod m (second prep pack): x_T = x_k0 + D'.  The first batch has completed (await), so B is free.
      const float4* src_hi = reinterpret_cast<const float4*>(p.lin_pack + CC_LIN_PACK_FLOATS + CC_LIN_WF_HI);
      const float4* src_lo = reinterpret_cast<const float4*>(p.lin_pack + CC_LIN_PACK_FLOATS + CC_LIN_WF_LO);
      float4* dhi = reinterpret_cast<float4*>(sh.bhi);
      float4* dlo = reinterpret_cast<float4*>(sh.blo);
      for (int i = tid; i < 1024; i += CC_LIN_THREADS) { dhi[i] = src_hi[i]; dlo[i] = src_lo[i]; }
      cc_lin_fence_async();
      cc_lin_round(tm);
      cc_lin_await(tm);
    }
    cc_lin_advance_state<NCH>(tm, x, S);
  }
  cc_lin_cp_wait_all();
  if (p.x_final && act)
#pragma unroll
    for (int i = 0; i < 8 * NCH; ++i)
      if (i < S) p.x_final[b * S + i] = x[i];
  cc_lin_epilogue(tm);
}

// ------------------------------------------------------------------------------------------------
// K2: open-loop reverse pass of a trainable linear model
// ------------------------------------------------------------------------------------------------
#define CC_LINW_TMEM_COLS 512
#define CC_LINW_COL_Q1 256   // Q accumulator, B = X_hi
#define CC_LINW_COL_Q2 320   // Q accumulator, B = X_lo
#define CC_LINW_OPA 8448     // float offset of operand A (Lam hi rows 0..63, lo rows 64..127): 16 row groups x 1152 floats
#define CC_LINW_OPB (CC_LINW_OPA + 16 * 1152)   // operand B (X hi rows 0..63, lo rows 64..127)
#define CC_LINW_STG (CC_LINW_OPB + 16 * 1152 + 64)   // cp.async staging of the NEXT block's loads: [x_k (64) ; ybar slots (13) ; u slots (13)][128 threads]
#define CC_LINW_SMEM_FLOATS (CC_LINW_STG + 90 * 128)
// float offset of operand row r (the thread's trajectory part is added separately): core matrix = 8 rows x 4 trajectories
// (128 B) padded to 144 B so that the 32 four-byte stores of a warp hit 32 banks
CC_DEV constexpr int cc_linw_row(int r) { return (r >> 3) * 1152 + (r & 7) * 4; }

struct CcLinwSync {
#ifndef CC_EMULATE
  uint64_t* q_done;
  uint32_t qpar;
  uint32_t opa, opb;  // shared-memory addresses
#else
  float* opa; float* opb; float* tmq;  // emulated Q accumulators [128 rows][128 cols]
#endif
};

// one round of the reverse kernel: the main TS-mode batch (d_ready) and, if with_q, the SS-mode Q batch (q_done)
CC_DEV void cc_linw_round(CcLinTm& t, CcLinwSync& s, bool with_q) {
#ifdef CC_EMULATE
  cc_lin_round(t);
  if (with_q && t.tid == 0) {
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < 128; ++c) {
        float acc = 0.f;
        for (int tr = 0; tr < 128; ++tr) {
          const int to = (tr >> 2) * 36 + (tr & 3);
          acc += cc_lin_tf32(s.opa[cc_linw_row(r) + to]) * cc_lin_tf32(s.opb[cc_linw_row(c) + to]);
        }
        s.tmq[r * 128 + c] = acc;
      }
  }
  __syncthreads();
#else
  cc_tmem_wait_st();
  cc_fence_proxy_async();  // the operand stores (generic proxy) must be visible to the MMAs (async proxy)
  cc_tc_fence_before();
  cc_mbar_arrive(t.a_ready);
  const uint32_t batch = t.batch++;
  if (t.warp == (int)(batch & 3u)) {
    cc_mbar_wait(t.a_ready, batch & 1u);
    cc_tc_fence_after();
    if (t.lane == 0) {
      cc_lin_issue<0, 8>(t.tbase, t.bhi, t.blo);
      cc_tc_commit(t.d_ready);
      if (with_q) {
        constexpr uint32_t idw = cc_umma_idesc_tf32(128, 64);
        const uint64_t da = cc_umma_desc_kmajor(s.opa, 144, 4608), db = cc_umma_desc_kmajor(s.opb, 144, 4608);
#pragma unroll
        for (int js = 0; js < 16; ++js) {  // 8 trajectories (two core matrices) per instruction
          const uint64_t ad = da + (uint64_t)((js * 288u) >> 4);
          cc_umma_tf32_ss(t.tbase + CC_LINW_COL_Q1, ad, db + (uint64_t)((js * 288u) >> 4), idw, js ? 1u : 0u);
          cc_umma_tf32_ss(t.tbase + CC_LINW_COL_Q2, ad, db + (uint64_t)((8u * 4608u + js * 288u) >> 4), idw, js ? 1u : 0u);
        }
      }
      cc_tc_commit(s.q_done);
    }
    __syncwarp();
  }
#endif
}
CC_DEV void cc_linw_await_q(CcLinwSync& s) {
#ifndef CC_EMULATE
  cc_mbar_wait(s.q_done, s.qpar);
  s.qpar ^= 1;
  cc_tc_fence_after();
#else
  (void)s;
#endif
}

__global__ void __launch_bounds__(CC_LIN_THREADS, 1) cc_linw_bwd_kernel(const CC_GRID_CONSTANT CcKParams gp) {
  CC_DIRECT_PARAMS(gp)
  (void)cta_smem;
  constexpr int NCH = 8;
  const CcNodeDev& M = p.nd[0];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const CcLinShared sh = cc_lin_carve(cc_smem);
  CcLinTm tm;
  CcLinwSync qs;
  float* opA = cc_smem + CC_LINW_OPA;
#ifndef CC_EMULATE
  opA += ((128u - (cc_smem_u32(opA) & 127u)) & 127u) / 4;  // operand buffers on a 128-byte boundary
#endif
  float* opB = opA + 16 * 1152;
  // (the prologue allocates 256 TMEM columns for the lin kernels; this kernel needs 512: own allocation below)
  const int S = M.S, I = M.I, O = M.O, m = p.lin_m, T = p.T;
  {
    const float* pack = p.lin_pack;
    const float4* src_hi = reinterpret_cast<const float4*>(pack + CC_LIN_WB_HI);
    const float4* src_lo = reinterpret_cast<const float4*>(pack + CC_LIN_WB_LO);
    float4* dhi = reinterpret_cast<float4*>(sh.bhi);
    float4* dlo = reinterpret_cast<float4*>(sh.blo);
    for (int i = tid; i < 1024; i += blockDim.x) { dhi[i] = src_hi[i]; dlo[i] = src_lo[i]; }
    for (int i = tid; i < 32 * 1152; i += blockDim.x) opA[i] = 0.f;  // both operands incl. padding
  }
#ifdef CC_EMULATE
  tm.tm = CC_SMEM_PTR + CC_LINW_SMEM_FLOATS;
  tm.bhi = sh.bhi; tm.blo = sh.blo; tm.tid = tid;
  for (int c = 0; c < 192; ++c) tm.tm[c * 128 + tid] = 0.f;
  qs.opa = opA; qs.opb = opB; qs.tmq = tm.tm + 192 * 128;
  __syncthreads();
#else
  uint64_t* q_done = sh.a_ready + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sh.a_ready + 3);
  if (warp == 0) {
    if (lane == 0) {
      cc_mbar_init(sh.a_ready, 128);
      cc_mbar_init(sh.d_ready, 1);
      cc_mbar_init(q_done, 1);
      cc_mbar_fence_init();
    }
    __syncwarp();
    cc_tmem_alloc(tmem_slot, CC_LINW_TMEM_COLS);
  }
  cc_fence_proxy_async();
  cc_tc_fence_before();
  __syncthreads();
  cc_tc_fence_after();
  tm.tbase = *tmem_slot;
  tm.lane_base = tm.tbase + ((uint32_t)(warp * 32) << 16);
  tm.bhi = cc_smem_u32(sh.bhi); tm.blo = cc_smem_u32(sh.blo);
  tm.a_ready = sh.a_ready; tm.d_ready = sh.d_ready;
  tm.batch = 0; tm.par = 0; tm.warp = warp; tm.lane = lane;
  qs.q_done = q_done; qs.qpar = 0; qs.opa = cc_smem_u32(opA); qs.opb = cc_smem_u32(opB);
  {
    float z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0.f;
    cc_lin_st<32>(tm, CC_LIN_COL_AHI, z); cc_lin_st<32>(tm, CC_LIN_COL_AHI + 32, z);
    cc_lin_st<32>(tm, CC_LIN_COL_ALO, z); cc_lin_st<32>(tm, CC_LIN_COL_ALO + 32, z);
  }
#endif
  const long long b0 = (long long)blockIdx.x * 128 + warp * 32;
  const long long rem = p.B - b0;
  const int nact = rem >= 32 ? 32 : (rem > 0 ? (int)rem : 0);
  const bool act = lane < nact;
  const long long b = b0 + lane;
  const int tho = (tid >> 2) * 36 + (tid & 3);  // this thread's (trajectory's) float offset inside an operand row group
  float* oa = opA + tho;
  float* ob = opB + tho;
  const long long yrow = (long long)(T + 1) * O;
  // ---- y0 = g(x0): its cotangent reaches only g (x0 is not trained): per-warp sums of ybar_0 (x) [x0 ; 1] ----
  {
    float* g0 = p.lin_q + (long long)gridDim.x * 128 * 64 + ((long long)blockIdx.x * 4 + warp) * 4 * 64;  // [cta][warp][O][64]
    for (int o = 0; o < O; ++o) {
      const float yb0 = act ? __ldg(p.ybar + b * yrow + o) : 0.f;
      for (int s = 0; s < 64; ++s) {
        float v = 0.f;
        if (s < S) v = (p.x0 && act) ? yb0 * __ldg(p.x0 + b * S + s) : 0.f;
        else if (s == CC_LIN_ONE) v = yb0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(CC_FULL, v, off);
        if (lane == 0) g0[o * 64 + s] = v;
      }
    }
  }
  float mu[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) mu[i] = 0.f;
  float qsum[64];  // running FP32 sums of this thread's row of Q (row = TMEM lane: Lam index n, hi part for lanes < 64, lo part above)
#pragma unroll
  for (int i = 0; i < 64; ++i) qsum[i] = 0.f;
  const int nblk = (T + m - 1) / m;
  float* stg_ = cc_smem + CC_LINW_STG + tid;  // (behind the operand buffers incl. their alignment shift of < 32 floats)
  auto prefetch_blk = [&](int bk) {
    if (bk >= 0 && act) {
      const int kb = bk * m;
      const int ns = T - kb < m ? T - kb : m;
      const float* tp = p.tape + ((long long)bk * S) * p.Bpad + b;
      for (int i = 0; i < S; ++i) cc_lin_cp4(stg_ + i * 128, tp + (long long)i * p.Bpad);
      for (int t = 0; t < ns * O; ++t) cc_lin_cp4(stg_ + (64 + t) * 128, p.ybar + b * yrow + (long long)(kb + 1) * O + t);  // row k+1 <-> step k
      for (int t = 0; t < ns * I; ++t) cc_lin_cp4(stg_ + (77 + t) * 128, p.in + b * (long long)T * I + (long long)kb * I + t);
    }
    cc_lin_cp_commit();
  };
  prefetch_blk(nblk - 1);
  bool pending = false;  // a Q batch is in flight / its result not yet flushed
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int k0 = blk * m;
    const int nst = T - k0 < m ? T - k0 : m;
    // ---- gather this block's operands: Lam = [mu_E ; w slots], X = [x_k ; u~ slots ; 1] ----
    // this block's loads were issued (cp.async) while the previous block was processed: global latency is covered
    cc_lin_cp_wait_all();
    float wv[CC_LIN_MAXM], uv[CC_LIN_MAXM], uraw[CC_LIN_MAXM];
#pragma unroll
    for (int t = 0; t < CC_LIN_MAXM; ++t) {
      wv[t] = (t < nst * O && act) ? stg_[(64 + t) * 128] : 0.f;
      uraw[t] = (t < nst * I && act) ? stg_[(77 + t) * 128] : 0.f;
      uv[t] = (t < nst * I && act) ? cc_ut(M.ut, M.u_scale, uraw[t]) : 0.f;
    }
    // the previous block's Q batch must have read the operand buffers before they are rewritten; flush its result
    if (pending) {
      cc_linw_await_q(qs);
#ifdef CC_EMULATE
      for (int c = 0; c < 64; ++c) qsum[c] += qs.tmq[tid * 128 + c] + qs.tmq[tid * 128 + 64 + c];
#else
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        float q1[16], q2[16];
        cc_lin_ld<16>(tm, CC_LINW_COL_Q1 + c, q1);
        cc_lin_ld<16>(tm, CC_LINW_COL_Q2 + c, q2);
        cc_lin_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) qsum[c + i] += q1[i] + q2[i];
      }
#endif
    }
    // TMEM A operand: state chunks = mu (already there from the previous advance, zeros for the first block); the w slots
    // with ONE 16-column store of columns 48..63 (state columns below S keep mu)
    {
      float arr[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = 48 + i, t = 62 - c;
        float v = 0.f;
        if (t >= 0 && t < CC_LIN_MAXM) v = wv[t < 0 ? 0 : (t < CC_LIN_MAXM ? t : 0)];
        arr[i] = c < S ? mu[c] : v;
      }
      cc_lin_store_split<16>(tm, 48, arr);
    }
    // shared-memory operands (K = trajectory): Lam rows / X rows, hi and lo parts; x_k streams from the tape (block-start state)
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      float lamv = mu[i], xv = 0.f;
      if (i < 50) {
        if (i < S && act) xv = stg_[i * 128];
      } else if (i <= 62) {  // slot columns
        lamv = i < S ? lamv : wv[62 - i];
        xv = i < S ? (act ? stg_[i * 128] : 0.f) : uv[62 - i];
      } else {
        lamv = 0.f; xv = 1.f;
      }
      const float lh = cc_lin_tf32(lamv), xh = cc_lin_tf32(xv);
      oa[cc_linw_row(i)] = lh; oa[cc_linw_row(64 + i)] = lamv - lh;
      ob[cc_linw_row(i)] = xh; ob[cc_linw_row(64 + i)] = xv - xh;
    }
    prefetch_blk(blk - 1);  // (the staging buffer has been consumed)
    cc_linw_round(tm, qs, true);
    pending = true;
    cc_lin_await(tm);
    // ubar of the block's steps
    if (p.ubar) {
      float d[16];
      cc_lin_ld<16>(tm, CC_LIN_COL_D + 48, d);
      cc_lin_wait_ld();
#pragma unroll
      for (int t = 0; t < CC_LIN_MAXM; ++t)
        if (t < nst * I && act) p.ubar[b * (long long)T * I + (long long)k0 * I + t] = d[14 - t] * cc_ut_grad(M.ut, M.u_scale, uraw[t]);
    }
    cc_lin_advance_state<NCH>(tm, mu, S);
  }
  if (pending) {
    cc_linw_await_q(qs);
#ifdef CC_EMULATE
    for (int c = 0; c < 64; ++c) qsum[c] += qs.tmq[tid * 128 + c] + qs.tmq[tid * 128 + 64 + c];
#else
#pragma unroll
    for (int c = 0; c < 64; c += 16) {
      float q1[16], q2[16];
      cc_lin_ld<16>(tm, CC_LINW_COL_Q1 + c, q1);
      cc_lin_ld<16>(tm, CC_LINW_COL_Q2 + c, q2);
      cc_lin_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) qsum[c + i] += q1[i] + q2[i];
    }
#endif
  }
  // per-CTA partial of Q: [cta][row = 128 (hi rows then lo rows)][64]
  {
    float* qp = p.lin_q + ((long long)blockIdx.x * 128 + tid) * 64;
#pragma unroll
    for (int c = 0; c < 64; c += 4) *reinterpret_cast<float4*>(qp + c) = make_float4(qsum[c], qsum[c + 1], qsum[c + 2], qsum[c + 3]);
  }
#ifndef CC_EMULATE
  cc_tc_fence_before();
  __syncthreads();
  if (warp == 0) cc_tmem_dealloc(tm.tbase, CC_LINW_TMEM_COLS);
#endif
}

// ------------------------------------------------------------------------------------------------
// fp64 post-processing  Q -> theta_bar
// ------------------------------------------------------------------------------------------------
#define CC_LIN_GRAD_SMEM_DOUBLES (4 * 4096)
#define CC_LIN_GRADF_SMEM_DOUBLES (6 * 4096 + 4 * 64 + 64 + 8)

// Q[r][c] = sum over CTAs (fixed order: deterministic) of the hi-row and lo-row partials, in double; grid = 64 (rows)
__global__ void __launch_bounds__(64) cc_lin_qreduce_kernel(const CC_GRID_CONSTANT CcLinGradArgs a) {
  const int r = blockIdx.x, c = threadIdx.x;
  double s = 0.0;
  for (int t = 0; t < a.n_cta; ++t) s += (double)a.qpart[((long long)t * 128 + r) * 64 + c] + (double)a.qpart[((long long)t * 128 + 64 + r) * 64 + c];
  a.work[r * 64 + c] = s;
  if (r < 4) {  // ybar_0 (x) [x0 ; 1] partials: [cta][warp][o][64]
    double g = 0.0;
    const float* g0 = a.qpart + (long long)a.n_cta * 128 * 64;
    for (int t = 0; t < a.n_cta * 4; ++t) g += (double)g0[((long long)t * 4 + r) * 64 + c];
    a.work[4096 + r * 64 + c] = g;
  }
}

// CTA j: contributions of in-block step j:  T1 = L_j Q ;  Mx_j = T1 R_j^T ; Mu_j, M1_j = columns of T1 ;
// Gb_j = Q[w-slot rows of step j] R_{j+q}^T ; db_j = Q[w-slot rows][one]
__global__ void __launch_bounds__(CC_LIN_PREP_THREADS, 1) cc_lin_grad_kernel(const CC_GRID_CONSTANT CcLinGradArgs a) {
  double* sm = CC_LIN_DSMEM;
  double* Q = sm;            // [64][64]
  double* Lj = Q + 4096;     // [S][64]  (Lam columns)
  double* Rj = Lj + 4096;    // [S][64]  (X columns)
  double* T1 = Rj + 4096;    // [S][64]
  const CcNodeDev& M = a.M;
  const int S = M.S, I = M.I, O = M.O, m = a.m, q = M.pre ? 0 : 1;
  const int j = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  const double* PW = a.tab + CC_LIN_TAB_PW;
  const double* RG = a.tab + CC_LIN_TAB_RG(m);
  const double* CU = a.tab + CC_LIN_TAB_CU(m);
  const double* GS = a.tab + CC_LIN_TAB_GS(m);
  auto Lentry = [&](int jj, int s, int c) -> double {  // L_jj[s][Lam column c]
    if (c < S) return PW[(size_t)(m - jj - 1) * 4096 + c * 64 + s];  // ((Phi^T)^{m-jj-1})[s][c] = Phi^{..}[c][s]
    const int t = 62 - c;
    if (t >= 0 && t < m * O) {
      const int i = t / O, o = t % O;
      if (i >= jj + 1 - q) return RG[((i - jj - 1 + q) * O + o) * 64 + s];
    }
    return 0.0;
  };
  auto Rentry = [&](int jj, int s, int c) -> double {  // R_jj[s][X column c], jj = 0..m
    if (c < S) return PW[(size_t)jj * 4096 + s * 64 + c];
    if (c == CC_LIN_ONE) return GS[jj * 64 + s];
    const int t = 62 - c;
    if (t >= 0 && t < m * I) {
      const int i = t / I, ii = t % I;
      if (i < jj) return CU[((jj - 1 - i) * I + ii) * 64 + s];
    }
    return 0.0;
  };
  for (int e = tid; e < 4096; e += nthr) {
    const int s = e / 64, c = e % 64;
    Q[e] = a.work[e];
    Lj[e] = s < S ? Lentry(j, s, c) : 0.0;
    Rj[e] = s < S ? Rentry(j, s, c) : 0.0;
  }
  __syncthreads();
  for (int e = tid; e < S * 64; e += nthr) {
    const int s = e / 64, c = e % 64;
    double acc = 0.0;
    for (int r = 0; r < 64; ++r) acc += Lj[s * 64 + r] * Q[r * 64 + c];
    T1[s * 64 + c] = acc;
  }
  __syncthreads();
  double* out = a.work + 4096 + 4 * 64 + (size_t)j * CC_LIN_GRAD_PER_J;
  for (int e = tid; e < 4096; e += nthr) {
    const int s = e / 64, s2 = e % 64;
    double acc = 0.0;
    if (s < S && s2 < S)
      for (int c = 0; c < 64; ++c) acc += T1[s * 64 + c] * Rj[s2 * 64 + c];
    out[e] = acc;
  }
  for (int e = tid; e < 256; e += nthr) {
    const int s = e / 4, ii = e % 4;
    out[4096 + e] = (s < S && ii < I) ? T1[s * 64 + CC_LIN_SLOT(j * I + ii)] : 0.0;
  }
  for (int s = tid; s < 64; s += nthr) out[4096 + 256 + s] = s < S ? T1[s * 64 + CC_LIN_ONE] : 0.0;
  // readout: rows of Q of the step's w slots against R_{j+q}
  for (int e = tid; e < 256; e += nthr) {
    const int o = e / 64, s2 = e % 64;
    double acc = 0.0;
    if (o < O && s2 < S) {
      const int row = CC_LIN_SLOT(j * O + o);
      for (int c = 0; c < 64; ++c) acc += Q[row * 64 + c] * Rentry(j + q, s2, c);
    }
    out[4096 + 256 + 64 + e] = acc;
  }
  if (tid < 4) out[4096 + 256 + 64 + 256 + tid] = tid < O ? Q[CC_LIN_SLOT(j * O + tid) * 64 + CC_LIN_ONE] : 0.0;
}

// final: sum over j, derivative of the matrix polynomials Phi(A), Gam(A) (tests/linblock_ref.py: matpoly_grad), theta_bar
__global__ void __launch_bounds__(CC_LIN_PREP_THREADS, 1) cc_lin_grad_final_kernel(const CC_GRID_CONSTANT CcLinGradArgs a) {
  double* sm = CC_LIN_DSMEM;
  double* Mx = sm;            // later: Abar
  double* At = Mx + 4096;     // h * A^T
  double* U = At + 4096;      // scratch
  double* V = U + 4096;
  double* W = V + 4096;
  double* Mb = W + 4096;      // M_beta = Mu B^T + M1 c^T
  double* Mu = Mb + 4096;     // [64][4]
  double* M1 = Mu + 256;      // [64]
  const CcNodeDev& M = a.M;
  const CcLayerDev& F = M.f.layer[0];
  const CcLayerDev& G = M.g.layer[0];
  const int S = M.S, I = M.I, O = M.O, m = a.m;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const double h = (double)M.dt, os = (double)M.out_scale;
  const double* tA = a.tab + CC_LIN_TAB_A(m);
  const double* tG = a.tab + CC_LIN_TAB_GAM(m);
  const double* part = a.work + 4096 + 4 * 64;
  for (int e = tid; e < 4096; e += nthr) {
    const int i = e / 64, jx = e % 64;
    double s = 0.0;
    for (int j = 0; j < m; ++j) s += part[(size_t)j * CC_LIN_GRAD_PER_J + e];
    Mx[e] = s;
    At[e] = h * tA[jx * 64 + i];
    U[e] = 0.0; V[e] = 0.0; W[e] = 0.0;
  }
  for (int e = tid; e < 256 + 64; e += nthr) {
    double s = 0.0;
    for (int j = 0; j < m; ++j) s += part[(size_t)j * CC_LIN_GRAD_PER_J + 4096 + e];
    Mu[e] = s;  // (M1 follows Mu)
  }
  __syncthreads();
  // readout gradient: g weight / bias (flat order of theta)
  for (int e = tid; e < 256 + 4; e += nthr) {
    double s = 0.0;
    for (int j = 0; j < m; ++j) s += part[(size_t)j * CC_LIN_GRAD_PER_J + 4096 + 256 + 64 + e];
    if (e < 256) {
      const int o = e / 64, s2 = e % 64;
      if (o < O && s2 < S) a.theta_bar[G.w_off + o * G.in_log + s2] = (float)(os * (s + a.work[4096 + o * 64 + s2]));
    } else {
      const int o = e - 256;
      if (o < O && G.b_off >= 0) a.theta_bar[G.b_off + o] = (float)(os * (s + a.work[4096 + o * 64 + CC_LIN_ONE]));
    }
  }
  // M_beta = Mu B^T + M1 c^T
  for (int e = tid; e < 4096; e += nthr) {
    const int s = e / 64, k = e % 64;
    double acc = 0.0;
    if (s < S && k < S) {
      for (int ii = 0; ii < I; ++ii) acc += Mu[s * 4 + ii] * (double)cc_weight_at(M, F, a.theta, k, M.row_v + ii);
      acc += M1[s] * (double)cc_weight_at(M, F, a.theta, k, M.row_one);
    }
    Mb[e] = acc;
  }
  __syncthreads();
  // Abar = sum_n coef_n sum_{a+b=n-1} (A^T)^a Mx (A^T)^b  (+ the same with the Gam coefficients on M_beta), in powers of X = hA^T:
  //   Abar = (1/h) * [ sum_{a,b} c_{a+b+1} X^a Mx X^b ] * h  ... written with coef on X^n directly: coefficient of X^a M X^b is
  //   phi_{a+b+1} * h  for Phi (d(hA)^n = h * ...)  and  h * phi_{a+b+2} * h for Gam.
  auto poly_apply = [&](const double* Min, int nmax, const double* cf, double* Out) {
    // Out += sum_{a+b <= nmax-1} cf[a+b+1] X^a Min X^b   via  U_a = X^a Min and Horner in X on the right
    // W_b = sum_a cf[a+b+1] U_a ; result = W_0 + (W_1 + (W_2 + W_3 X) X) X
    // computed as: for b from nmax-1 down to 0: V = W_b + V X, where W_b needs all U_a with a <= nmax-1-b
    for (int e = tid; e < 4096; e += nthr) V[e] = 0.0;
    __syncthreads();
    for (int bb = nmax - 1; bb >= 0; --bb) {
      // V <- V X
      if (bb != nmax - 1) {
        cc_lin_matmul(W, V, At, S, tid, nthr);
        for (int e = tid; e < 4096; e += nthr) V[e] = W[e];
        __syncthreads();
      }
      // V += sum_{a <= nmax-1-bb} cf[a+bb+1] X^a Min : U runs through X^a Min
      for (int e = tid; e < 4096; e += nthr) U[e] = Min[e];
      __syncthreads();
      for (int aa = 0; aa <= nmax - 1 - bb; ++aa) {
        if (aa > 0) {
          cc_lin_matmul(W, At, U, S, tid, nthr);
          for (int e = tid; e < 4096; e += nthr) U[e] = W[e];
          __syncthreads();
        }
        const double c = cf[aa + bb + 1];
        for (int e = tid; e < 4096; e += nthr) V[e] += c * U[e];
        __syncthreads();
      }
    }
    for (int e = tid; e < 4096; e += nthr) Out[e] += V[e];
    __syncthreads();
  };
  double* Abar = a.work;  // reuse the Q slot of the work buffer (Q is no longer needed)
  for (int e = tid; e < 4096; e += nthr) Abar[e] = 0.0;
  __syncthreads();
  if (M.integrator == CC_INT_RK4) {
    // d/dA of Phi = sum phi_n (hA)^n: coefficient of (hA^T)^a M (hA^T)^b is phi_{a+b+1} * h
    const double cphi[5] = {0.0, h * 1.0, h * 0.5, h / 6.0, h / 24.0};
    poly_apply(Mx, 4, cphi, Abar);
    // Gam = h sum_{n<=3} phi_{n+1} (hA)^n
    const double cgam[4] = {0.0, h * h * 0.5, h * h / 6.0, h * h / 24.0};
    poly_apply(Mb, 3, cgam, Abar);
  } else if (M.integrator == CC_INT_EE) {
    for (int e = tid; e < 4096; e += nthr) Abar[e] = h * Mx[e];
    __syncthreads();
  } else {
    for (int e = tid; e < 4096; e += nthr) Abar[e] = Mx[e];
    __syncthreads();
  }
  // f weight [A | B] and bias c:  Bbar = Gam^T Mu, cbar = Gam^T M1
  for (int e = tid; e < S * (S + I + 1); e += nthr) {
    const int n = e / (S + I + 1), k = e % (S + I + 1);
    if (k < S) {
      a.theta_bar[F.w_off + n * F.in_log + k] = (float)Abar[n * 64 + k];
    } else {
      double acc = 0.0;
      const bool isb = k < S + I;
      for (int s = 0; s < S; ++s) acc += tG[s * 64 + n] * (isb ? Mu[s * 4 + (k - S)] : M1[s]);
      if (isb) a.theta_bar[F.w_off + n * F.in_log + k] = (float)acc;
      else if (F.b_off >= 0) a.theta_bar[F.b_off + n] = (float)acc;
    }
  }
}
