// cc_lin_tm.cuh -- the tensor-memory / MMA layer of the blocked linear kernels behind one small interface, so that the
// thread-level code of cc_lin_kernels.cuh is the same source for sm_100a (tcgen05 + TMEM + mbarriers, cc_tmem.cuh) and
// for the CPU emulation of tests/emu (TMEM = an array in the emulated shared memory, the MMA batch = a loop executed by
// thread 0 between two barriers, with the same tf32 truncation of the operands: the emulation also predicts the
// rounding error of the 3xTF32 products).
//
// One CTA = 128 threads = 128 trajectories = the 128 TMEM lanes.  Per CTA: D[64] (fp32 accumulator), A_hi[64], A_lo[64]
// columns; B = the block matrix (hi / lo) in shared memory.  One "round": every thread writes its A row, publish();
// the MMA batch D = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi runs (3 x 8 tcgen05.mma, M = 128, N = 64, K = 8); await(); every
// thread reads its D row.
#pragma once
#include "cc_kernels.cuh"
#include "cc_lin.h"
#ifndef CC_EMULATE
#include "cc_tmem.cuh"
#endif

#define CC_LIN_THREADS 128
#define CC_LIN_TMEM_COLS 256
#define CC_LIN_COL_D 0
#define CC_LIN_COL_AHI 64
#define CC_LIN_COL_ALO 128

struct CcLinTm {
#ifdef CC_EMULATE
  float* tm;        // [192 columns][128 lanes]
  const float* bhi; // B operand (MMA layout) in emulated shared memory
  const float* blo;
  int tid;
#else
  uint32_t lane_base;  // TMEM address of this thread's lane, column 0
  uint32_t tbase;      // TMEM base of the CTA
  uint32_t bhi, blo;   // shared-memory addresses of the B operand
  uint64_t* a_ready;
  uint64_t* d_ready;
  uint32_t batch;      // rounds published so far
  uint32_t par;        // d_ready parity
  int warp, lane;
#endif
};

CC_DEV float cc_lin_tf32(float v) {  // truncation to tf32 (what the tensor core does with a 32-bit operand)
  unsigned u;
  memcpy(&u, &v, 4);
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}
CC_DEV float cc_lin_tf32_rn(float v) {  // round to nearest tf32 (ties away): the operand splits below are unbiased
  unsigned u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

#ifdef CC_EMULATE
template <int NC>
CC_DEV void cc_lin_ld(const CcLinTm& t, int col, float (&r)[NC]) {
  for (int i = 0; i < NC; ++i) r[i] = t.tm[(col + i) * 128 + t.tid];
}
template <int NC>
CC_DEV void cc_lin_st(const CcLinTm& t, int col, const float (&r)[NC]) {
  for (int i = 0; i < NC; ++i) t.tm[(col + i) * 128 + t.tid] = r[i];
}
CC_DEV float cc_lin_ld1(const CcLinTm& t, int col) { return t.tm[col * 128 + t.tid]; }
CC_DEV void cc_lin_st1(const CcLinTm& t, int col, float v) { t.tm[col * 128 + t.tid] = v; }
CC_DEV void cc_lin_wait_ld() {}
CC_DEV void cc_lin_wait_st() {}
// chunks [c0, c1) of the contraction (8 columns each)
CC_DEV void cc_lin_round(CcLinTm& t, int c0 = 0, int c1 = 8) {
  __syncthreads();
  if (t.tid == 0) {
    for (int l = 0; l < 128; ++l)
      for (int n = 0; n < 64; ++n) {
        float acc = 0.f;
        for (int pass = 0; pass < 3; ++pass)
          for (int k = 8 * c0; k < 8 * c1; ++k) {
            const int off = ((k >> 2) * 64 + n) * 4 + (k & 3);
            const float ahi = cc_lin_tf32(t.tm[(CC_LIN_COL_AHI + k) * 128 + l]), alo = cc_lin_tf32(t.tm[(CC_LIN_COL_ALO + k) * 128 + l]);
            const float bh = cc_lin_tf32(t.bhi[off]), bl = cc_lin_tf32(t.blo[off]);
            acc += pass == 0 ? alo * bh : (pass == 1 ? ahi * bl : ahi * bh);
          }
        t.tm[(CC_LIN_COL_D + n) * 128 + l] = acc;
      }
  }
  __syncthreads();
}
CC_DEV void cc_lin_await(CcLinTm&) {}
#else
template <int NC>
CC_DEV void cc_lin_ld(const CcLinTm& t, int col, float (&r)[NC]) {
  uint32_t u[NC];
  if constexpr (NC == 32) cc_tmem_ld32(t.lane_base + col, u);
  else if constexpr (NC == 16) cc_tmem_ld16(t.lane_base + col, u);
  else if constexpr (NC == 8) cc_tmem_ld8(t.lane_base + col, u);
  else cc_tmem_ld4(t.lane_base + col, u);
#pragma unroll
  for (int i = 0; i < NC; ++i) r[i] = __uint_as_float(u[i]);
}
template <int NC>
CC_DEV void cc_lin_st(const CcLinTm& t, int col, const float (&r)[NC]) {
  uint32_t u[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) u[i] = __float_as_uint(r[i]);
  if constexpr (NC == 32) cc_tmem_st32(t.lane_base + col, u);
  else if constexpr (NC == 16) cc_tmem_st16(t.lane_base + col, u);
  else if constexpr (NC == 8) cc_tmem_st8(t.lane_base + col, u);
  else cc_tmem_st4(t.lane_base + col, u);
}
CC_DEV float cc_lin_ld1(const CcLinTm& t, int col) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(t.lane_base + (uint32_t)col));
  return __uint_as_float(r);
}
CC_DEV void cc_lin_st1(const CcLinTm& t, int col, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(t.lane_base + (uint32_t)col), "r"(__float_as_uint(v)) : "memory");
}
CC_DEV void cc_lin_wait_ld() { cc_tmem_wait_ld(); }
CC_DEV void cc_lin_wait_st() { cc_tmem_wait_st(); }
// D = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi over the contraction chunks [C0, C1); issued by ONE thread
template <int C0, int C1>
CC_DEV void cc_lin_issue(uint32_t tbase, uint32_t bhi, uint32_t blo) {
  constexpr uint32_t kstep = 2u * 64u * 16u, lbo = 64u * 16u;
  constexpr uint32_t idesc = cc_umma_idesc_tf32(128, 64);
  const uint32_t aHi = tbase + CC_LIN_COL_AHI, aLo = tbase + CC_LIN_COL_ALO;
#pragma unroll
  for (int js = C0; js < C1; ++js) cc_umma_tf32_ts(tbase, aLo + 8 * js, cc_umma_desc_kmajor(bhi + js * kstep, lbo, 128), idesc, js > C0 ? 1u : 0u);
#pragma unroll
  for (int js = C0; js < C1; ++js) cc_umma_tf32_ts(tbase, aHi + 8 * js, cc_umma_desc_kmajor(blo + js * kstep, lbo, 128), idesc, 1u);
#pragma unroll
  for (int js = C0; js < C1; ++js) cc_umma_tf32_ts(tbase, aHi + 8 * js, cc_umma_desc_kmajor(bhi + js * kstep, lbo, 128), idesc, 1u);
}
// publish the A row just written and start the MMA batch: all 128 threads arrive; the issuing warp (rotating) waits for the
// last arrival and its lane 0 issues the batch and commits onto d_ready
CC_DEV void cc_lin_round(CcLinTm& t, int c0 = 0, int c1 = 8) {
  cc_tmem_wait_st();
  cc_tc_fence_before();
  cc_mbar_arrive(t.a_ready);
  const uint32_t batch = t.batch++;
  if (t.warp == (int)(batch & 3u)) {
    cc_mbar_wait(t.a_ready, batch & 1u);
    cc_tc_fence_after();
    if (t.lane == 0) {
      if (c0 == 0 && c1 == 4) cc_lin_issue<0, 4>(t.tbase, t.bhi, t.blo);
      else cc_lin_issue<0, 8>(t.tbase, t.bhi, t.blo);
      cc_tc_commit(t.d_ready);
    }
    __syncwarp();
  }
}
CC_DEV void cc_lin_await(CcLinTm& t) {
  cc_mbar_wait(t.d_ready, t.par);
  t.par ^= 1;
  cc_tc_fence_after();
}
#endif

// make shared-memory writes of this thread (generic proxy) visible to the tensor core (async proxy) before the next publish
CC_DEV void cc_lin_fence_async() {
#ifndef CC_EMULATE
  cc_fence_proxy_async();
#endif
}

// split NC values into tf32 hi / exact lo and store them at columns c0.. of A_hi / A_lo
template <int NC>
CC_DEV void cc_lin_store_split(const CcLinTm& t, int c0, const float (&z)[NC]) {
  float hi[NC], lo[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {  // z = hi + lo with |lo| <= 2^-12 |z|, lo rounded (not truncated) to tf32: error <= 2^-24 |z|
    hi[i] = cc_lin_tf32_rn(z[i]);
    lo[i] = cc_lin_tf32_rn(z[i] - hi[i]);
  }
  cc_lin_st<NC>(t, CC_LIN_COL_AHI + c0, hi);
  cc_lin_st<NC>(t, CC_LIN_COL_ALO + c0, lo);
}
CC_DEV void cc_lin_store_split1(const CcLinTm& t, int col, float v) {
  const float hi = cc_lin_tf32_rn(v);
  cc_lin_st1(t, CC_LIN_COL_AHI + col, hi);
  cc_lin_st1(t, CC_LIN_COL_ALO + col, cc_lin_tf32_rn(v - hi));
}
