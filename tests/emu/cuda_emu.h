// cuda_emu.h — TEST INFRASTRUCTURE.  A tiny CPU emulation of the CUDA execution model so that the
// device code in chain_control_b200/csrc can be compiled with g++ (-DCC_EMULATE) and its indexing /
// barrier logic checked against the oracle in this GPU-less container (with ASan for out-of-bounds).
// Every CUDA thread of a CTA is a ucontext fiber; __syncthreads() yields to a round-robin scheduler.
// Races between threads of one CTA are NOT modelled (fibers run one at a time, to the next barrier).
// The product library never includes this file: it is compiled only into tests/emu/libccb200_emu.so.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct uint3e { unsigned x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace ccemu {
struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  uint3e tid;
  bool done = false;
  int wait_kind = 0;  // 0 runnable, 1 waiting at __syncthreads, 2 waiting at a warp barrier
};
struct WarpState {
  int arrived = 0;
  unsigned gen = 0;
  float xchg[32];
};
struct State {
  ucontext_t sched;
  Fiber* cur = nullptr;
  uint3e bid{0, 0, 0}, bdim{1, 1, 1}, gdim{1, 1, 1};
  float* smem = nullptr;
  std::function<void()> body;
  std::vector<WarpState> warps;
  int cta_arrived = 0;
  unsigned cta_gen = 0;
  int alive = 0;
  long progress = 0;  // bumped whenever a thread passes a barrier or exits
};
inline State& st() {
  static thread_local State s;
  return s;
}
inline void trampoline() {
  State& s = st();
  s.body();
  s.cur->done = true;
  swapcontext(&s.cur->ctx, &s.sched);
}
template <class F>
inline void launch(int grid, int block, size_t smem_bytes, F f) {
  State& s = st();
  for (int b = 0; b < grid; ++b) {
    // exact-size heap allocation so that ASan sees out-of-bounds shared-memory accesses
    float* smem = static_cast<float*>(aligned_alloc(16, (smem_bytes + 15) / 16 * 16 + 16));
    memset(smem, 0xFF, smem_bytes);  // NaN-poison: uninitialised reads show up in the results
    s.smem = smem;
    s.bid = {unsigned(b), 0, 0};
    s.bdim = {unsigned(block), 1, 1};
    s.gdim = {unsigned(grid), 1, 1};
    s.body = f;
    std::vector<Fiber> fibers(block);
    for (int t = 0; t < block; ++t) {
      Fiber& fb = fibers[t];
      fb.stack.resize(256 * 1024);
      fb.tid = {unsigned(t), 0, 0};
      getcontext(&fb.ctx);
      fb.ctx.uc_stack.ss_sp = fb.stack.data();
      fb.ctx.uc_stack.ss_size = fb.stack.size();
      fb.ctx.uc_link = nullptr;
      makecontext(&fb.ctx, (void (*)())trampoline, 0);
    }
    s.warps.assign((block + 31) / 32, WarpState());
    s.cta_arrived = 0;
    s.alive = block;
    // round-robin: a fiber runs until it blocks at a barrier (it re-checks its condition itself)
    // or finishes.  Deadlock (nobody can progress) = barrier divergence in the kernel.
    long idle_rounds = 0;
    while (s.alive > 0) {
      const long before = s.progress;
      for (int t = 0; t < block; ++t) {
        Fiber& fb = fibers[t];
        if (fb.done) continue;
        s.cur = &fb;
        swapcontext(&s.sched, &fb.ctx);
        if (fb.done) { --s.alive; ++s.progress; }
      }
      idle_rounds = s.progress != before ? 0 : idle_rounds + 1;
      if (idle_rounds > 4) {
        fprintf(stderr, "ccemu: deadlock: %d threads wait at a barrier that can never complete\n", s.alive);
        abort();
      }
    }
    free(smem);
  }
}
}  // namespace ccemu

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define threadIdx (ccemu::st().cur->tid)
#define blockIdx (ccemu::st().bid)
#define blockDim (ccemu::st().bdim)
#define gridDim (ccemu::st().gdim)
static inline void __syncthreads() {
  ccemu::State& s = ccemu::st();
  ccemu::Fiber* me = s.cur;
  const unsigned gen = s.cta_gen;
  ++s.progress;
  if (++s.cta_arrived == (int)s.bdim.x) {  // every thread of the CTA must arrive (exited threads = bug)
    s.cta_arrived = 0;
    ++s.cta_gen;
    return;
  }
  me->wait_kind = 1;
  while (s.cta_gen == gen) swapcontext(&me->ctx, &s.sched);
  me->wait_kind = 0;
}
static inline int ccemu_warp_size_of(int warp) {
  const int n = (int)ccemu::st().bdim.x - warp * 32;
  return n < 32 ? n : 32;
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
  (void)mask;
  ccemu::State& s = ccemu::st();
  ccemu::Fiber* me = s.cur;
  const int w = me->tid.x / 32;
  ccemu::WarpState& ws = s.warps[w];
  const unsigned gen = ws.gen;
  ++s.progress;
  if (++ws.arrived == ccemu_warp_size_of(w)) {
    ws.arrived = 0;
    ++ws.gen;
    return;
  }
  me->wait_kind = 2;
  while (ws.gen == gen) swapcontext(&me->ctx, &s.sched);
  me->wait_kind = 0;
}
// full-mask shuffles (all lanes of the warp must participate)
static inline float __shfl_xor_sync(unsigned mask, float v, int lane_mask) {
  ccemu::State& s = ccemu::st();
  const int w = s.cur->tid.x / 32, l = s.cur->tid.x % 32;
  s.warps[w].xchg[l] = v;
  __syncwarp(mask);
  const float r = s.warps[w].xchg[l ^ lane_mask];
  __syncwarp(mask);
  return r;
}
static inline float __shfl_sync(unsigned mask, float v, int src) {
  ccemu::State& s = ccemu::st();
  const int w = s.cur->tid.x / 32, l = s.cur->tid.x % 32;
  s.warps[w].xchg[l] = v;
  __syncwarp(mask);
  const float r = s.warps[w].xchg[src & 31];
  __syncwarp(mask);
  return r;
}
static inline float __ldg(const float* p) { return *p; }
