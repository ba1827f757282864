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
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace ccemu {
struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  uint3e tid;
  bool done = false;
};
struct State {
  ucontext_t sched;
  Fiber* cur = nullptr;
  uint3e bid{0, 0, 0}, bdim{1, 1, 1}, gdim{1, 1, 1};
  float* smem = nullptr;
  std::function<void()> body;
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
    int alive = block;
    while (alive > 0) {
      int finished_now = 0, yielded = 0;
      for (int t = 0; t < block; ++t) {
        Fiber& fb = fibers[t];
        if (fb.done) continue;
        s.cur = &fb;
        swapcontext(&s.sched, &fb.ctx);
        if (fb.done) ++finished_now; else ++yielded;
      }
      alive -= finished_now;
      if (finished_now && yielded) {
        fprintf(stderr, "ccemu: barrier divergence: %d threads exited while %d wait at __syncthreads\n",
                finished_now, yielded);
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
  swapcontext(&s.cur->ctx, &s.sched);
}
static inline float __ldg(const float* p) { return *p; }
