// tcgen05 probe (scratch): D[128 x 64] = A[128 x 56] . W[64 x 56]^T with A in TMEM (TS mode), B = weights in
// shared memory (K-major, no swizzle), 3xTF32 split for fp32-class accuracy; plus the timing skeleton of the
// planned rollout pipeline (workers: tcgen05.ld D -> fma -> tcgen05.st A ; one thread: 21 MMAs + commit).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define KP 56   // padded contraction
#define NP 64   // padded outputs
#define KCH (KP / 4)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ uint64_t make_bdesc(uint32_t saddr) {
  // K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B contiguous (128 B); SBO (next 8 rows) = 128 B; LBO (next 16 B K chunk) = NP*16 B
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(((NP * 16) >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((128 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version 1 (sm_100)
  return d;
}
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct Smem {
  float bhi[KCH * NP * 4];  // [kchunk][n][4]
  float blo[KCH * NP * 4];
  uint64_t a_ready[2];
  uint64_t d_ready[2];
  uint32_t tmem_base;
};

// mode 0: one product, written to out[128][64];  mode 1: MMA issue throughput; mode 2: ping-pong with n_tiles tiles
__global__ void __launch_bounds__(288, 1) tc_probe(const float* A, const float* W, float* out, long long* cyc, int mode, int reps, int n_tiles, int terms) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  Smem& S = *reinterpret_cast<Smem*>(smraw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // stage weights: W[n][k] row-major (n < 64 rows given, k < 56) -> hi/lo, chunked
  for (int i = tid; i < NP * KP; i += blockDim.x) {
    const int n = i / KP, k = i % KP;
    const float w = W[i];
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    const float lo = w - hi;
    const int off = ((k >> 2) * NP + n) * 4 + (k & 3);
    S.bhi[off] = hi;
    S.blo[off] = lo;
  }
  if (warp == 8) {
    if (lane == 0) {
      for (int t = 0; t < 2; ++t) { mbar_init(&S.a_ready[t], 128); mbar_init(&S.d_ready[t], 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // make generic-proxy smem writes (weights) visible to the async proxy (UMMA reads smem through it)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = S.tmem_base;
  // per tile: D cols [0,64), A_hi [64,120), A_lo [128,184)  (+ tile * 192)
  if (warp < 8 && (warp >> 2) < n_tiles) {
    const int tile = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + tile * 192;
    float x[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) x[k] = A[row * KP + k];
    uint32_t par = 0;
    const int iters = mode == 2 ? reps : 1;
    if (mode == 3 || mode == 4) {
      long long t0 = clock64();
      float acc = 0.f;
      for (int it = 0; it < reps; ++it) {
        if (mode == 3) {
#pragma unroll
          for (int c = 0; c < 14; ++c) {
            uint32_t h[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = __float_as_uint(x[(c * 8 + i) % KP]) + it;
            tmem_st8(lane_addr + 64 + c * 8, h);
          }
          tmem_wait_st();
        } else {
          float d[NP];
#pragma unroll
          for (int c = 0; c < NP / 8; ++c) tmem_ld8(lane_addr + c * 8, d + c * 8);
          tmem_wait_ld();
#pragma unroll
          for (int n = 0; n < NP; ++n) acc += d[n];
        }
      }
      long long t1 = clock64();
      if (lane == 0 && (warp & 3) == 0) cyc[1 + tile] = t1 - t0;
      out[tile * 128 + row] = acc;
    } else {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // write A hi / lo
#pragma unroll
      for (int c = 0; c < KP / 8; ++c) {
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v = x[c * 8 + i];
          h[i] = __float_as_uint(v) & 0xFFFFE000u;
          l[i] = __float_as_uint(v - __uint_as_float(h[i]));
        }
        tmem_st8(lane_addr + 64 + c * 8, h);
        tmem_st8(lane_addr + 128 + c * 8, l);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&S.a_ready[tile]);
      mbar_wait(&S.d_ready[tile], par);
      par ^= 1;
      tc_fence_after();
      float d[NP];
#pragma unroll
      for (int c = 0; c < NP / 8; ++c) tmem_ld8(lane_addr + c * 8, d + c * 8);
      tmem_wait_ld();
      if (mode == 2) {
        // next "stage input": x <- x + 1e-3 * d  (keeps the chain dependent, like an RK stage)
#pragma unroll
        for (int k = 0; k < KP; ++k) x[k] = fmaf(1e-3f, d[k], x[k]);
      } else {
#pragma unroll
        for (int n = 0; n < NP; ++n) out[row * NP + n] = d[n];
      }
    }
    long long t1 = clock64();
    if (mode == 2 && lane == 0 && (warp & 3) == 0) cyc[1 + tile] = t1 - t0;
    if (mode == 2) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k) s += x[k];
      out[(tile * 128 + row)] = s;
    }
    }
  } else if (warp == 8 && lane == 0) {
    // MMA issuer
    const uint32_t bhi = smem_u32(S.bhi), blo = smem_u32(S.blo);
    if (mode == 3 || mode == 4) {
    } else if (mode == 1) {
      mbar_wait(&S.a_ready[0], 0);
      tc_fence_after();
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r)
        for (int j = 0; j < 3 * (KP / 8); ++j) {
          const int js = j % (KP / 8);
          tc_mma_ts(tbase, tbase + 256 + js * 8, make_bdesc(bhi + js * 2 * NP * 16), (kIdesc & ~(63u << 17)) | ((uint32_t)(terms >> 3) << 17), (r | j) ? 1u : 0u);
        }
      tc_commit(&S.d_ready[0]);
      mbar_wait(&S.d_ready[0], 0);
      long long t1 = clock64();
      cyc[0] = t1 - t0;
      // let the workers finish: they wait on d_ready[0] phase 0 as well (already complete)
    } else {
      const int iters = mode == 2 ? reps : 1;
      uint32_t par = 0;
      for (int it = 0; it < iters; ++it) {
        for (int tile = 0; tile < n_tiles; ++tile) {
          mbar_wait(&S.a_ready[tile], par);
          tc_fence_after();
          const uint32_t d = tbase + tile * 192, ahi = d + 64, alo = d + 128;
          uint32_t acc = 0;
          if (terms >= 3) {
            for (int js = 0; js < KP / 8; ++js) { tc_mma_ts(d, alo + js * 8, make_bdesc(bhi + js * 2 * NP * 16), kIdesc, acc); acc = 1; }
            for (int js = 0; js < KP / 8; ++js) { tc_mma_ts(d, ahi + js * 8, make_bdesc(blo + js * 2 * NP * 16), kIdesc, acc); acc = 1; }
          }
          for (int js = 0; js < KP / 8; ++js) { tc_mma_ts(d, ahi + js * 8, make_bdesc(bhi + js * 2 * NP * 16), kIdesc, acc); acc = 1; }
          tc_commit(&S.d_ready[tile]);
        }
        par ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
  }
}

int main(int argc, char** argv) {
  const int M = 256;
  std::vector<float> A(M * KP), W(NP * KP);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : W) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.14f;
  float *dA, *dW, *dO;
  long long* dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dO, M * NP * 4); cudaMalloc(&dC, 64);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, M * NP * 4);
  cudaFuncSetAttribute(tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  auto check = [&](const char* what) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", what, cudaGetErrorString(e)); exit(1); }
  };
  for (int terms = 1; terms <= 3; terms += 2) {
    tc_probe<<<1, 288, sizeof(Smem) + 1024>>>(dA, dW, dO, dC, 0, 1, 1, terms);
    check("mode0");
    std::vector<float> O(128 * NP);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0, maxerr32 = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < NP; ++n) {
        double r = 0; float r32 = 0.f;
        for (int k = 0; k < KP; ++k) { r += (double)A[m * KP + k] * W[n * KP + k]; r32 = fmaf(A[m * KP + k], W[n * KP + k], r32); }
        maxerr = fmax(maxerr, fabs(O[m * NP + n] - r));
        maxerr32 = fmax(maxerr32, fabs(r32 - r));
        maxref = fmax(maxref, fabs(r));
      }
    printf("terms=%d  max|err|=%.3e  (fp32 fma chain: %.3e)  max|ref|=%.3e  D[0][0..3]=%g %g %g %g\n", terms, maxerr, maxerr32, maxref, O[0], O[1], O[2], O[3]);
  }
  // MMA issue throughput
  {
    const int reps = 400;
    for (int N = 16; N <= 256; N *= 2) {
    tc_probe<<<1, 288, 65536>>>(dA, dW, dO, dC, 1, reps, 1, N);
    check("mode1");
    long long c[4];
    cudaMemcpy(c, dC, 32, cudaMemcpyDeviceToHost);
    printf("MMA throughput: %lld cycles for %d MMAs (M128 N%d K8 tf32, TS) = %.1f cyc/MMA\n", c[0], reps * 21, N, (double)c[0] / (reps * 21));
    }
    for (int mode = 3; mode <= 4; ++mode) for (int nt = 1; nt <= 2; ++nt) {
      tc_probe<<<1, 288, 65536>>>(dA, dW, dO, dC, mode, 1000, nt, 3);
      check("mode34");
      long long c[4];
      cudaMemcpy(c, dC, 32, cudaMemcpyDeviceToHost);
      const int cols = mode == 3 ? 112 : 64;
      printf("%s only, tiles=%d: %.1f cyc per %d cols x 128 lanes (%.1f B/clk/SM)\n", mode == 3 ? "tcgen05.st" : "tcgen05.ld", nt, (double)c[1] / 1000, cols, nt * 128.0 * cols * 4 / ((double)c[1] / 1000));
    }
  }
  for (int nt = 1; nt <= 2; ++nt)
    for (int terms = 1; terms <= 3; terms += 2) {
      const int reps = 2000;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      tc_probe<<<148, 288, sizeof(Smem) + 1024>>>(dA, dW, dO, dC, 2, reps, nt, terms);
      cudaEventRecord(e1);
      check("mode2");
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long c[4];
      cudaMemcpy(c, dC, 32, cudaMemcpyDeviceToHost);
      printf("ping-pong tiles=%d terms=%d: %.1f cyc/stage (tile0) %.1f (tile1); 148 CTAs: %.3f ms -> %.3e tile-stages/s/GPU -> %.3e traj-steps/s (4 stages)\n", nt, terms,
             (double)c[1] / reps, nt > 1 ? (double)c[2] / reps : 0.0, ms, 148.0 * nt * reps / (ms * 1e-3), 148.0 * nt * reps * 128 / 4 / (ms * 1e-3));
    }
  return 0;
}
