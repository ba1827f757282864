// ping-pong timeline probe (scratch): who waits for what when two tiles share the tensor pipe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../chain_control_b200/csrc/cc_tmem.cuh"
#define KP 56
#define NP 64
struct Smem { float bhi[KP / 4 * NP * 4]; float blo[KP / 4 * NP * 4]; uint64_t a_ready[2]; uint64_t d_ready[2]; uint32_t tmem_base; unsigned int lock; };
constexpr uint32_t kIdesc = cc_umma_idesc_tf32(128, NP);
__global__ void __launch_bounds__(320, 1) k(const float* A, const float* W, float* out, long long* ev, int reps, int n_tiles, int terms, int rec_it, int two_issuers) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  Smem& S = *reinterpret_cast<Smem*>(smraw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NP * KP; i += blockDim.x) {
    const int n = i / KP, kk = i % KP;
    const float w = W[i];
    const float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    const int off = ((kk >> 2) * NP + n) * 4 + (kk & 3);
    S.bhi[off] = hi; S.blo[off] = w - hi;
  }
  if (warp == 8) {
    if (lane == 0) { for (int t = 0; t < 2; ++t) { cc_mbar_init(&S.a_ready[t], 128); cc_mbar_init(&S.d_ready[t], 1); } S.lock = 0; cc_mbar_fence_init(); }
    __syncwarp();
    cc_tmem_alloc(&S.tmem_base, 512);
  }
  cc_fence_proxy_async();
  cc_tc_fence_before();
  __syncthreads();
  cc_tc_fence_after();
  const uint32_t tbase = S.tmem_base;
  if (warp < 8 && (warp >> 2) < n_tiles) {
    const int tile = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t la = tbase + ((uint32_t)((warp & 3) * 32) << 16) + tile * 192;
    float x[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) x[i] = A[row * KP + i];
    uint32_t par = 0;
    const bool rec = blockIdx.x == 0 && lane == 0 && (warp & 3) == 0;
    for (int it = 0; it < reps; ++it) {
      long long* e = ev + ((it - rec_it) * 3 + tile) * 8;
      const bool r = rec && it >= rec_it && it < rec_it + 4;
      if (r) e[0] = clock64();
      {
        uint32_t h[32], l[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { h[i] = __float_as_uint(x[i]) & 0xFFFFE000u; l[i] = __float_as_uint(x[i] - __uint_as_float(h[i])); }
        cc_tmem_st32(la + 64, h); cc_tmem_st32(la + 128, l);
      }
      {
        uint32_t h[16], l[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { h[i] = __float_as_uint(x[32 + i]) & 0xFFFFE000u; l[i] = __float_as_uint(x[32 + i] - __uint_as_float(h[i])); }
        cc_tmem_st16(la + 64 + 32, h); cc_tmem_st16(la + 128 + 32, l);
      }
      {
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { h[i] = __float_as_uint(x[48 + i]) & 0xFFFFE000u; l[i] = __float_as_uint(x[48 + i] - __uint_as_float(h[i])); }
        cc_tmem_st8(la + 64 + 48, h); cc_tmem_st8(la + 128 + 48, l);
      }
      if (r) e[1] = clock64();
      cc_tmem_wait_st();
      if (r) e[2] = clock64();
      cc_tc_fence_before();
      cc_mbar_arrive(&S.a_ready[tile]);
      if (r) e[3] = clock64();
      cc_mbar_wait(&S.d_ready[tile], par);
      par ^= 1;
      cc_tc_fence_after();
      if (r) e[4] = clock64();
      uint32_t d0[32], d1[16], d2[8];
      cc_tmem_ld32(la, d0); cc_tmem_ld16(la + 32, d1); cc_tmem_ld8(la + 48, d2);
      cc_tmem_wait_ld();
      if (r) e[5] = clock64();
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = fmaf(1e-3f, __uint_as_float(d0[i]), x[i]);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[32 + i] = fmaf(1e-3f, __uint_as_float(d1[i]), x[32 + i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[48 + i] = fmaf(1e-3f, __uint_as_float(d2[i]), x[48 + i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < KP; ++i) s += x[i];
    out[(blockIdx.x * 2 + tile) * 128 + row] = s;
  } else if (warp >= 8 && lane == 0 && (two_issuers ? (warp - 8) < n_tiles : warp == 8)) {
    const uint32_t bhi = cc_smem_u32(S.bhi), blo = cc_smem_u32(S.blo);
    uint32_t par = 0;
    for (int it = 0; it < reps; ++it) {
      for (int tile = two_issuers ? warp - 8 : 0; tile < (two_issuers ? warp - 8 + 1 : n_tiles); ++tile) {
        long long* e = ev + ((it - rec_it) * 3 + 2) * 8 + tile * 4;
        const bool r = blockIdx.x == 0 && it >= rec_it && it < rec_it + 4;
        if (r) e[3] = clock64();
        cc_mbar_wait(&S.a_ready[tile], par);
        cc_tc_fence_after();
        if (r) e[0] = clock64();
        const uint32_t d = tbase + tile * 192, ahi = d + 64, alo = d + 128;
        uint32_t acc = 0;
        if (two_issuers == 2) while (atomicCAS(&S.lock, 0u, 1u) != 0u) __nanosleep(32);
        if (terms >= 3) {
#pragma unroll
          for (int js = 0; js < KP / 8; ++js) { cc_umma_tf32_ts(d, alo + js * 8, cc_umma_desc_kmajor(bhi + js * 2 * NP * 16, NP * 16, 128), kIdesc, acc); acc = 1; }
#pragma unroll
          for (int js = 0; js < KP / 8; ++js) { cc_umma_tf32_ts(d, ahi + js * 8, cc_umma_desc_kmajor(blo + js * 2 * NP * 16, NP * 16, 128), kIdesc, acc); acc = 1; }
        }
#pragma unroll
        for (int js = 0; js < KP / 8; ++js) { cc_umma_tf32_ts(d, ahi + js * 8, cc_umma_desc_kmajor(bhi + js * 2 * NP * 16, NP * 16, 128), kIdesc, acc); acc = 1; }
        if (r) e[1] = clock64();
        cc_tc_commit(&S.d_ready[tile]);
        if (two_issuers == 2) atomicExch(&S.lock, 0u);
        if (r) e[2] = clock64();
      }
      par ^= 1;
    }
  }
  cc_tc_fence_before();
  __syncthreads();
  if (warp == 8) cc_tmem_dealloc(tbase, 512);
}
int main() {
  std::vector<float> A(128 * KP), W(NP * KP);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : W) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.14f;
  float *dA, *dW, *dO; long long* dE;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dO, 148 * 256 * 4); cudaMalloc(&dE, 4096);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int two = 1; two < 3; ++two) for (int nt = 2; nt <= 2; ++nt) for (int terms = 3; terms <= 3; terms += 2) {
    const int reps = 1000;
    cudaMemset(dE, 0, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<148, 320, 65536>>>(dA, dW, dO, dE, reps, nt, terms, 500, two);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long ev[4 * 3 * 8];
    cudaMemcpy(ev, dE, sizeof(ev), cudaMemcpyDeviceToHost);
    printf("== two_issuers=%d tiles=%d terms=%d: %.3f ms, %.0f cyc/iter(at 1.965GHz), %.3e traj-steps/s(4 stages)\n", two, nt, terms, ms, ms * 1e-3 * 1.965e9 / reps, 148.0 * nt * reps * 128 / 4 / (ms * 1e-3));
    const long long t00 = ev[0];
    for (int it = 0; it < 3; ++it) {
      for (int tile = 0; tile < nt; ++tile) {
        long long* q = ev + (it * 3 + tile) * 8;
        printf("  it%d tile%d worker: start %lld | st issued +%lld | wait_st +%lld | arrive +%lld | d_ready wake +%lld | ld done +%lld\n", it, tile, q[0] - t00, q[1] - q[0], q[2] - q[1], q[3] - q[2], q[4] - q[3], q[5] - q[4]);
      }
      for (int tile = 0; tile < nt; ++tile) {
        long long* q = ev + (it * 3 + 2) * 8 + tile * 4;
        printf("  it%d tile%d issuer: wait begins %lld | a_ready wake %lld | mma issued +%lld | commit +%lld\n", it, tile, q[3] - t00, q[0] - t00, q[1] - q[0], q[2] - q[1]);
      }
    }
  }
  return 0;
}
