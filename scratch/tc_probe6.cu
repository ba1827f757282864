// MN-major operand probe (scratch): which LBO/SBO assignment does tcgen05.mma kind::tf32 expect for SWIZZLE_NONE MN-major
// operands?  (1) TS mode, B = the forward weight buffer of cc_mlp_kernels.cuh re-read as W^T; (2) SS mode, both operands
// MN-major with K = trajectory (weight gradient).  Integer-valued data -> exact comparison with the CPU.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../chain_control_b200/csrc/cc_tmem.cuh"
#define NP 64
#define KP 80  // staged columns (72 logical + 8 so that N' = 80 stays inside the buffer)
struct Smem {
  float w[KP / 4 * NP * 4];     // element (n, k) at ((k/4)*NP + n)*4 + k%4
  float opa[32 * 512];          // (t, r): r-group r/4 at 2048 B, t-group t/8 at 128 B, t%8 at 16 B
  float opb[20 * 512];
  uint64_t bar;
  uint32_t tmem_base;
};
__host__ __device__ constexpr uint32_t idesc(int M, int N, int amn, int bmn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__global__ void __launch_bounds__(128, 1) k(const float* W, const float* Dl, const float* H, float* out0, float* out1, float* out2, int variant, int n1) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  Smem& S = *reinterpret_cast<Smem*>(smraw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NP * KP; i += 128) {
    const int n = i / KP, kk = i % KP;
    S.w[((kk >> 2) * NP + n) * 4 + (kk & 3)] = W[i];
  }
  for (int r = 0; r < 128; ++r) S.opa[(r / 4) * 512 + (tid / 8) * 32 + (tid % 8) * 4 + (r % 4)] = Dl[tid * 128 + r];
  for (int j = 0; j < 80; ++j) S.opb[(j / 4) * 512 + (tid / 8) * 32 + (tid % 8) * 4 + (j % 4)] = H[tid * 80 + j];
  if (warp == 0) {
    if (lane == 0) { cc_mbar_init(&S.bar, 1); cc_mbar_fence_init(); }
    __syncwarp();
    cc_tmem_alloc(&S.tmem_base, 512);
  }
  cc_fence_proxy_async();
  cc_tc_fence_before();
  __syncthreads();
  cc_tc_fence_after();
  const uint32_t tbase = S.tmem_base;
  const uint32_t la = tbase + ((uint32_t)(warp * 32) << 16);
  // A of test 1: delta[t][n] = Dl[t*128 + n], n < 64, at columns 256..319
  {
    uint32_t a[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) a[i] = __float_as_uint(Dl[tid * 128 + i]);
    cc_tmem_st64(la + 256, a);
    cc_tmem_wait_st();
  }
  cc_tc_fence_before();
  __syncthreads();
  cc_tc_fence_after();
  if (tid == 0) {
    const uint32_t wb = cc_smem_u32(S.w), oa = cc_smem_u32(S.opa), ob = cc_smem_u32(S.opb);
    // test 1: D1[t][k] = sum_n delta[t][n] W[n][k]  -> columns 0..n1-1
    const uint32_t id1 = idesc(128, n1, 0, 1);
    for (int js = 0; js < 8; ++js) {
      const uint64_t bd = variant == 0 ? cc_umma_desc_kmajor(wb + js * 128, 128, NP * 16) : cc_umma_desc_kmajor(wb + js * 128, NP * 16, 128);
      cc_umma_tf32_ts(tbase, tbase + 256 + 8 * js, bd, id1, js ? 1u : 0u);
    }
    // test 2: D2[r][j] = sum_t delta[t][r] h[t][j]  -> columns 128..207
    const uint32_t id2 = idesc(128, 80, 1, 1);
    for (int c = 0; c < 16; ++c) {
      const uint64_t ad = variant == 0 ? cc_umma_desc_kmajor(oa + c * 128, 128, 2048) : cc_umma_desc_kmajor(oa + c * 128, 2048, 128);
      const uint64_t bd = variant == 0 ? cc_umma_desc_kmajor(ob + c * 128, 128, 2048) : cc_umma_desc_kmajor(ob + c * 128, 2048, 128);
      cc_umma_tf32_ss(tbase + 128, ad, bd, id2, c ? 1u : 0u);
    }
    // control: K-major forward product D0[t][n] = sum_{k<64} delta[t][k] W[n][k] -> columns 400..463
    for (int js = 0; js < 8; ++js) cc_umma_tf32_ts(tbase + 400, tbase + 256 + 8 * js, cc_umma_desc_kmajor(wb + js * 2 * NP * 16, NP * 16, 128), idesc(128, 64, 0, 0), js ? 1u : 0u);
    cc_tc_commit(&S.bar);
  }
  cc_mbar_wait(&S.bar, 0);
  cc_tc_fence_after();
  {
    uint32_t d[16];
    for (int c0 = 0; c0 < 80; c0 += 16) {
      cc_tmem_ld16(la + c0, d);
      cc_tmem_wait_ld();
      for (int i = 0; i < 16; ++i) out1[tid * 80 + c0 + i] = __uint_as_float(d[i]);
      cc_tmem_ld16(la + 128 + c0, d);
      cc_tmem_wait_ld();
      for (int i = 0; i < 16; ++i) out2[tid * 80 + c0 + i] = __uint_as_float(d[i]);
      if (c0 < 64) {
        cc_tmem_ld16(la + 400 + c0, d);
        cc_tmem_wait_ld();
        for (int i = 0; i < 16; ++i) out0[tid * 64 + c0 + i] = __uint_as_float(d[i]);
      }
    }
  }
  cc_tc_fence_before();
  __syncthreads();
  if (warp == 0) cc_tmem_dealloc(tbase, 512);
}
int main() {
  std::vector<float> W(NP * KP), D(128 * 128), H(128 * 80);
  srand(7);
  auto rnd = [] { return (float)(rand() % 17 - 8); };
  for (auto& v : W) v = rnd();
  for (auto& v : D) v = rnd();
  for (auto& v : H) v = rnd();
  float *dW, *dD, *dH, *o0, *o1, *o2; cudaMalloc(&o0, 128 * 64 * 4);
  cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dH, H.size() * 4); cudaMalloc(&o1, 128 * 80 * 4); cudaMalloc(&o2, 128 * 80 * 4);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dH, H.data(), H.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024);
  for (int variant = 0; variant < 2; ++variant)
    for (int n1 = 64; n1 <= 80; n1 += 16) {
      cudaMemset(o1, 0, 128 * 80 * 4); cudaMemset(o2, 0, 128 * 80 * 4);
      k<<<1, 128, sizeof(Smem) + 1024>>>(dW, dD, dH, o0, o1, o2, variant, n1);
      cudaError_t e0 = cudaGetLastError(); if (e0 != cudaSuccess) { printf("launch err %s\n", cudaGetErrorString(e0)); return 1; }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d n1 %d: err %s\n", variant, n1, cudaGetErrorString(e)); return 1; }
      std::vector<float> r1(128 * 80), r2(128 * 80);
      cudaMemcpy(r1.data(), o1, r1.size() * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(r2.data(), o2, r2.size() * 4, cudaMemcpyDeviceToHost);
      std::vector<float> r0(128 * 64);
      cudaMemcpy(r0.data(), o0, r0.size() * 4, cudaMemcpyDeviceToHost);
      double e1 = 0, e2 = 0, ec = 0;
      for (int t = 0; t < 128; ++t)
        for (int n = 0; n < 64; ++n) {
          double s = 0;
          for (int kk = 0; kk < 64; ++kk) s += (double)D[t * 128 + kk] * W[n * KP + kk];
          ec = fmax(ec, fabs(s - r0[t * 64 + n]));
        }
      printf("control K-major err %.1f; r1[5][0..3] = %.0f %.0f %.0f %.0f ; r2[5][0..3] = %.0f %.0f %.0f %.0f\n", ec, r1[400], r1[401], r1[402], r1[403], r2[400], r2[401], r2[402], r2[403]);
      for (int t = 0; t < 128; ++t)
        for (int kk = 0; kk < n1; ++kk) {
          double s = 0;
          for (int n = 0; n < 64; ++n) s += (double)D[t * 128 + n] * W[n * KP + kk];
          e1 = fmax(e1, fabs(s - r1[t * 80 + kk]));
        }
      for (int r = 0; r < 128; ++r)
        for (int j = 0; j < 80; ++j) {
          double s = 0;
          for (int t = 0; t < 128; ++t) s += (double)D[t * 128 + r] * H[t * 80 + j];
          e2 = fmax(e2, fabs(s - r2[r * 80 + j]));
        }
      printf("variant %d (%s) n1=%d: TS W^T max err %.1f | SS MN-major max err %.1f\n", variant, variant == 0 ? "lbo=K-group stride, sbo=MN-group stride" : "swapped", n1, e1, e2);
    }
  return 0;
}
