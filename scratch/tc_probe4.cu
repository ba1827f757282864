// kind::f16 TS probe: D[128 x 64] = A_f16[128 x 64](TMEM, 2 per column) . W_f16[64 x 64]^T  with the scaled 3-term split
//   D = D1 + 2^-11 D2,  D1 = a_hi.w_hi,  D2 = a_hi.w_lo' + a_lo'.w_hi   (lo' = lo * 2^11 in fp16)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../chain_control_b200/csrc/cc_tmem.cuh"
#define KP 64
#define NP 64
struct Smem { __half bhi[KP / 8 * NP * 8]; __half blo[KP / 8 * NP * 8]; uint64_t bar; uint32_t tmem_base; };
// idesc kind::f16: D fp32 (1<<4), A fmt F16 = 0, B fmt F16 = 0
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) k(const float* A, const float* W, float* out) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  Smem& S = *reinterpret_cast<Smem*>(smraw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // B: K-major, no swizzle, 16-bit: core matrix = 8 rows x 16 B = 8 rows x 8 K-elements; chunk c = k/8: offset (c*NP + n)*8 + k%8
  for (int i = tid; i < NP * KP; i += blockDim.x) {
    const int n = i / KP, kk = i % KP;
    const float w = W[i];
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn((w - __half2float(hi)) * 2048.f);
    const int off = ((kk >> 3) * NP + n) * 8 + (kk & 7);
    S.bhi[off] = hi; S.blo[off] = lo;
  }
  if (warp == 0) {
    if (lane == 0) { cc_mbar_init(&S.bar, 1); cc_mbar_fence_init(); }
    __syncwarp();
    cc_tmem_alloc(&S.tmem_base, 512);
  }
  cc_fence_proxy_async(); cc_tc_fence_before(); __syncthreads(); cc_tc_fence_after();
  const uint32_t tb = S.tmem_base;
  const int row = warp * 32 + lane;
  const uint32_t la = tb + ((uint32_t)(warp * 32) << 16);
  // A hi at cols [128, 160), lo' at [160, 192): element k of the row in column k/2, low half = even k
  uint32_t hi[32], lo[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const float v0 = A[row * KP + 2 * c], v1 = A[row * KP + 2 * c + 1];
    const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
    const __half l0 = __float2half_rn((v0 - __half2float(h0)) * 2048.f), l1 = __float2half_rn((v1 - __half2float(h1)) * 2048.f);
    hi[c] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    lo[c] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
  }
  cc_tmem_st32(la + 128, hi); cc_tmem_st32(la + 160, lo);
  cc_tmem_wait_st(); cc_tc_fence_before(); __syncthreads(); cc_tc_fence_after();
  if (tid == 0) {
    const uint32_t bhi = cc_smem_u32(S.bhi), blo = cc_smem_u32(S.blo);
    // per MMA: K = 16 = 2 chunks of 8 (16 B each): LBO = NP*16 B, SBO = 128 B; A advances 8 columns per K = 16
    for (int js = 0; js < KP / 16; ++js) mma_f16_ts(tb, tb + 128 + js * 8, cc_umma_desc_kmajor(bhi + js * 2 * NP * 16, NP * 16, 128), kIdesc, js ? 1u : 0u);      // D1 = hi.hi
    for (int js = 0; js < KP / 16; ++js) mma_f16_ts(tb + 64, tb + 128 + js * 8, cc_umma_desc_kmajor(blo + js * 2 * NP * 16, NP * 16, 128), kIdesc, js ? 1u : 0u);  // D2 = hi.lo'
    for (int js = 0; js < KP / 16; ++js) mma_f16_ts(tb + 64, tb + 160 + js * 8, cc_umma_desc_kmajor(bhi + js * 2 * NP * 16, NP * 16, 128), kIdesc, 1u);           // D2 += lo'.hi
    cc_tc_commit(&S.bar);
  }
  cc_mbar_wait(&S.bar, 0); cc_tc_fence_after();
  uint32_t d1[64], d2[64];
  cc_tmem_ld64(la, d1); cc_tmem_ld64(la + 64, d2); cc_tmem_wait_ld();
#pragma unroll
  for (int n = 0; n < 64; ++n) out[row * 64 + n] = fmaf(__uint_as_float(d2[n]), 1.f / 2048.f, __uint_as_float(d1[n]));
  cc_tc_fence_before(); __syncthreads();
  if (warp == 0) cc_tmem_dealloc(tb, 512);
}
int main() {
  std::vector<float> A(128 * KP), W(NP * KP);
  srand(1);
  for (auto& v : A) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 3.f;
  for (auto& v : W) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.14f;
  float *dA, *dW, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<<<1, 128, 65536>>>(dA, dW, dO);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> O(128 * 64);
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0, e32 = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
    double r = 0; float r32 = 0;
    for (int kk = 0; kk < KP; ++kk) { r += (double)A[m * KP + kk] * W[n * KP + kk]; r32 = fmaf(A[m * KP + kk], W[n * KP + kk], r32); }
    maxerr = fmax(maxerr, fabs(O[m * 64 + n] - r)); e32 = fmax(e32, fabs(r32 - r)); maxref = fmax(maxref, fabs(r));
  }
  printf("f16 scaled 3-term: max|err| = %.3e (fp32 chain %.3e), max|ref| = %.3e, D[0][0..3] = %g %g %g %g\n", maxerr, e32, maxref, O[0], O[1], O[2], O[3]);
  return 0;
}
