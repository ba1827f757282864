// SS-mode kind::tf32 probe for the weight-gradient product: D[128 x 64] = A[128 rows x 128 traj] . B[64 rows x 128 traj]^T,
// both operands K-major (K = trajectory) without swizzle, written by thread = trajectory: LBO = 144 B (padded core matrices,
// conflict-free 4-byte stores), SBO = 32 * 144 B.  Also: rounding of a long accumulation chain, and cycles per MMA.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../chain_control_b200/csrc/cc_tmem.cuh"
#define LBO 144
#define SBO (32 * 144)
#define REG (16 * SBO)
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
struct Ctl { uint64_t bar; uint32_t tmem_base; };
__global__ void __launch_bounds__(128, 1) k(const float* A, const float* Bm, float* out, int reps, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* opA = sm;
  unsigned char* opB = sm + REG;
  Ctl& S = *reinterpret_cast<Ctl*>(sm + 2 * REG);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // thread t = trajectory t writes element (row, t) of both operands
  const int tb_off = (tid >> 2) * LBO + (tid & 3) * 4;
  for (int r = 0; r < 128; ++r) *reinterpret_cast<float*>(opA + (r >> 3) * SBO + (r & 7) * 16 + tb_off) = A[r * 128 + tid];
  for (int r = 0; r < 64; ++r) *reinterpret_cast<float*>(opB + (r >> 3) * SBO + (r & 7) * 16 + tb_off) = Bm[r * 128 + tid];
  if (warp == 0) {
    if (lane == 0) { cc_mbar_init(&S.bar, 1); cc_mbar_fence_init(); }
    __syncwarp();
    cc_tmem_alloc(&S.tmem_base, 512);
  }
  cc_fence_proxy_async(); cc_tc_fence_before(); __syncthreads(); cc_tc_fence_after();
  const uint32_t tb = S.tmem_base;
  constexpr uint32_t idesc = cc_umma_idesc_tf32(128, 64);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t a = cc_smem_u32(opA), b = cc_smem_u32(opB);
    t0 = clock64();
    for (int r = 0; r < reps; ++r)
#pragma unroll
      for (int js = 0; js < 16; ++js) mma_ss(tb, cc_umma_desc_kmajor(a + js * 2 * LBO, LBO, SBO), cc_umma_desc_kmajor(b + js * 2 * LBO, LBO, SBO), idesc, (r | js) ? 1u : 0u);
    cc_tc_commit(&S.bar);
  }
  cc_mbar_wait(&S.bar, 0); cc_tc_fence_after();
  if (tid == 0) { t1 = clock64(); cyc[0] = t1 - t0; }
  uint32_t r[64];
  cc_tmem_ld64(tb + ((uint32_t)(warp * 32) << 16), r);
  cc_tmem_wait_ld();
  for (int j = 0; j < 64; ++j) out[tid * 64 + j] = __uint_as_float(r[j]);
  cc_tc_fence_before(); __syncthreads();
  if (warp == 0) cc_tmem_dealloc(tb, 512);
}
static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
#include <cstring>
int main() {
  std::vector<float> A(128 * 128), B(64 * 128), out(128 * 64);
  float *dA, *dB, *dO; long long* dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dC, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * REG + 64);
  for (int test = 0; test < 3; ++test) {
    // 0: random tf32-exact operands, reps 1 (layout check)  1: stacked hi/lo scheme on fp32 data  2: rounding of a same-sign chain
    srand(1 + test);
    std::vector<float> z(64 * 128), a(64 * 128);
    for (auto& v : z) v = (float)rand() / RAND_MAX * 2 - 1;
    for (auto& v : a) v = (float)rand() / RAND_MAX * 2 - 1;
    int reps = 1;
    if (test == 0) {
      for (auto& v : A) v = tf32_trunc((float)rand() / RAND_MAX * 2 - 1);
      for (auto& v : B) v = tf32_trunc((float)rand() / RAND_MAX * 2 - 1);
    } else if (test == 1) {
      // A rows n < 64: z_hi, rows 64 + n: z_lo ; two launches: B = a_hi, then B = a_lo (host adds the results)
      for (int n = 0; n < 64; ++n) for (int t = 0; t < 128; ++t) { const float h = tf32_trunc(z[n * 128 + t]); A[n * 128 + t] = h; A[(64 + n) * 128 + t] = z[n * 128 + t] - h; }
    } else {
      for (auto& v : A) v = tf32_trunc(0.5f + 0.5f * (float)rand() / RAND_MAX);
      for (auto& v : B) v = tf32_trunc(0.5f + 0.5f * (float)rand() / RAND_MAX);
      reps = 256;  // 4096 dependent accumulations
    }
    std::vector<double> ref(128 * 64, 0.0);
    std::vector<float> acc(128 * 64, 0.f);
    const int passes = test == 1 ? 2 : 1;
    for (int ps = 0; ps < passes; ++ps) {
      if (test == 1) for (int j = 0; j < 64; ++j) for (int t = 0; t < 128; ++t) { const float h = tf32_trunc(a[j * 128 + t]); B[j * 128 + t] = ps == 0 ? h : a[j * 128 + t] - h; }
      cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
      k<<<1, 128, 2 * REG + 64>>>(dA, dB, dO, reps, dC);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
      for (size_t i = 0; i < out.size(); ++i) acc[i] += out[i];
    }
    long long cyc; cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0, bias = 0;
    if (test == 1) {
      for (int n = 0; n < 64; ++n) for (int j = 0; j < 64; ++j) {
        double r = 0; for (int t = 0; t < 128; ++t) r += (double)z[n * 128 + t] * a[j * 128 + t];
        const double got = (double)acc[n * 64 + j] + acc[(64 + n) * 64 + j];
        maxerr = fmax(maxerr, fabs(got - r)); maxref = fmax(maxref, fabs(r));
      }
    } else {
      for (int m = 0; m < 128; ++m) for (int j = 0; j < 64; ++j) {
        double r = 0; for (int t = 0; t < 128; ++t) r += (double)A[m * 128 + t] * B[j * 128 + t];
        r *= reps;
        maxerr = fmax(maxerr, fabs(acc[m * 64 + j] - r)); maxref = fmax(maxref, fabs(r)); bias += (acc[m * 64 + j] - r) / r;
      }
    }
    printf("test %d reps %d: max abs err %.3e (max |ref| %.3e, rel %.3e) mean signed rel err %.3e | %lld cycles for %d MMAs = %.1f / MMA\n", test, reps, maxerr, maxref, maxerr / maxref, bias / (128 * 64), cyc, reps * 16, (double)cyc / (reps * 16));
  }
  return 0;
}
