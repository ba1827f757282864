// TMEM ld/st shape throughput (scratch)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../chain_control_b200/csrc/cc_tmem.cuh"
__shared__ uint32_t s_base;
template <int SH>
__global__ void __launch_bounds__(288, 1) k(float* out, long long* cyc, int mode, int reps, int nwarps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8) cc_tmem_alloc(&s_base, 512);
  cc_tc_fence_before();
  __syncthreads();
  cc_tc_fence_after();
  const uint32_t tb = s_base;
  if (warp < nwarps) {
    const uint32_t la = tb + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
      if (mode == 0) {  // ld 64 cols
        if (SH == 8) { uint32_t r[8][8];
#pragma unroll
          for (int c = 0; c < 8; ++c) cc_tmem_ld8(la + c * 8, r[c]);
          cc_tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += __uint_as_float(r[c][i]);
        } else if (SH == 16) { uint32_t r[4][16];
#pragma unroll
          for (int c = 0; c < 4; ++c) cc_tmem_ld16(la + c * 16, r[c]);
          cc_tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc += __uint_as_float(r[c][i]);
        } else if (SH == 32) { uint32_t r[2][32];
#pragma unroll
          for (int c = 0; c < 2; ++c) cc_tmem_ld32(la + c * 32, r[c]);
          cc_tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[c][i]);
        } else { uint32_t r[64];
          cc_tmem_ld64(la, r);
          cc_tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 64; ++i) acc += __uint_as_float(r[i]);
        }
      } else if (mode == 1) {  // st 128 cols
        if (SH == 8) {
#pragma unroll
          for (int c = 0; c < 16; ++c) { uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = it + i + c;
            cc_tmem_st8(la + c * 8, r); }
        } else if (SH == 16) {
#pragma unroll
          for (int c = 0; c < 8; ++c) { uint32_t r[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = it + i + c;
            cc_tmem_st16(la + c * 16, r); }
        } else if (SH == 32) {
#pragma unroll
          for (int c = 0; c < 4; ++c) { uint32_t r[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = it + i + c;
            cc_tmem_st32(la + c * 32, r); }
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c) { uint32_t r[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) r[i] = it + i + c;
            cc_tmem_st64(la + c * 64, r); }
        }
        cc_tmem_wait_st();
      } else {  // chunked ld16 -> wait -> use, 4 chunks (dependent waits)
        uint32_t r[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          cc_tmem_ld16(la + c * 16, r);
          cc_tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc += __uint_as_float(r[i]);
        }
      }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = acc;
  }
  cc_tc_fence_before();
  __syncthreads();
  if (warp == 8) cc_tmem_dealloc(tb, 512);
}
template <int SH> void run(float* o, long long* c) {
  for (int mode = 0; mode < 3; ++mode)
    for (int nw = 1; nw <= 8; nw *= 2) {
      if (mode == 2 && SH != 16) continue;
      k<SH><<<1, 288>>>(o, c, mode, 2000, nw);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); exit(1); }
      long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      const int cols = mode == 1 ? 128 : 64;
      printf("%s x%d warps=%d: %.1f cyc per %d cols/warp  (%.1f B/clk/SM)\n", mode == 0 ? "ld" : mode == 1 ? "st" : "ld16-chunked", SH, nw, h / 2000.0, cols, nw * 32.0 * cols * 4 / (h / 2000.0));
    }
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 4096); cudaMalloc(&c, 64);
  run<8>(o, c); run<16>(o, c); run<32>(o, c); run<64>(o, c);
  return 0;
}
