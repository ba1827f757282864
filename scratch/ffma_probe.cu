// FFMA issue-rate probes: immediate/constant operands vs three register operands (scratch)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe_imm(float* sink, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 0.999f, c = 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) { a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c); a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c); }
  }
  float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456f) sink[0] = s;
}
// acc[i][j] += x[i] * y[j]: the SGEMM register-tile pattern (3 register operands, operand reuse possible)
__global__ void probe_tile(float* sink, const float* in, int iters) {
  float acc[4][8];
  float x[4], y[8];
  for (int i = 0; i < 4; ++i) x[i] = in[threadIdx.x + i];
  for (int j = 0; j < 8; ++j) y[j] = in[threadIdx.x + 4 + j];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(x[i], y[j], acc[i][j]);
      // perturb operands a little so the loop is not folded (cheap: 2 ops per 32 FMAs)
      x[u & 3] += 1e-7f; y[u & 7] -= 1e-7f;
    }
  }
  float s = 0.f;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
  if (s == 123.456f) sink[0] = s;
}
// a_i = fma(a_i, b_i, c_i) with all-distinct registers
__global__ void probe_reg(float* sink, const float* in, int iters) {
  float a[8], b[8], c[8];
  for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; c[i] = in[threadIdx.x + 16 + i]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b[i], c[i]);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456f) sink[0] = s;
}
int main() {
  float *sink, *in;
  cudaMalloc(&sink, 64); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2048, blocks = 148 * 4, threads = 512;
  for (int which = 0; which < 3; ++which) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) probe_imm<<<blocks, threads>>>(sink, iters);
      if (which == 1) probe_reg<<<blocks, threads>>>(sink, in, iters);
      if (which == 2) probe_tile<<<blocks, threads>>>(sink, in, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (which == 2) ? (double)blocks * threads * iters * 4 * 32 : (double)blocks * threads * iters * 16 * 8;
    printf("%s: %.3f ms -> %.1f TFLOP/s\n", which == 0 ? "imm operands" : which == 1 ? "3 distinct regs" : "4x8 register tile", ms, 2 * fmas / ms / 1e9);
  }
  // single warp per SMSP
  for (int which = 1; which < 3; ++which) {
    cudaEventRecord(e0);
    if (which == 1) probe_reg<<<148, 128>>>(sink, in, iters);
    if (which == 2) probe_tile<<<148, 128>>>(sink, in, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (which == 2) ? 148.0 * 128 * iters * 4 * 32 : 148.0 * 128 * iters * 16 * 8;
    printf("1 warp/SMSP %s: %.3f ms -> %.1f TFLOP/s\n", which == 1 ? "3 distinct regs" : "4x8 register tile", ms, 2 * fmas / ms / 1e9);
  }
  return 0;
}
