// cc_aux.cuh — small helper kernels: fused MSE value+cotangent (cc/utils/utils.py:53-55),
// deterministic reduction of per-CTA gradient partials into flat theta_bar, FP32 peak probe.
#pragma once
#include "cc_kernels.cuh"

// ybar = 2*(yhat - y)*inv_n ; block partials of sum((yhat-y)^2)*inv_n reduced in fixed order
__global__ void cc_mse_kernel(const float* y, const float* yhat, float* ybar, float* block_part, long long n, float inv_n) {
#ifndef CC_EMULATE
  __shared__ float red[256];
#else
  float* red = ccemu::st().smem;
#endif
  const int tid = threadIdx.x;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = yhat[i] - y[i];
    s = fmaf(d, d, s);
    ybar[i] = 2.f * inv_n * d;
  }
  red[tid] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (tid < w) red[tid] += red[tid + w];
    __syncthreads();
  }
  if (tid == 0) block_part[blockIdx.x] = red[0] * inv_n;
}
__global__ void cc_sum_partials_kernel(const float* part, int n, float* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += part[i];
    out[0] = s;
  }
}

// theta_bar[j] = sum over warp tiles (fixed order) of the gradient-record entry that maps to
// parameter j.  One thread per (layer, n, r) entry of the node's record ([N][K] blocks).
__global__ void cc_reduce_grad_kernel(const CcKParams p, int node, float* theta_bar, int n_tiles) {
  const CcNodeDev& nd = p.nd[node];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // index inside the node's record
  if (idx >= nd.g_total) return;
  for (int w = 0; w < 2; ++w) {
    const CcMlpDev& mlp = w ? nd.g : nd.f;
    for (int l = 0; l < mlp.L; ++l) {
      const CcLayerDev& L = mlp.layer[l];
      const int sz = L.N * L.K;
      if (idx >= L.g_off && idx < L.g_off + sz) {
        const int n = (idx - L.g_off) / L.K, r = (idx - L.g_off) % L.K;
        const int lc = cc_row_to_logical(nd, L, r);
        int dst = -1;
        if (lc >= 0) dst = L.w_off + n * L.in_log + lc;
        else if (lc == -2 && L.b_off >= 0) dst = L.b_off + n;
        if (dst < 0) return;
        float s = 0.f;
        for (int t = 0; t < n_tiles; ++t) s += p.gpart[(long long)t * p.g_rec + nd.g_base + idx];
        theta_bar[dst] = s;
        return;
      }
    }
  }
}

// 8 independent FFMA chains per thread, register resident: 2*8*iters*unroll flops per thread
__global__ void cc_fp32_probe_kernel(float* sink, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float b = 0.999f, c = 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
      a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
  }
  const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456f) sink[0] = s;
}

// acc[i][j] += x[i] * y[j]: the SGEMM register-tile pattern (3 register operands, operand reuse possible)
__global__ void cc_fp32_tile_probe_kernel(float* sink, const float* in, int iters) {
  float acc[4][8];
  float x[4], y[8];
  for (int i = 0; i < 4; ++i) x[i] = in[threadIdx.x + i];
  for (int j = 0; j < 8; ++j) y[j] = in[threadIdx.x + 4 + j];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(x[i], y[j], acc[i][j]);
      x[u & 3] += 1e-7f;
      y[u & 7] -= 1e-7f;
    }
  }
  float s = 0.f;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 8; ++j) s += acc[i][j];
  if (s == 123.456f) sink[0] = s;
}
