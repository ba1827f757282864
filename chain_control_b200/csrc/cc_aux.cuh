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

// theta_bar[j] = sum over warp tiles (fixed order) of the gradient-record entry that maps to parameter j.
// A CTA owns 32 consecutive record entries (one coalesced 128 B line per record) x CC_RED_SLICES slices of the tile
// range; each slice keeps 4 running sums, the slices are combined in a fixed order through shared memory -> 64 short
// chains per output instead of one n_tiles-long one (rounding error and latency both drop).
#define CC_RED_SLICES 16
__global__ void __launch_bounds__(32 * CC_RED_SLICES) cc_reduce_grad_kernel(const CcKParams p, int node, float* theta_bar, int n_tiles) {
#ifndef CC_EMULATE
  __shared__ float red[CC_RED_SLICES][33];
#else
  float(*red)[33] = reinterpret_cast<float(*)[33]>(ccemu::st().smem);
#endif
  const CcNodeDev& nd = p.nd[node];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + lane;  // index inside the node's record
  int dst = -1;
  if (idx < nd.g_total) {
    for (int w = 0; w < 2; ++w) {
      const CcMlpDev& mlp = w ? nd.g : nd.f;
      for (int l = 0; l < mlp.L; ++l) {
        const CcLayerDev& L = mlp.layer[l];
        if (idx >= L.g_off && idx < L.g_off + L.N * L.K) {
          const int n = (idx - L.g_off) / L.K, r = (idx - L.g_off) % L.K;
          const int lc = cc_row_to_logical(nd, L, r);
          if (lc >= 0) dst = L.w_off + n * L.in_log + lc;
          else if (lc == -2 && L.b_off >= 0) dst = L.b_off + n;
        }
      }
    }
  }
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (dst >= 0) {
    const float* src = p.gpart + nd.g_base + idx;
    const long long rec = p.g_rec;
    int t = slice;
    for (; t + 3 * CC_RED_SLICES < n_tiles; t += 4 * CC_RED_SLICES) {
      s0 += src[(long long)t * rec];
      s1 += src[(long long)(t + CC_RED_SLICES) * rec];
      s2 += src[(long long)(t + 2 * CC_RED_SLICES) * rec];
      s3 += src[(long long)(t + 3 * CC_RED_SLICES) * rec];
    }
    for (; t < n_tiles; t += CC_RED_SLICES) s0 += src[(long long)t * rec];
  }
  red[slice][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (slice == 0 && dst >= 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CC_RED_SLICES; ++i) s += red[i][lane];
    theta_bar[dst] = s;
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
