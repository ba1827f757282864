// cc_mlp.cu -- translation unit of the deep-MLP tcgen05 forward family (cc_mlp_kernels.cuh); exposes cc_launch_mlp_fwd.
// Also compiled (included by cc_abi.cu) into the CPU emulation of tests/emu.
#ifndef CC_EMULATE
#include <cuda_runtime.h>
#endif
#include <cstdio>

#include <cstdlib>

#include "cc_mlp128_kernels.cuh"
#include "cc_mlp2_kernels.cuh"
#include "cc_mlpb_kernels.cuh"

namespace {
template <int PAD, bool SEG>
int mlp_launch_t(const CcMlpArgs& a, void* stream, char* err, int errlen) {
  using TM = CcMlpTm<PAD>;
  const long long grid = (a.B + 127) / 128;
#ifdef CC_EMULATE
  (void)stream; (void)err; (void)errlen;
  const size_t smem = ((size_t)CC_MLP_SMEM_FLOATS(PAD, a.L) + (size_t)TM::COLS * 128) * 4;
  ccemu::launch((int)grid, CC_MLP_THREADS, smem, [&]() { cc_mlp_fwd_kernel<PAD, SEG>(a); });
  return CC_OK;
#else
  const size_t smem = (size_t)CC_MLP_SMEM_FLOATS(PAD, a.L) * 4;
  cudaError_t e = cudaFuncSetAttribute(cc_mlp_fwd_kernel<PAD, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    cc_mlp_fwd_kernel<PAD, SEG><<<(unsigned)grid, CC_MLP_THREADS, smem, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "cc_mlp_fwd_kernel<%d> launch failed: %s", PAD, cudaGetErrorString(e));
    return CC_ERR_CUDA;
  }
  return CC_OK;
#endif
}
template <int PAD, bool SEG>
int mlp2_launch_t(const CcMlpArgs& a, void* stream, char* err, int errlen) {
  using TM = CcMlpTm<PAD>;
  const long long grid = (a.B + 127) / 128;
#ifdef CC_EMULATE
  (void)stream; (void)err; (void)errlen;
  const size_t smem = ((size_t)CC_MLP_SMEM_FLOATS(PAD, a.L) + (size_t)(TM::COLS + 8) * 128) * 4;
  ccemu::launch((int)grid, CC_MLP2_THREADS, smem, [&]() { cc_mlp2_fwd_kernel<PAD, SEG>(a); });
  return CC_OK;
#else
  const size_t smem = (size_t)CC_MLP_SMEM_FLOATS(PAD, a.L) * 4;
  cudaError_t e = cudaFuncSetAttribute(cc_mlp2_fwd_kernel<PAD, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    cc_mlp2_fwd_kernel<PAD, SEG><<<(unsigned)grid, CC_MLP2_THREADS, smem, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "cc_mlp2_fwd_kernel<%d> launch failed: %s", PAD, cudaGetErrorString(e));
    return CC_ERR_CUDA;
  }
  return CC_OK;
#endif
}

// segment replay (k1 > 0) takes the SEG instantiation; the plain forward pass the other one
template <int PAD>
int mlp_launch(const CcMlpArgs& a, void* stream, char* err, int errlen) {
  return a.k1 > 0 ? mlp_launch_t<PAD, true>(a, stream, err, errlen) : mlp_launch_t<PAD, false>(a, stream, err, errlen);
}
template <int PAD>
int mlp2_launch(const CcMlpArgs& a, void* stream, char* err, int errlen) {
  return a.k1 > 0 ? mlp2_launch_t<PAD, true>(a, stream, err, errlen) : mlp2_launch_t<PAD, false>(a, stream, err, errlen);
}

int mlp128_launch(const CcMlpArgs& a, void* stream, char* err, int errlen) {
  const long long grid = (a.B + 127) / 128;
  if (!a.pack) {
    snprintf(err, errlen, "deep-MLP tcgen05 family (width > 64): workspace missing (cc_rollout_fwd_workspace_bytes)");
    return CC_ERR_WORKSPACE;
  }
#ifdef CC_EMULATE
  (void)stream;
  ccemu::launch(2 * a.L, 256, 0, [&]() { cc_mlp128_prep_kernel(a); });
  const size_t smem = ((size_t)CC_MLP128_SMEM_FLOATS + (size_t)CcMlpTm<128>::COLS * 128) * 4;
  ccemu::launch((int)grid, CC_MLP128_WORKERS, smem, [&]() { cc_mlp128_fwd_kernel(a); });
  return CC_OK;
#else
  const size_t smem = (size_t)CC_MLP128_SMEM_FLOATS * 4;
  cc_mlp128_prep_kernel<<<2 * a.L, 256, 0, (cudaStream_t)stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaFuncSetAttribute(cc_mlp128_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    cc_mlp128_fwd_kernel<<<(unsigned)grid, CC_MLP128_THREADS, smem, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "cc_mlp128_fwd_kernel launch failed: %s", cudaGetErrorString(e));
    return CC_ERR_CUDA;
  }
  return CC_OK;
#endif
}
template <int L, int NH>
int mlpb_launch(const CcMlpArgs& a, void* stream, char* err, int errlen) {
  const long long grid = (a.B + 127) / 128;
#ifdef CC_EMULATE
  (void)stream; (void)err; (void)errlen;
  const size_t smem = ((size_t)CC_MLPB_SMEM_FLOATS + 16 + (size_t)512 * 128) * 4;
  ccemu::launch((int)grid, 128 * NH, smem, [&]() { cc_mlpb_bwd_kernel<L, NH>(a); });
  return CC_OK;
#else
  const size_t smem = (size_t)CC_MLPB_SMEM_FLOATS * 4;
  cudaError_t e = cudaFuncSetAttribute(cc_mlpb_bwd_kernel<L, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    cc_mlpb_bwd_kernel<L, NH><<<(unsigned)grid, 128 * NH + 128, smem, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "cc_mlpb_bwd_kernel<%d, %d> launch failed: %s", L, NH, cudaGetErrorString(e));
    return CC_ERR_CUDA;
  }
  return CC_OK;
#endif
}
}  // namespace

size_t cc_mlpb_workspace_bytes(int L, long long B) {
  return ((size_t)2 * L * CC_MLPB_SLAB_FLOATS + (size_t)((B + 127) / 128) * CC_MLPB_CTA_FLOATS(L)) * 4;
}

// prep (operand slabs) -> reverse kernel (per-CTA sums) -> reduction into theta_bar.  Segment replay (ckpt_every > 1) calls it
// once per segment, last segment first: `phases` bit 0 = run the preparation (first call), bit 1 = run the reduction (last call)
int cc_launch_mlpb_bwd(CcMlpArgs a, float* theta_bar, int P, void* ws, void* stream, char* err, int errlen, int phases) {
  a.pack = (float*)ws;
  a.recs = a.pack + (size_t)2 * a.L * CC_MLPB_SLAB_FLOATS;
  const char* fe = getenv("CC_B200_MLPB_FLUSH");
  a.flush_every = fe && atoi(fe) >= 1 ? atoi(fe) : 2;
  const int n_cta = (int)((a.B + 127) / 128);
  if (phases & 1) {
#ifdef CC_EMULATE
    ccemu::launch(2 * a.L, 256, 0, [&]() { cc_mlpb_prep_kernel(a); });
#else
    cc_mlpb_prep_kernel<<<2 * a.L, 256, 0, (cudaStream_t)stream>>>(a);
    if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) {
      snprintf(err, errlen, "cc_mlpb_prep_kernel launch failed: %s", cudaGetErrorString(e));
      return CC_ERR_CUDA;
    }
#endif
  }
  // worker threads per trajectory: 2 (32 columns each).  CC_B200_MLPB_NH=4 selects the 16-column variant (640 threads):
  // measured 9 % slower at (64,64,2) x 65,536 (10.06 vs 9.15 ms per 32 steps) -- kept because the tests run both.
  const char* nhe = getenv("CC_B200_MLPB_NH");
  const int nh = nhe && atoi(nhe) == 4 ? 4 : 2;
  int rc;
  if (a.L == 2) rc = nh == 2 ? mlpb_launch<2, 2>(a, stream, err, errlen) : mlpb_launch<2, 4>(a, stream, err, errlen);
  else if (a.L == 3) rc = nh == 2 ? mlpb_launch<3, 2>(a, stream, err, errlen) : mlpb_launch<3, 4>(a, stream, err, errlen);
  else { snprintf(err, errlen, "deep-MLP tcgen05 reverse pass: %d layers not built", a.L); return CC_ERR_UNSUPPORTED; }
  if (rc) return rc;
  if (!(phases & 2)) return CC_OK;
#ifdef CC_EMULATE
  ccemu::launch((P + 127) / 128, 128, 0, [&]() { cc_mlpb_reduce_kernel(a, theta_bar, n_cta, P); });
#else
  cc_mlpb_reduce_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a, theta_bar, n_cta, P);
  if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) {
    snprintf(err, errlen, "cc_mlpb_reduce_kernel launch failed: %s", cudaGetErrorString(e));
    return CC_ERR_CUDA;
  }
#endif
  return CC_OK;
}

// shared memory of one CTA (bytes) for padded width `pad` and `L` layers: the plan checks it against the 227 KB budget
size_t cc_mlp_smem_bytes(int pad, int L) {
  if (pad == 128) return (size_t)CC_MLP128_SMEM_FLOATS * 4;  // (streamed weights: independent of L)
  const size_t kp = (size_t)pad + 8;
  return ((size_t)L * 2 * kp * pad + (size_t)CC_TINY_O * (pad + 4) + 8) * 4;
}

// global workspace: the pre-split weight chunks of the streamed (width 65..128) variant
size_t cc_mlp_workspace_bytes(int pad, int L) { return pad == 128 ? (size_t)2 * L * CC_MLP128_CHUNK_BYTES : 0; }

int cc_launch_mlp_fwd(const CcMlpArgs& a, int pad, void* stream, char* err, int errlen) {
  // two threads per trajectory (cc_mlp2_kernels.cuh): measured on B200 (gpurun_out/r02e_sweep_nh*.jsonl) +4 % at large batch and
  // +8 % in the one-wave regime for the 64-wide variant; the 32-wide one-thread kernel keeps 3 CTAs per SM and stays ahead
  // (72 vs 62 TFLOP/s algorithmic).  CC_B200_MLP_NH=1 / 2 forces a variant (tests run both).
  const char* nh = getenv("CC_B200_MLP_NH");
  const int force = nh ? atoi(nh) : 0;
  switch (pad) {
    case 16: return mlp_launch<16>(a, stream, err, errlen);
    case 32: return force == 2 ? mlp2_launch<32>(a, stream, err, errlen) : mlp_launch<32>(a, stream, err, errlen);
    case 64: return force == 1 ? mlp_launch<64>(a, stream, err, errlen) : mlp2_launch<64>(a, stream, err, errlen);
    case 128: return mlp128_launch(a, stream, err, errlen);
    default: snprintf(err, errlen, "deep-MLP tcgen05 family: padded width %d not built", pad); return CC_ERR_UNSUPPORTED;
  }
}
