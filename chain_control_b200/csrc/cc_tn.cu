// cc_tn.cu -- the rollout / BPTT kernels for ONE register-tile width (compile with -DCC_TN_VALUE=8|13|16;
// three translation units so that the build parallelises).  Exposes cc_launch_{fwd,bwd}_tn<TN>.
#ifndef CC_EMULATE
#include <cuda_runtime.h>
#endif
#include <cstdio>

#include "cc_kernels.cuh"
#include "cc_kernels_bwd.cuh"

#ifndef CC_TN_VALUE
#error "define CC_TN_VALUE"
#endif
#define CC_CAT2(a, b) a##b
#define CC_CAT(a, b) CC_CAT2(a, b)

namespace {
template <class Kern>
int CC_CAT(cc_launch1_tn, CC_TN_VALUE)(Kern kern, const CcKParams& p, void* stream, char* err, int errlen) {
  const long long grid = (p.n_wtiles + p.wpc - 1) / p.wpc;
  const size_t smem = ((size_t)p.cta_floats + (size_t)p.wpc * p.warp_floats + CC_PARAM_FLOATS) * 4;
#ifdef CC_EMULATE
  (void)stream; (void)err; (void)errlen;
  ccemu::launch((int)grid, p.wpc * 32, smem, [&]() { kern(p); });
  return CC_OK;
#else
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    kern<<<(unsigned)grid, p.wpc * 32, smem, (cudaStream_t)stream>>>(p);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "kernel launch failed: %s", cudaGetErrorString(e));
    return CC_ERR_CUDA;
  }
  return CC_OK;
#endif
}
}  // namespace

#define CC_DISPATCH_CM(kern)                                                                                   \
  (p.ctrl_mode == CC_CTRL_NONE ? CC_CAT(cc_launch1_tn, CC_TN_VALUE)(kern<CC_TN_VALUE, CC_CTRL_NONE>, p, stream, err, errlen) \
   : p.ctrl_mode == CC_CTRL_TINY ? CC_CAT(cc_launch1_tn, CC_TN_VALUE)(kern<CC_TN_VALUE, CC_CTRL_TINY>, p, stream, err, errlen) \
                                 : CC_CAT(cc_launch1_tn, CC_TN_VALUE)(kern<CC_TN_VALUE, CC_CTRL_TILE>, p, stream, err, errlen))

int CC_CAT(cc_launch_fwd_tn, CC_TN_VALUE)(const CcKParams& p, void* stream, char* err, int errlen) {
  return CC_DISPATCH_CM(cc_rollout_fwd_kernel);
}
int CC_CAT(cc_launch_bwd_tn, CC_TN_VALUE)(const CcKParams& p, void* stream, char* err, int errlen) {
  return CC_DISPATCH_CM(cc_rollout_bwd_kernel);
}
#undef CC_DISPATCH_CM
