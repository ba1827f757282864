// cc_tc.cu -- translation unit of the tcgen05 rollout kernels (cc_tc_kernels.cuh); exposes cc_launch_{fwd,bwd}_tc.
#include <cuda_runtime.h>
#include <cstdio>

#include "cc_tc_kernels.cuh"
#include "cc_wpt_kernels.cuh"
#include "cc_tcw_kernels.cuh"

namespace {
template <class Kern>
int launch_tc(Kern kern, const CcKParams& p, bool bwd, void* stream, char* err, int errlen) {
  const int SR = 8 * p.tc_nch;
  const long long grid = (p.B + 127) / 128;
  const size_t gp = bwd ? (size_t)(p.nd[0].S + p.nd[0].O) * CC_CKU * 128 + (size_t)(p.nd[0].f.layer[0].act != CC_ACT_IDENTITY ? 9 : 5) * CC_TINY_S * 128 : 0;  // lane-private controller gradient accumulators + stage record
  const size_t smem = ((size_t)p.tc_off + CC_TC_REGION_FLOATS(SR, p.tc_KP, p.tc_NP) + gp) * 4;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    kern<<<(unsigned)grid, CC_TC_THREADS, smem, (cudaStream_t)stream>>>(p);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "tcgen05 kernel launch failed: %s", cudaGetErrorString(e));
    return CC_ERR_CUDA;
  }
  return CC_OK;
}
}  // namespace

// CFG = 521: the compact controller's shape (S_c = 5, 2 inputs, 1 output, time-invariant) as compile-time constants
static int ctrl_cfg(const CcKParams& p) {
  const CcNodeDev& C = p.nd[0];
  return (p.n_nodes == 2 && C.S == 5 && C.I == 2 && C.O == 1 && !C.has_t && C.ut == CC_UT_IDENTITY &&
          C.f.layer[0].act == CC_ACT_IDENTITY && C.g.layer[0].act == CC_ACT_IDENTITY) ? 521 : 0;
}
#define CC_TC_FWD(NCH, CM, FACT) (ctrl_cfg(p) == 521 && CM == CC_CTRL_TINY ? launch_tc(cc_tc_fwd_kernel<NCH, CM, FACT, (CM == CC_CTRL_TINY ? 521 : 0)>, p, false, stream, err, errlen) \
                                                                           : launch_tc(cc_tc_fwd_kernel<NCH, CM, FACT, 0>, p, false, stream, err, errlen))
// ---- family 2: latency kernels, one warp per trajectory (cc_wpt_kernels.cuh) ----
template <class Kern>
static int launch_wpt(Kern kern, const CcKParams& p, bool bwd, void* stream, char* err, int errlen) {
  const long long grid = (p.B + 3) / 4;
  const size_t extra = bwd ? (size_t)(p.nd[0].S + p.nd[0].O) * CC_CKU * 128 + (size_t)(p.nd[0].f.layer[0].act != CC_ACT_IDENTITY ? 9 : 5) * CC_TINY_S * 128 : 0;
  const size_t smem = ((size_t)((CC_PARAM_FLOATS + 15) / 16 * 16) + CC_WPT_REGION_FLOATS + extra) * 4;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    kern<<<(unsigned)grid, CC_WPT_THREADS, smem, (cudaStream_t)stream>>>(p);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "latency kernel launch failed: %s", cudaGetErrorString(e));
    return CC_ERR_CUDA;
  }
  return CC_OK;
}

int cc_launch_fwd_tc(const CcKParams& p, void* stream, char* err, int errlen) {
  if (p.tc == 2) {
    const bool c521 = ctrl_cfg(p) == 521;
    if (p.n_nodes == 1) return launch_wpt(cc_wpt_fwd_kernel<CC_CTRL_NONE, 0>, p, false, stream, err, errlen);
    return c521 ? launch_wpt(cc_wpt_fwd_kernel<CC_CTRL_TINY, 521>, p, false, stream, err, errlen) : launch_wpt(cc_wpt_fwd_kernel<CC_CTRL_TINY, 0>, p, false, stream, err, errlen);
  }
  const bool closed = p.n_nodes == 2;
  const bool fact = p.nd[p.n_nodes - 1].f.layer[0].act != CC_ACT_IDENTITY;  // non-linear final activation of f: separate instantiation
  if (p.tc_nch == 4) {
    if (fact) return closed ? CC_TC_FWD(4, CC_CTRL_TINY, true) : CC_TC_FWD(4, CC_CTRL_NONE, true);
    return closed ? CC_TC_FWD(4, CC_CTRL_TINY, false) : CC_TC_FWD(4, CC_CTRL_NONE, false);
  }
  if (fact) return closed ? CC_TC_FWD(7, CC_CTRL_TINY, true) : CC_TC_FWD(7, CC_CTRL_NONE, true);
  return closed ? CC_TC_FWD(7, CC_CTRL_TINY, false) : CC_TC_FWD(7, CC_CTRL_NONE, false);
}
// open-loop reverse pass with a trainable depth-0 model (latency kernels)
static int launch_wpt_obwd(const CcKParams& p, void* stream, char* err, int errlen) {
  const long long grid = (p.B + 3) / 4;
  const size_t smem = ((size_t)((CC_PARAM_FLOATS + 15) / 16 * 16) + CC_WPT_OB_FLOATS(p.nd[0].S) + 4 * CC_WPT_ZB) * 4;
  auto go = [&](auto kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      kern<<<(unsigned)grid, CC_WPT_THREADS, smem, (cudaStream_t)stream>>>(p);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
      snprintf(err, errlen, "latency kernel launch failed: %s", cudaGetErrorString(e));
      return (int)CC_ERR_CUDA;
    }
    return (int)CC_OK;
  };
  return p.nd[0].K0 <= 52 ? go(cc_wpt_obwd_kernel<52>) : go(cc_wpt_obwd_kernel<56>);
}

// open-loop reverse pass with a trainable linear depth-0 model at large batch: tcgen05 incl. the weight gradient (cc_tcw_kernels.cuh)
static int launch_tcw(const CcKParams& p, void* stream, char* err, int errlen) {
  const int SR = 8 * p.tc_nch;
  const long long grid = (p.B + 127) / 128;
  const size_t smem = ((size_t)p.tc_off + CC_TCW_REGION_FLOATS(SR)) * 4;
  auto go = [&](auto kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      kern<<<(unsigned)grid, CC_TCW_THREADS, smem, (cudaStream_t)stream>>>(p);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
      snprintf(err, errlen, "tcgen05 weight-gradient kernel launch failed: %s", cudaGetErrorString(e));
      return (int)CC_ERR_CUDA;
    }
    return (int)CC_OK;
  };
  return p.tc_nch == 4 ? go(cc_tcw_bwd_kernel<4>) : go(cc_tcw_bwd_kernel<7>);
}

int cc_launch_bwd_tc(const CcKParams& p, void* stream, char* err, int errlen) {
  if (p.tc == 3) return launch_tcw(p, stream, err, errlen);
  if (p.tc == 2 && p.n_nodes == 1) return launch_wpt_obwd(p, stream, err, errlen);
  if (p.tc == 2)
    return ctrl_cfg(p) == 521 ? launch_wpt(cc_wpt_bwd_kernel<CC_CTRL_TINY, 521>, p, true, stream, err, errlen) : launch_wpt(cc_wpt_bwd_kernel<CC_CTRL_TINY, 0>, p, true, stream, err, errlen);
  const bool c521 = ctrl_cfg(p) == 521;
  if (p.tc_nch == 4) return c521 ? launch_tc(cc_tc_bwd_kernel<4, CC_CTRL_TINY, 521>, p, true, stream, err, errlen) : launch_tc(cc_tc_bwd_kernel<4, CC_CTRL_TINY, 0>, p, true, stream, err, errlen);
  return c521 ? launch_tc(cc_tc_bwd_kernel<7, CC_CTRL_TINY, 521>, p, true, stream, err, errlen) : launch_tc(cc_tc_bwd_kernel<7, CC_CTRL_TINY, 0>, p, true, stream, err, errlen);
}
