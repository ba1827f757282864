// cc_lin.cu -- translation unit of the blocked linear family (cc_lin_kernels.cuh, cc_lin_prep.cuh); exposes
// cc_launch_lin_prep / cc_launch_lin.  Also compiled (included by cc_abi.cu) into the CPU emulation of tests/emu.
#ifndef CC_EMULATE
#include <cuda_runtime.h>
#endif
#include <cstdio>

#include "cc_lin_kernels.cuh"
#ifdef CC_LIN_OPEN
#include "cc_lin_open.cuh"
#endif
#include "cc_lin_prep.cuh"

namespace {
template <class Kern, class Arg>
int lin_launch1(Kern kern, const Arg& a, long long grid, int threads, size_t smem, void* stream, char* err, int errlen, const char* what) {
#ifdef CC_EMULATE
  (void)stream; (void)err; (void)errlen; (void)what;
  ccemu::launch((int)grid, threads, smem, [&]() { kern(a); });
  return CC_OK;
#else
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    kern<<<(unsigned)grid, threads, smem, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "%s launch failed: %s", what, cudaGetErrorString(e));
    return CC_ERR_CUDA;
  }
  return CC_OK;
#endif
}
#ifdef CC_EMULATE
constexpr int kEmuTmemFloats = 192 * 128;
#else
constexpr int kEmuTmemFloats = 0;
#endif
int ctrl_cfg_lin(const CcKParams& p) {
  const CcNodeDev& C = p.nd[0];
  return (p.n_nodes == 2 && C.S == 5 && C.I == 2 && C.O == 1 && !C.has_t && C.ut == CC_UT_IDENTITY &&
          C.f.layer[0].act == CC_ACT_IDENTITY && C.g.layer[0].act == CC_ACT_IDENTITY) ? 521 : 0;
}
}  // namespace

int cc_launch_lin_prep(const CcNodeDev& M, const float* theta, float* pack, double* tab, int m, int form, const CcNodeDev* C, const float* theta_c,
                       void* stream, char* err, int errlen) {
  CcLinPrepArgs a;
  a.M = M; a.theta = theta; a.pack = pack; a.tab = tab; a.m = m; a.form = form;
  a.ctrl = C ? 1 : 0;
  a.C = C ? *C : M;
  a.theta_c = theta_c;
  return lin_launch1(cc_lin_prep_kernel, a, 1, CC_LIN_PREP_THREADS, (size_t)CC_LIN_PREP_SMEM_DOUBLES * 8, stream, err, errlen, "lin prep kernel");
}

#ifdef CC_LIN_OPEN
int cc_launch_lin_grad(const CcLinGradArgs& a, void* stream, char* err, int errlen) {
  int rc = lin_launch1(cc_lin_qreduce_kernel, a, 64, 64, 0, stream, err, errlen, "lin Q reduction kernel");
  if (rc) return rc;
  rc = lin_launch1(cc_lin_grad_kernel, a, a.m, CC_LIN_PREP_THREADS, (size_t)CC_LIN_GRAD_SMEM_DOUBLES * 8, stream, err, errlen, "lin gradient kernel");
  if (rc) return rc;
  return lin_launch1(cc_lin_grad_final_kernel, a, 1, CC_LIN_PREP_THREADS, (size_t)CC_LIN_GRADF_SMEM_DOUBLES * 8, stream, err, errlen, "lin gradient kernel (final)");
}
#endif

int cc_launch_lin_cgrad(const CcLinCgradArgs& a, void* stream, char* err, int errlen) {
  return lin_launch1(cc_lin_cgrad_kernel, a, 1, 64, 8 * 64 * 8, stream, err, errlen, "lin controller gradient kernel");
}

#define CC_LIN_NCH(kern, ...)                                                                             \
  (nch == 4 ? lin_launch1(kern<4, __VA_ARGS__>, p, grid, CC_LIN_THREADS, smem, stream, err, errlen, #kern) \
   : nch == 7 ? lin_launch1(kern<7, __VA_ARGS__>, p, grid, CC_LIN_THREADS, smem, stream, err, errlen, #kern) \
              : lin_launch1(kern<8, __VA_ARGS__>, p, grid, CC_LIN_THREADS, smem, stream, err, errlen, #kern))

int cc_launch_lin(const CcKParams& p, bool bwd, void* stream, char* err, int errlen) {
  const long long grid = (p.B + 127) / 128;
  const int S = p.nd[p.n_nodes - 1].S;
  const int nch = S <= 32 ? 4 : (S <= 56 ? 7 : 8);
  if (p.n_nodes == 2) {
    const bool c521 = ctrl_cfg_lin(p) == 521;
    const bool stg = p.lin_stage != 0;
    const bool lc = p.lin_lc != 0;  // (the 521 shape is linear by construction: it is only instantiated together with LC)
    if (!bwd) {
      const size_t smem = ((size_t)CC_LIN_SM_STAGE + (stg ? 4 * 3 * CC_LIN_TILE : 0) + kEmuTmemFloats) * 4;
      if (stg) return (lc && c521) ? CC_LIN_NCH(cc_lin_cl_fwd_kernel, 521, true, true) : (lc ? CC_LIN_NCH(cc_lin_cl_fwd_kernel, 0, true, true) : CC_LIN_NCH(cc_lin_cl_fwd_kernel, 0, true, false));
      return (lc && c521) ? CC_LIN_NCH(cc_lin_cl_fwd_kernel, 521, false, true) : (lc ? CC_LIN_NCH(cc_lin_cl_fwd_kernel, 0, false, true) : CC_LIN_NCH(cc_lin_cl_fwd_kernel, 0, false, false));
    }
    const CcNodeDev& C = p.nd[0];
    const size_t gpf = (lc && c521) ? 0 : (size_t)(C.S + C.O) * CC_CKU * 128;
    const size_t keepf = lc ? 0 : (size_t)(C.f.layer[0].act != CC_ACT_IDENTITY ? 9 : 5) * CC_TINY_S * 128;
    const char* sgb = getenv("CC_B200_LIN_STAGE_BWD");
    const bool stgb = stg && lc && c521 && sgb && atoi(sgb) == 1;  // measured: no gain (2.55 vs 2.41 ms at 65,536 x 1001), off by default  // register accumulators leave room for the refs / ybar tiles
    const size_t ringf = 4 * CC_LIN_RING_D * 32 * CC_LIN_RING_STRIDE;  // per-warp cp.async ring of the tape / reference / cotangent loads
    const size_t smem = ((size_t)CC_LIN_SM_STAGE + (stgb ? 4 * 4 * CC_LIN_TILE : 0) + gpf + keepf + ringf + kEmuTmemFloats) * 4;
    if (stgb) return CC_LIN_NCH(cc_lin_cl_bwd_kernel, 521, true, true);
    return (lc && c521) ? CC_LIN_NCH(cc_lin_cl_bwd_kernel, 521, false, true) : (lc ? CC_LIN_NCH(cc_lin_cl_bwd_kernel, 0, false, true) : CC_LIN_NCH(cc_lin_cl_bwd_kernel, 0, false, false));
  }
#ifdef CC_LIN_OPEN
  // open loop
  if (!bwd) {
    const size_t smem = ((size_t)CC_LIN_SM_STAGE + 4 * 4 * CC_LIN_TILE + kEmuTmemFloats) * 4;
    return CC_LIN_NCH(cc_lin_ol_fwd_kernel, 0);
  }
  const size_t smem = ((size_t)CC_LINW_SMEM_FLOATS + kEmuTmemFloats * 2) * 4;
  return lin_launch1(cc_linw_bwd_kernel, p, grid, CC_LIN_THREADS, smem, stream, err, errlen, "cc_linw_bwd_kernel");
#else
  snprintf(err, errlen, "open-loop lin kernels not built");
  return CC_ERR_UNSUPPORTED;
#endif
}
