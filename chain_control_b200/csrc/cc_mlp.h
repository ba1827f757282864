// cc_mlp.h -- host/device interface of the deep-MLP tcgen05 forward family (cc_mlp_kernels.cuh, cc_mlp.cu).
#pragma once
#include <stddef.h>

#include "../../include/cc_b200.h"

struct CcMlpLayerArg {
  int w_off, b_off;  // offsets into theta (b_off < 0: no bias)
  int in_log, N;     // logical input / output width
  int act;
};
struct CcMlpArgs {
  int S, I, O, L;  // state / input / output dims, number of Linear layers of f (2..4)
  CcMlpLayerArg f[CC_MAX_LAYERS];
  CcMlpLayerArg g;  // readout Linear [O][S]
  int integrator, pre, ut;
  float u_scale, out_scale, dt;
  const float* theta;
  const float* u;   // [B][T][I]
  const float* x0;  // [B][S] or null
  float* y;         // [B][T+1][O]
  float* x_final;   // [B][S] or null
  float* tape;      // [n_ckpt][S][Bpad] pre-step states (the FFMA reverse pass reads them) or null
  float* ttab;      // [2][T] step times behind the tape
  float* pack;      // (widths 65..128) the pre-split weight chunks in the workspace, cc_mlp128_kernels.cuh
  int ckpt_every;
  long long B, Bpad;
  int T;
  // reverse pass (cc_mlpb_kernels.cuh)
  float* tape_xT;     // [S][Bpad] final state behind the step times (written by the forward kernels when set)
  const float* ybar;  // [B][T+1][O]
  float* recs;        // per-CTA sums and stage-input slots
  int flush_every;    // steps between flushes of the weight-gradient accumulators
  // segment replay (checkpointed recomputation, ckpt_every > 1): the kernels walk the steps [k0, k1) of the T-step problem
  // (u / y / ybar keep their full-length rows); k1 = 0 means the whole range
  int k0, k1;
  const float* carry_in;  // forward: [S][Bpad] state at step k0 (a checkpoint) instead of x0; null = x0
  int seg_tape;           // forward: the tape is segment-local (row k - k0, every step)
  int seg_last;           // reverse: this segment ends at step T (terminal readout term, adjoint starts at zero)
  int seg_accum;          // reverse: per-CTA sums already hold later segments' sums (do not zero)
  float* lam_io;          // reverse: [S][Bpad] adjoint carried between segments (read unless seg_last, always written), or null
};


size_t cc_mlp_smem_bytes(int pad, int L);
size_t cc_mlp_workspace_bytes(int pad, int L);
int cc_launch_mlp_fwd(const CcMlpArgs& a, int pad, void* stream, char* err, int errlen);
// reverse pass (widths <= 64): workspace = [2 L slabs][per-CTA records]; theta_bar gets all P parameters
size_t cc_mlpb_workspace_bytes(int L, long long B);
int cc_launch_mlpb_bwd(CcMlpArgs a, float* theta_bar, int P, void* ws, void* stream, char* err, int errlen, int phases = 3);
