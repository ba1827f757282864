/* cc_oracle.c — C restatement (fp32 and fp64, OpenMP over trajectory blocks) of chain_control's
 * neural-ODE unroll and its reverse pass.  TEST INFRASTRUCTURE: the checker for the CUDA path and
 * the CPU baseline of bench.py ("kind": "port").  Never linked into libccb200.so.
 *
 * PARITY PINNED BY REFERENCE EXECUTION: tests/test_golden.py requires the fp64 entry points of this
 * file to reproduce, to 1e-9, outputs / losses / gradients obtained by executing the reference's own
 * source files on a NumPy stand-in of jax/equinox (tests/golden/make_golden.py, refshim.py; fixtures
 * tests/golden/*.npz).  Also validated against oracle/np_oracle.py, torch autograd and finite
 * differences (tests/test_c_oracle.py, tests/test_oracle.py).
 *
 * Follows: cc/core/abstract.py:54-70,97-127; cc/examples/neural_ode_model.py:106-175;
 * neural_ode_controller.py:108-149; *_compact_example.py; nn_lib/integrate.py:1-28;
 * nn_lib/mlp_network.py:5-35; cc/utils/utils.py:53-55.  Like XLA's scan transpose the backward
 * stores every per-step residual of the forward (no rematerialisation).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cc_b200.h"

#define NB 16 /* trajectories per block = SIMD lanes */

int64_t cco_param_count(const CcModule* m) {
  int64_t p = 0;
  for (int which = 0; which < 2; ++which) {
    const CcMlp* mlp = which ? &m->g : &m->f;
    for (int l = 0; l < mlp->n_layers; ++l)
      p += (int64_t)mlp->sizes[l] * mlp->sizes[l + 1] + (mlp->use_bias ? mlp->sizes[l + 1] : 0);
  }
  return p;
}

int cco_max_threads(void) { return omp_get_max_threads(); }

#define REAL float
#define SFX _f32
#define TANH tanhf
#define ATAN atanf
#include "cc_oracle_impl.h"
#undef REAL
#undef SFX
#undef TANH
#undef ATAN

#define REAL double
#define SFX _f64
#define TANH tanh
#define ATAN atan
#include "cc_oracle_impl.h"
