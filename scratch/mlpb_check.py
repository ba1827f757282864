"""GPU check of the deep-MLP tcgen05 reverse pass (cc_mlpb_kernels.cuh) against the fp64 C oracle + timing."""
import os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
os.environ.setdefault('CC_B200_VERBOSE', '1')
import numpy as np, torch
from cchelpers import conv, l2rel, normwise
from oracle import np_oracle as O, c_oracle as CO
from chain_control_b200 import ops
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device='cuda')
cases = {
  'A': (O.make_node(40, 1, 1, f_width=48, f_depth=2), 301, 57),
  'B': (O.make_node(64, 2, 2, f_width=64, f_depth=1, f_act=O.ACT_TANH, f_act_final=O.ACT_TANH, readout_pre_step=True, u_transform=O.UT_ARCTAN), 257, 40),
  'C': (O.make_node(64, 1, 1, f_width=64, f_depth=2, f_act_final=O.ACT_TANH), 257, 120),
  'D': (O.make_node(20, 1, 1, f_width=33, f_depth=2, integrator=O.INT_EE, f_act=O.ACT_TANH), 130, 33),
}
for name in (sys.argv[1:] or list(cases)):
    spec, B, T = cases[name]
    theta = O.init_params(spec, 1).astype(np.float32)
    rng = np.random.default_rng(0)
    u = rng.uniform(-1, 1, size=(B, T, spec.input_dim)).astype(np.float32)
    x0 = (0.1 * rng.normal(size=(B, spec.state_dim))).astype(np.float32)
    ybar = rng.normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    ref = CO.rollout(spec, theta, u, x0=x0, ybar=ybar, want_ubar=False, dtype=np.float64)
    y, _, tape = ops.rollout_fwd(conv(spec), T_(theta), T_(u), T_(x0), want_tape=True)
    tb, _ = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar), T_(x0))
    torch.cuda.synchronize()
    print(f'case {name}: y err {normwise(y.cpu().numpy(), ref["y"]):.2e}  theta_bar l2rel {l2rel(tb.cpu().numpy(), ref["theta_bar"]):.2e}', flush=True)
