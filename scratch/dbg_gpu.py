import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from cchelpers import conv, l2rel, normwise
from ccspecs import controller_variants, model_variants
from oracle import c_oracle as CO, np_oracle as O
from chain_control_b200 import ops
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device='cuda')
mspec, cspec = model_variants()['full_s5_w10_d2'], controller_variants()['full_s6_ft_tv']
for (B, Tr, use_noise) in [(6, 19, True), (290, 201, True), (290, 201, False), (32, 201, True), (290, 40, True)]:
    thc, thm = O.init_params(cspec, 2).astype(np.float32), O.init_params(mspec, 1).astype(np.float32)
    refs = 3 * np.random.default_rng(0).uniform(-1, 1, size=(B, Tr, 1)).astype(np.float32)
    noise = (0.02 * np.random.default_rng(1).normal(size=(B, Tr, 1))).astype(np.float32) if use_noise else None
    ybar = np.random.default_rng(6).normal(size=(B, Tr, 1)).astype(np.float32)
    ref = CO.closedloop(cspec, mspec, thc, thm, refs, noise, ybar=ybar, dtype=np.float64)
    ref32 = CO.closedloop(cspec, mspec, thc, thm, refs, noise, ybar=ybar, dtype=np.float32)
    yh, us, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), None if noise is None else T_(noise), want_tape=True, want_u=True)
    tb = ops.closedloop_bwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), tape, T_(ybar), None if noise is None else T_(noise))
    print(B, Tr, use_noise, 'y', normwise(yh.cpu().numpy(), ref['yhat']), 'u', normwise(us.cpu().numpy(), ref['u']), 'grad', l2rel(tb.cpu().numpy(), ref['theta_c_bar']), '| f32 oracle vs f64: y', normwise(ref32['yhat'], ref['yhat']), 'grad', l2rel(ref32['theta_c_bar'], ref['theta_c_bar']))
