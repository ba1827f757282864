"""GPU: tcgen05 weight-gradient reverse pass (cc_tcw_kernels.cuh) vs the fp64 oracle and the FFMA kernels; timing (scratch)."""
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import np_oracle as O, c_oracle as CO
from ccspecs import model_variants
from cchelpers import conv
from chain_control_b200 import ops, _lib
dev = 'cuda'
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device=dev)
def nerr(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
os.environ['CC_B200_TC'] = '1'; os.environ['CC_B200_WPT_MAXB'] = '0'
variants = {'compact_s50_d0': model_variants()['compact_s50_d0'],
            'compact_s20_i2_o3': O.make_node(20, 2, 3, f_depth=0, readout_pre_step=True, u_transform=O.UT_ARCTAN),
            'compact_s51_nobias': O.make_node(51, 1, 1, f_depth=0, readout_pre_step=True, f_use_bias=False, g_use_bias=False)}
for name, ms in variants.items():
    for (B, T) in ((5, 1), (130, 2), (300, 37)):
        th = O.init_params(ms, 1, stable=True).astype(np.float32)
        I = ms.input_dim
        u = np.random.default_rng(0).uniform(-1, 1, (B, T, I)).astype(np.float32)
        ref0 = CO.rollout(ms, th, u, dtype=np.float64)
        ybar = np.random.default_rng(1).normal(size=ref0['y'].shape).astype(np.float32)
        ref = CO.rollout(ms, th, u, ybar=ybar, want_ubar=True, dtype=np.float64)
        out = {}
        for tcw in ('1', '0'):
            os.environ['CC_B200_TCW'] = tcw
            y, _, tape = ops.rollout_fwd(conv(ms), T_(th), T_(u), want_tape=True)
            tb, ub = ops.rollout_bwd(conv(ms), T_(th), T_(u), tape, T_(ybar), want_ubar=True)
            torch.cuda.synchronize()
            out[tcw] = (tb.cpu().numpy(), ub.cpu().numpy())
        print('%-20s B=%4d T=%3d  tcw: grad %.2e ubar %.2e | ffma: grad %.2e ubar %.2e' % (name, B, T,
              nerr(out['1'][0], ref['theta_bar']), nerr(out['1'][1], ref['ubar']), nerr(out['0'][0], ref['theta_bar']), nerr(out['0'][1], ref['ubar'])), flush=True)
if len(sys.argv) > 1:
    def timeit(fn, n=3):
        fn(); torch.cuda.synchronize(); ts = []
        for _ in range(n):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        return min(ts)
    ms = model_variants()['compact_s50_d0']; spec = conv(ms); theta = T_(O.init_params(ms, 1, stable=True))
    os.environ.pop('CC_B200_WPT_MAXB')
    for B in (16384, 65536):
        T = 1000
        u = torch.rand(B, T, 1, device=dev) * 2 - 1
        ybar = torch.randn(B, T + 1, 1, device=dev) / (B * T)
        y, _, tape = ops.rollout_fwd(spec, theta, u, want_tape=True)
        res = {}
        for tcw in ('1', '0'):
            os.environ['CC_B200_TCW'] = tcw
            ms_ = timeit(lambda: ops.rollout_bwd(spec, theta, u, tape, ybar))
            res[tcw] = ops.rollout_bwd(spec, theta, u, tape, ybar)[0]
            print('B=%6d T=%d reverse pass tcw=%s: %.2f ms' % (B, T, tcw, ms_), flush=True)
        print('   tcw vs ffma gradient: %.2e' % nerr(res['1'].cpu().numpy(), res['0'].cpu().numpy()))
