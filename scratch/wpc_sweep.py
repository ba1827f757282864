import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import np_oracle as O
from ccspecs import model_variants, controller_variants
from cchelpers import conv
from chain_control_b200 import ops
dev = 'cuda'
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device=dev)
def timeit(fn, n=2):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
ms_ = model_variants()['compact_s50_d0']; cs_ = controller_variants()['compact_s5_d0']
spec = conv(ms_); cs = conv(cs_)
theta = T_(O.init_params(ms_, 1, stable=True)); thc = T_(O.init_params(cs_, 2))
T = 200
for wpc in [1, 2, 3, 4, 5, 6, 7, 8]:
    os.environ['CC_B200_WPC'] = str(wpc)
    B = 148 * wpc * 32   # exactly one CTA per SM, one wave
    u = torch.rand(B, T, 1, device=dev) * 2 - 1
    try:
        f0 = timeit(lambda: ops.rollout_fwd(spec, theta, u))
        c0 = timeit(lambda: ops.closedloop_fwd(cs, spec, thc, theta, u))
        print('wpc=%d B=%d open fwd %.2f ms (%.2f us/step/CTA) -> %.3e steps/s | closed fwd %.2f ms -> %.3e' % (wpc, B, f0, f0 / T * 1e3, B * T / f0 * 1e3, c0, B * T / c0 * 1e3))
    except Exception as e:
        print('wpc', wpc, 'failed', str(e)[:100])
