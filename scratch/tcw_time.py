import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import np_oracle as O
from ccspecs import model_variants
from cchelpers import conv
from chain_control_b200 import ops
ms = model_variants()['compact_s50_d0']; spec = conv(ms); theta = torch.tensor(O.init_params(ms, 1, stable=True).astype(np.float32), device='cuda')
B, T = int(sys.argv[1]), 1000
u = torch.rand(B, T, 1, device='cuda') * 2 - 1; ybar = torch.randn(B, T + 1, 1, device='cuda') / (B * T)
y, _, tape = ops.rollout_fwd(spec, theta, u, want_tape=True)
for dbg in sys.argv[2:]:
    os.environ['CC_B200_TCW_DEBUG'] = dbg
    ops.rollout_bwd(spec, theta, u, tape, ybar); torch.cuda.synchronize(); ts = []
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ops.rollout_bwd(spec, theta, u, tape, ybar); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print('B=%d debug=%s: %.2f ms' % (B, dbg, min(ts)), flush=True)
