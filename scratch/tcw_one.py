import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import np_oracle as O
from cchelpers import conv
from chain_control_b200 import ops
os.environ['CC_B200_TC'] = sys.argv[2]; os.environ['CC_B200_WPT_MAXB'] = '0'; os.environ['CC_B200_TCW'] = sys.argv[1]
ms = O.make_node(20, 2, 3, f_depth=0, readout_pre_step=True, u_transform=O.UT_ARCTAN)
th = torch.tensor(O.init_params(ms, 1, stable=True).astype(np.float32), device='cuda')
T = int(sys.argv[3]); u = torch.rand(5, T, 2, device='cuda'); ybar = torch.randn(5, T + 1, 3, device='cuda')
y, _, tape = ops.rollout_fwd(conv(ms), th, u, want_tape=True)
torch.cuda.synchronize(); print('fwd ok', flush=True)
tb, ub = ops.rollout_bwd(conv(ms), th, u, tape, ybar, want_ubar=True)
torch.cuda.synchronize(); print('bwd ok', tb[:4], flush=True)
