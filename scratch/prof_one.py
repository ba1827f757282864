"""one launch of each hot kernel at B=65536 for ncu (scratch)."""
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import np_oracle as O
from ccspecs import model_variants, controller_variants
from cchelpers import conv
from chain_control_b200 import ops
dev = 'cuda'
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device=dev)
which = sys.argv[1]; T = int(sys.argv[2]); B = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
ms_ = model_variants()['compact_s50_d0']; cs_ = controller_variants()['compact_s5_d0']
spec = conv(ms_); cs = conv(cs_)
theta = T_(O.init_params(ms_, 1, stable=True)); thc = T_(O.init_params(cs_, 2))
u = torch.rand(B, T, 1, device=dev) * 2 - 1
ybar = torch.randn(B, T + 1, 1, device=dev) / (B * T)
if which == 'open_fwd':
    ops.rollout_fwd(spec, theta, u)
elif which == 'open_bwd':
    y, _, tape = ops.rollout_fwd(spec, theta, u, want_tape=True)
    ops.rollout_bwd(spec, theta, u, tape, ybar)
elif which == 'closed_fwd':
    ops.closedloop_fwd(cs, spec, thc, theta, u)
elif which == 'closed_bwd':
    y, _, tape = ops.closedloop_fwd(cs, spec, thc, theta, u, want_tape=True)
    ops.closedloop_bwd(cs, spec, thc, theta, u, tape, ybar[:, :T])
torch.cuda.synchronize()
print('done', which)
