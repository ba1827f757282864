"""scratch: one FFMA closed-loop forward + reverse pass of the CLI architecture (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from chain_control_b200 import ops
from oracle import np_oracle as O
from cchelpers import conv
dev = "cuda"
cspec = O.make_node(5, 2, 1, f_width=10, f_depth=2)
mspec = O.make_node(3, 1, 1, f_depth=0, f_use_bias=False, g_use_bias=False, integrator=O.INT_EE, readout_pre_step=True)
thc = torch.tensor(O.init_params(cspec, 2).astype(np.float32), device=dev)
thm = torch.tensor(O.init_params(mspec, 1, stable=True).astype(np.float32), device=dev)
B, T = int(os.environ.get("PB", 960)), int(os.environ.get("PT", 100))
refs = torch.rand(B, T, 1, device=dev) * 2 - 1
for it in range(2):
    yh, _, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), thc, thm, refs, want_tape=True)
    ybar = (yh - refs) / yh.numel()
    g = ops.closedloop_bwd(conv(cspec), conv(mspec), thc, thm, refs, tape, ybar)
torch.cuda.synchronize()
print("ok", float(g.norm()))
