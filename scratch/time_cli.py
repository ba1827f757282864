"""scratch: the CLI workflow's kernel calls (deep controller 5/10/10 + LTI model S=3, 30 x 1501) timed separately."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from chain_control_b200 import arch as A, ops
from oracle import np_oracle as O
from cchelpers import conv
dev = "cuda"
cspec = O.make_node(5, 2, 1, f_width=10, f_depth=2)
mspec = O.make_node(3, 1, 1, f_depth=0, f_use_bias=False, g_use_bias=False, integrator=O.INT_EE, readout_pre_step=True)
thc = torch.tensor(O.init_params(cspec, 2).astype(np.float32), device=dev)
thm = torch.tensor(O.init_params(mspec, 1, stable=True).astype(np.float32), device=dev)
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return 1e3 * (time.time() - t0) / n
os.environ["CC_B200_VERBOSE"] = "1"
for B in (30, 960, 30720):
    refs = torch.rand(B, 1501, 1, device=dev) * 2 - 1
    yh, _, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), thc, thm, refs, want_tape=True)
    os.environ.pop("CC_B200_VERBOSE", None)
    ybar = (yh - refs) / yh.numel()
    f = t(lambda: ops.closedloop_fwd(conv(cspec), conv(mspec), thc, thm, refs, want_tape=True))
    b = t(lambda: ops.closedloop_bwd(conv(cspec), conv(mspec), thc, thm, refs, tape, ybar))
    print(f"B={B}: fwd {f:.2f} ms ({1e3*f/1501:.2f} us/step), bwd {b:.2f} ms ({1e3*b/1501:.2f} us/step)")
