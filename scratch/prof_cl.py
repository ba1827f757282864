"""scratch: one closed-loop forward + reverse pass at the bench size (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as BN
from chain_control_b200 import ops
dev = torch.device("cuda:0")
ctrl, model = BN.specs()
thc = torch.tensor(BN.init_theta(ctrl, 2, False), device=dev)
thm = torch.tensor(BN.init_theta(model, 1, True), device=dev)
B, T = int(os.environ.get("PB", 65536)), int(os.environ.get("PT", 1001))
refs = torch.tensor(BN.make_refs(B, T, 7), device=dev)
for it in range(2):
    yh, _, tape = ops.closedloop_fwd(ctrl, model, thc, thm, refs, want_tape=True)
    ybar = (2.0 / yh.numel()) * (yh - refs)
    g = ops.closedloop_bwd(ctrl, model, thc, thm, refs, tape, ybar)
torch.cuda.synchronize()
print("ok", float(g.norm()))
