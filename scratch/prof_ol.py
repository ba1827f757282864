"""scratch: one open-loop forward + reverse pass (ModelTrainer, cfg4 size) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as BN
from chain_control_b200 import ops
dev = torch.device("cuda:0")
ctrl, model = BN.specs()
thm = torch.tensor(BN.init_theta(model, 1, True), device=dev)
B, T = int(os.environ.get("PB", 65536)), int(os.environ.get("PT", 1000))
u = torch.rand(B, T, 1, device=dev) * 2 - 1
for it in range(2):
    y, _, tape = ops.rollout_fwd(model, thm, u, want_tape=True)
    ybar = (2.0 / y.numel()) * y
    g, _ = ops.rollout_bwd(model, thm, u, tape, ybar)
torch.cuda.synchronize()
print("ok", float(g.norm()))
