"""scratch: one forward + reverse rollout on the deep-MLP tcgen05 families (for ncu): PS = state/width, PB, PT."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chain_control_b200 import arch as A, ops
dev = torch.device("cuda:0")
S = int(os.environ.get("PS", 64)); B = int(os.environ.get("PB", 18944)); T = int(os.environ.get("PT", 16))
spec = A.make_node_spec(S, 1, 1, 0.01, f_width=S, f_depth=2)
rng = np.random.default_rng(S); parts = []
for mlp in (spec.f, spec.g):
    for i, o in zip(mlp.sizes[:-1], mlp.sizes[1:]):
        parts += [0.5 * rng.uniform(-1, 1, size=o * i) / np.sqrt(i), rng.uniform(-1, 1, size=o) / np.sqrt(i)]
th = torch.tensor(np.concatenate(parts).astype(np.float32), device=dev)
u = torch.rand(B, T, 1, device=dev) * 2 - 1
ybar = torch.randn(B, T + 1, 1, device=dev)
for it in range(int(os.environ.get("PR", 1))):
    y, _, tape = ops.rollout_fwd(spec, th, u, want_tape=True)
    tb, _ = ops.rollout_bwd(spec, th, u, tape, ybar)
torch.cuda.synchronize()
print("ok", float(tb.norm()))
