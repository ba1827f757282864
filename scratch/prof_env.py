"""scratch: one rollout of the two-segment data source (for ncu / timing): PB episodes x PT control steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chain_control_b200.env import make_env
B = int(os.environ.get("PB", 16384)); T = int(os.environ.get("PT", 1000))
env = make_env("two_segments_v1", n_envs=B)
acts = torch.rand(B, T, device="cuda") * 2 - 1
for _ in range(int(os.environ.get("PR", 1))):
    obs = env.rollout(acts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if os.environ.get("PTIME"):
    for B2 in (1024, 16384, 148 * 128 * 4, 262144):
        env2 = make_env("two_segments_v1", n_envs=B2); a2 = torch.rand(B2, T, device="cuda") * 2 - 1
        env2.rollout(a2); e0.record(); env2.rollout(a2); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"B={B2} T={T}: {ms:.2f} ms  {B2 * T / ms * 1e3:.3e} env-steps/s  {10 * B2 * T / ms * 1e3:.3e} physics-steps/s")
print("ok", float(obs["xpos_of_segment_end"].abs().max()))
