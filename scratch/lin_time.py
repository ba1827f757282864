"""scratch: kernel times of the blocked linear family vs the older families (closed loop cfg2x / cfg2, open loop cfg4 / cfg1)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench as BN
from chain_control_b200 import ops

dev = torch.device("cuda:0")
ctrl, model = BN.specs()
thc = torch.tensor(BN.init_theta(ctrl, 2, False), device=dev)
thm = torch.tensor(BN.init_theta(model, 1, True), device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def closed(B, T, env):
    for k, v in env.items(): os.environ[k] = v
    refs = torch.tensor(BN.make_refs(B, T, 7), device=dev)
    out = {}
    st = {}
    def f():
        st["yh"], _, st["tape"] = ops.closedloop_fwd(ctrl, model, thc, thm, refs, want_tape=True)
    out["fwd_ms"] = timeit(f)
    ybar = (2.0 / st["yh"].numel()) * (st["yh"] - refs)
    def b():
        st["g"] = ops.closedloop_bwd(ctrl, model, thc, thm, refs, st["tape"], ybar)
    out["bwd_ms"] = timeit(b)
    out["g_norm"] = float(st["g"].norm())
    out["yh_abs"] = float(st["yh"].abs().mean())
    for k in env: os.environ.pop(k, None)
    return out

def open_(B, T, env):
    for k, v in env.items(): os.environ[k] = v
    u = torch.rand(B, T, 1, device=dev) * 2 - 1
    out = {}; st = {}
    def f():
        st["y"], _, st["tape"] = ops.rollout_fwd(model, thm, u, want_tape=True)
    out["fwd_ms"] = timeit(f)
    ybar = (2.0 / st["y"].numel()) * st["y"]
    def b():
        st["g"], _ = ops.rollout_bwd(model, thm, u, st["tape"], ybar)
    out["bwd_ms"] = timeit(b)
    out["g_norm"] = float(st["g"].norm())
    for k in env: os.environ.pop(k, None)
    return out

res = {}
for name, B, T in [("cfg2x", 65536, 1001), ("cfg2", 64, 1001), ("b8192", 8192, 1001)]:
    for tag, env in [("lin", {}), ("lin_nostagebwd", {"CC_B200_LIN_STAGE_BWD": "0"}), ("lin_nostage", {"CC_B200_LIN_STAGE": "0"}), ("lin_nolc", {"CC_B200_LIN_LC": "0"})]:
        r = closed(B, T, env)
        res[f"closed_{name}_{tag}"] = r
        print(f"closed {name} {tag}: {r}", flush=True)
json.dump(res, open("gpurun_out/lin_time2.json", "w"), indent=1)
