"""timing of the deep-MLP reverse pass (family 6) and of the FFMA reverse pass it replaces"""
import os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from cchelpers import conv
from oracle import np_oracle as O
from chain_control_b200 import ops
def run(S, W, depth, B, T, env=None, reps=3):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k); os.environ[k] = v
    try:
        spec = O.make_node(S, 1, 1, f_width=W, f_depth=depth)
        theta = torch.tensor(O.init_params(spec, 1).astype(np.float32), device='cuda')
        g = torch.Generator(device='cuda').manual_seed(0)
        u = torch.rand((B, T, 1), device='cuda', generator=g) * 2 - 1
        ybar = torch.randn((B, T + 1, 1), device='cuda', generator=g)
        cs = conv(spec)
        y, _, tape = ops.rollout_fwd(cs, theta, u, want_tape=True)
        tb, _ = ops.rollout_bwd(cs, theta, u, tape, ybar)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = tbk = 0.0
        for _ in range(reps):
            e[0].record(); y, _, tape = ops.rollout_fwd(cs, theta, u, want_tape=True)
            e[1].record(); tb, _ = ops.rollout_bwd(cs, theta, u, tape, ybar)
            e[2].record(); torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1]) / reps; tbk += e[1].elapsed_time(e[2]) / reps
        n = B * T
        print(f'({S},{W},{depth}) B={B} T={T} env={env}: fwd {tf:.2f} ms ({n / tf / 1e3:.3e} traj-steps/s)  bwd {tbk:.2f} ms ({n / tbk / 1e3:.3e})  fwd+bwd {n / (tf + tbk) / 1e3:.3e} traj-steps/s, |tb| {float(tb.norm()):.4e}', flush=True)
    finally:
        for k, v in old.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
if __name__ == '__main__':
    T = int(os.environ.get('TT', '64'))
    run(64, 64, 2, 18944, T)
    run(64, 64, 2, 65536, T)
    run(64, 64, 2, 65536, T, {'CC_B200_MLPB_FLUSH': '8'})
    run(64, 64, 1, 65536, T)
    run(32, 32, 2, 65536, T, {'CC_B200_MLPB': '0'}, reps=1)
    run(32, 32, 2, 65536, T, {'CC_B200_MLPB': '1'})
    run(5, 10, 2, 65536, T, {'CC_B200_MLPB': '0'}, reps=1)
    run(5, 10, 2, 65536, T, {'CC_B200_MLPB': '1'})
    run(40, 48, 2, 65536, T, {'CC_B200_MLPB': '0'}, reps=1)
    run(40, 48, 2, 65536, T)
