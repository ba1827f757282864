"""BASELINE.json configs 3 (forward-only sweep) and 5 (T = 10,000 closed-loop BPTT) on one B200 -> JSON lines."""
import sys, os, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from chain_control_b200 import arch as A, ops, _lib
dev = 'cuda'
def init(spec, seed, stable=True):
    rng = np.random.default_rng(seed); out = []
    for mlp in (spec.f, spec.g):
        L = mlp.n_layers
        for li, (i, o) in enumerate(zip(mlp.sizes[:-1], mlp.sizes[1:])):
            W = rng.uniform(-1, 1, size=(o, i)) / np.sqrt(i)
            if stable and mlp is spec.f and L == 1: W[:, :spec.state_dim] -= np.eye(spec.state_dim)
            if stable and L > 1: W *= 0.5
            out += [W.ravel(), rng.uniform(-1, 1, size=o) / np.sqrt(i)]
    return torch.tensor(np.concatenate(out).astype(np.float32), device=dev)
def timeit(fn, n=2):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
def flops(spec):
    mf = sum(i * o for i, o in zip(spec.f.sizes[:-1], spec.f.sizes[1:])); mg = sum(i * o for i, o in zip(spec.g.sizes[:-1], spec.g.sizes[1:]))
    return 2 * (4 * mf + mg)
res = []
# cfg3: forward-only, T = 1000, u ~ U(-1, 1)
T = 1000
archs = {'S50_d0': A.make_node_spec(50, 1, 1, 0.01, f_depth=0, readout_pre_step=True, u_transform=A.UT_ARCTAN),
         'S32_W32_d2': A.make_node_spec(32, 1, 1, 0.01, f_width=32, f_depth=2, f_act_final=A.ACT_TANH),
         'S64_W64_d2': A.make_node_spec(64, 1, 1, 0.01, f_width=64, f_depth=2, f_act_final=A.ACT_TANH),
         'S128_W128_d2': A.make_node_spec(128, 1, 1, 0.01, f_width=128, f_depth=2, f_act_final=A.ACT_TANH)}
for name, spec in archs.items():
    th = init(spec, 1)
    for lb in (10, 12, 14, 16, 18, 20):
        B = 2 ** lb
        if B * T * 8 > 40e9: continue
        try:
            u = torch.rand(B, T, 1, device=dev) * 2 - 1
            ms = timeit(lambda: ops.rollout_fwd(spec, th, u))
            r = dict(config='cfg3', arch=name, B=B, T=T, ms=ms, traj_steps_per_s=B * T / ms * 1e3, alg_tflops=B * T * flops(spec) / ms / 1e9)
        except Exception as e:
            r = dict(config='cfg3', arch=name, B=B, T=T, error=str(e)[:160])
        res.append(r); print(json.dumps(r), flush=True)
        if 'error' in r: break
        del u
# cfg5: closed loop, T = 10,000, B in {64, 4096}, stable parameters, tape every step (ckpt_every = 1)
model = archs['S50_d0']; ctrl = A.make_node_spec(5, 2, 1, 0.01, f_depth=0, readout_pre_step=True, out_scale=0.5)
thm, thc = init(model, 1), init(ctrl, 2, stable=False)
for B in (64, 4096):
    Tr = 10001
    refs = (torch.sign(torch.randn(B, 1, 1, device=dev)) * torch.rand(B, 1, 1, device=dev) * 3).expand(B, Tr, 1).contiguous()
    def step():
        yh, _, tape = ops.closedloop_fwd(ctrl, model, thc, thm, refs, want_tape=True)
        loss, yb = ops.mse_value_and_cotangent(refs, yh)
        return ops.closedloop_bwd(ctrl, model, thc, thm, refs, tape, yb), loss
    ms = timeit(step)
    g, loss = step()
    r = dict(config='cfg5', B=B, T=Tr, ms=ms, traj_steps_per_s=B * Tr / ms * 1e3, alg_tflops=B * Tr * 41870 / ms / 1e9, loss=float(loss), grad_finite=bool(torch.isfinite(g).all()), tape_MB=6 * 4 * max(B, 32) * Tr / 1e6)
    res.append(r); print(json.dumps(r), flush=True)
json.dump(res, open('gpurun_out/sweep_r01.json', 'w'), indent=1)
