"""BASELINE.json configs[2] (forward-only sweep B = 1K..1M, MLP width 32..128) on one B200 -> JSON lines.
Each (architecture, B) is timed on the default plan and, where both exist, on the FFMA family (CC_B200_MLP=0)."""
import sys, os, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import ctypes as C
import numpy as np, torch
from chain_control_b200 import arch as A, ops, _lib
dev = 'cuda'
out_path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/r02_sweep_cfg3.jsonl'
def init(spec, seed):
    rng = np.random.default_rng(seed); out = []
    for mlp in (spec.f, spec.g):
        for i, o in zip(mlp.sizes[:-1], mlp.sizes[1:]):
            W = rng.uniform(-1, 1, size=(o, i)) / np.sqrt(i)
            if mlp.n_layers > 1: W *= 0.5
            out += [W.ravel(), rng.uniform(-1, 1, size=o) / np.sqrt(i)]
    return torch.tensor(np.concatenate(out).astype(np.float32), device=dev)
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
def flops(spec):
    mf = sum(i * o for i, o in zip(spec.f.sizes[:-1], spec.f.sizes[1:])); mg = sum(i * o for i, o in zip(spec.g.sizes[:-1], spec.g.sizes[1:]))
    return 2 * (4 * mf + mg)
T = 1000
archs = {'S5_W10_d2 (reference default)': A.make_node_spec(5, 1, 1, 0.01),
         'S32_W32_d2': A.make_node_spec(32, 1, 1, 0.01, f_width=32, f_depth=2, f_act_final=A.ACT_TANH),
         'S64_W64_d2': A.make_node_spec(64, 1, 1, 0.01, f_width=64, f_depth=2, f_act_final=A.ACT_TANH),
         'S128_W128_d2': A.make_node_spec(128, 1, 1, 0.01, f_width=128, f_depth=2, f_act_final=A.ACT_TANH)}
only = os.environ.get('SWEEP_ARCHS')
f = open(out_path, 'w')
for name, spec in archs.items():
    if only and not any(k in name for k in only.split(',')): continue
    th = init(spec, 1)
    for lb in (10, 12, 14, 16, 17, 18, 20):
        B = 2 ** lb
        Tb = T if lb < 20 else 250
        u = torch.rand(B, Tb, 1, device=dev) * 2 - 1
        for fam_env in (None, '0'):
            if fam_env is None: os.environ.pop('CC_B200_MLP', None)
            else: os.environ['CC_B200_MLP'] = fam_env
            try:
                fam = _lib.lib().cc_kernel_family_at(None, C.byref(spec.to_c()), 0, B)
                if fam_env == '0' and lb > 17: continue
                ms = timeit(lambda: ops.rollout_fwd(spec, th, u))
                r = dict(config='cfg3', arch=name, B=B, T=Tb, family=fam, ms=ms, traj_steps_per_s=B * Tb / ms * 1e3, alg_tflops=B * Tb * flops(spec) / ms / 1e9)
            except Exception as e:
                r = dict(config='cfg3', arch=name, B=B, T=Tb, family_env=fam_env, error=str(e)[:160])
            print(json.dumps(r), flush=True); f.write(json.dumps(r) + '\n'); f.flush()
        del u
os.environ.pop('CC_B200_MLP', None)
