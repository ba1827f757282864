"""first GPU contact: forward parity vs the C oracle + raw timing (scratch, not a test)."""
import sys, time, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import np_oracle as O, c_oracle as CO
from ccspecs import model_variants, controller_variants
from chain_control_b200 import arch as A, ops, _lib
def conv(s):
    return A.NodeSpec(s.state_dim, s.input_dim, s.output_dim, A.MlpSpec(s.f.sizes,s.f.act,s.f.act_final,s.f.use_bias), A.MlpSpec(s.g.sizes,s.g.act,s.g.act_final,s.g.use_bias), s.dt, s.integrator, s.f_time_variant, s.g_time_variant, s.g_feedthrough_u, s.readout_pre_step, s.u_transform, s.u_scale, s.out_scale)
dev = 'cuda'
print(torch.cuda.get_device_name(0))
# parity
for name, spec in model_variants().items():
    B, T = 300, 200
    theta = O.init_params(spec, 1).astype(np.float32); u = np.random.default_rng(0).uniform(-1, 1, (B, T, spec.input_dim)).astype(np.float32)
    ref = CO.rollout(spec, theta, u, dtype=np.float64)['y']
    y, _, _ = ops.rollout_fwd(conv(spec), torch.tensor(theta, device=dev), torch.tensor(u, device=dev))
    y = y.cpu().numpy()
    err = (np.abs(y - ref).max(axis=(1, 2)) / np.abs(ref).max(axis=(1, 2))).max()
    print('open', name, 'normwise err vs fp64: %.2e' % err)
mspec = model_variants()['compact_s50_d0']
for cname, cspec in controller_variants().items():
    B, Tr = 300, 200
    thc = O.init_params(cspec, 2).astype(np.float32); thm = O.init_params(mspec, 1).astype(np.float32)
    refs = (3 * np.random.default_rng(0).uniform(-1, 1, (B, Tr, 1))).astype(np.float32)
    ref = CO.closedloop(cspec, mspec, thc, thm, refs, dtype=np.float64)['yhat']
    y, _, _ = ops.closedloop_fwd(conv(cspec), conv(mspec), torch.tensor(thc, device=dev), torch.tensor(thm, device=dev), torch.tensor(refs, device=dev))
    err = (np.abs(y.cpu().numpy() - ref).max(axis=(1, 2)) / np.abs(ref).max(axis=(1, 2))).max()
    print('closed', cname, 'normwise err vs fp64: %.2e' % err)
# timing
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
L = _lib.lib()
sink = torch.zeros(4, device=dev)
iters = 4096
ms = timeit(lambda: L.check(L.cc_fp32_peak_probe(sink.data_ptr(), iters, torch.cuda.current_stream().cuda_stream), 'probe'))
flops = 148 * 4 * 512 * iters * 16 * 8 * 2
print('fp32 probe: %.3f ms -> %.1f TFLOP/s' % (ms, flops / ms / 1e9))
spec = conv(model_variants()['compact_s50_d0'])
theta = torch.tensor(O.init_params(model_variants()['compact_s50_d0'], 1, stable=True).astype(np.float32), device=dev)
import os
for B in [64, 4096, 65536]:
    T = 1000
    u = torch.rand(B, T, 1, device=dev) * 2 - 1
    for bt in ([None] if B < 65536 else [None, 64, 96, 112, 128, 160, 224]):
        if bt: os.environ['CC_B200_BT'] = str(bt)
        else: os.environ.pop('CC_B200_BT', None)
        try:
            ms = timeit(lambda: ops.rollout_fwd(spec, theta, u))
            print('open fwd S=50 B=%d T=%d BT=%s: %.2f ms -> %.3e traj-steps/s  (%.1f TFLOP/s alg)' % (B, T, bt, ms, B * T / ms * 1e3, B * T * 20500 / ms / 1e9))
        except Exception as e:
            print('BT', bt, 'failed', e)
os.environ.pop('CC_B200_BT', None)
cs = conv(controller_variants()['compact_s5_d0']); thc = torch.tensor(O.init_params(controller_variants()['compact_s5_d0'], 2).astype(np.float32), device=dev)
for B in [64, 65536]:
    refs = torch.rand(B, 1001, 1, device=dev) * 2 - 1
    ms = timeit(lambda: ops.closedloop_fwd(cs, spec, thc, theta, refs))
    print('closed fwd B=%d: %.2f ms -> %.3e traj-steps/s (%.1f TFLOP/s alg)' % (B, ms, B * 1001 / ms * 1e3, B * 1001 * 20790 / ms / 1e9))
