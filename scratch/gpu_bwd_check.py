"""GPU: backward parity vs the C oracle (fp64) + raw timing (scratch)."""
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import np_oracle as O, c_oracle as CO
from ccspecs import model_variants, controller_variants
from chain_control_b200 import arch as A, ops, _lib
def conv(s):
    return A.NodeSpec(s.state_dim, s.input_dim, s.output_dim, A.MlpSpec(s.f.sizes,s.f.act,s.f.act_final,s.f.use_bias), A.MlpSpec(s.g.sizes,s.g.act,s.g.act_final,s.g.use_bias), s.dt, s.integrator, s.f_time_variant, s.g_time_variant, s.g_feedthrough_u, s.readout_pre_step, s.u_transform, s.u_scale, s.out_scale)
dev = 'cuda'
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device=dev)
def nerr(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'
if not quick:
  for name, spec in model_variants().items():
    B, T = 300, 200
    theta = O.init_params(spec, 1).astype(np.float32); u = np.random.default_rng(0).uniform(-1, 1, (B, T, spec.input_dim)).astype(np.float32)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    ref = CO.rollout(spec, theta, u, ybar=ybar, want_ubar=True, dtype=np.float64)
    y, _, tape = ops.rollout_fwd(conv(spec), T_(theta), T_(u), want_tape=True)
    tb, ub = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar), want_ubar=True)
    print('open', name, 'theta_bar L2 err %.2e  ubar err %.2e' % (nerr(tb.cpu().numpy(), ref['theta_bar']), nerr(ub.cpu().numpy(), ref['ubar'])))
  for mname in ['compact_s50_d0', 'full_s5_w10_d2']:
    mspec = model_variants()[mname]
    for cname, cspec in controller_variants().items():
      B, Tr = 300, 200
      thc = O.init_params(cspec, 2).astype(np.float32); thm = O.init_params(mspec, 1).astype(np.float32)
      refs = (3 * np.random.default_rng(0).uniform(-1, 1, (B, Tr, 1))).astype(np.float32)
      ybar = np.random.default_rng(6).normal(size=(B, Tr, 1)).astype(np.float32)
      ref = CO.closedloop(cspec, mspec, thc, thm, refs, ybar=ybar, dtype=np.float64)
      yh, _, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), want_tape=True)
      tb = ops.closedloop_bwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), tape, T_(ybar))
      print('closed', mname, cname, 'theta_c_bar L2 err %.2e' % nerr(tb.cpu().numpy(), ref['theta_c_bar']))
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
ms_ = model_variants()['compact_s50_d0']; cs_ = controller_variants()['compact_s5_d0']
spec = conv(ms_); cs = conv(cs_)
theta = T_(O.init_params(ms_, 1, stable=True)); thc = T_(O.init_params(cs_, 2))
bts = [None] + [int(x) for x in os.environ.get('BTS', '').split(',') if x]
for B in [64, 65536]:
    T = 1000
    u = torch.rand(B, T, 1, device=dev) * 2 - 1
    ybar = torch.randn(B, T + 1, 1, device=dev) / (B * T)
    refs = torch.rand(B, T + 1, 1, device=dev) * 2 - 1
    ybc = torch.randn(B, T + 1, 1, device=dev) / (B * T)
    for bt in (bts if B > 64 else [None]):
        if bt: os.environ['CC_B200_BT'] = str(bt)
        else: os.environ.pop('CC_B200_BT', None)
        try:
            f0 = timeit(lambda: ops.rollout_fwd(spec, theta, u))
            y, _, tape = ops.rollout_fwd(spec, theta, u, want_tape=True)
            f1 = timeit(lambda: ops.rollout_fwd(spec, theta, u, want_tape=True))
            b1 = timeit(lambda: ops.rollout_bwd(spec, theta, u, tape, ybar))
            print('open  B=%d BT=%s fwd %.2f ms, fwd+tape %.2f ms, bwd %.2f ms -> fwd %.3e  fwd+bwd %.3e steps/s (%.1f / %.1f TFLOP/s alg)' % (B, bt, f0, f1, b1, B*T/f0*1e3, B*T/(f1+b1)*1e3, B*T*20500/f0/1e9, B*T*61500/(f1+b1)/1e9))
            del tape
            c0 = timeit(lambda: ops.closedloop_fwd(cs, spec, thc, theta, refs))
            yh, _, tape = ops.closedloop_fwd(cs, spec, thc, theta, refs, want_tape=True)
            c1 = timeit(lambda: ops.closedloop_fwd(cs, spec, thc, theta, refs, want_tape=True))
            cb = timeit(lambda: ops.closedloop_bwd(cs, spec, thc, theta, refs, tape, ybc))
            Tr = T + 1
            print('closed B=%d BT=%s fwd %.2f ms, fwd+tape %.2f ms, bwd %.2f ms -> fwd %.3e  fwd+bwd %.3e steps/s (%.1f / %.1f TFLOP/s alg)' % (B, bt, c0, c1, cb, B*Tr/c0*1e3, B*Tr/(c1+cb)*1e3, B*Tr*20790/c0/1e9, B*Tr*41870/(c1+cb)/1e9))
            del tape
        except Exception as e:
            print('BT', bt, 'failed', e)
