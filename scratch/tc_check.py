"""GPU: tcgen05 path vs the fp64 C oracle and vs the FFMA path; raw timing (scratch)."""
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import np_oracle as O, c_oracle as CO
from ccspecs import model_variants, controller_variants
from chain_control_b200 import arch as A, ops
def conv(s):
    return A.NodeSpec(s.state_dim, s.input_dim, s.output_dim, A.MlpSpec(s.f.sizes,s.f.act,s.f.act_final,s.f.use_bias), A.MlpSpec(s.g.sizes,s.g.act,s.g.act_final,s.g.use_bias), s.dt, s.integrator, s.f_time_variant, s.g_time_variant, s.g_feedthrough_u, s.readout_pre_step, s.u_transform, s.u_scale, s.out_scale)
dev = 'cuda'
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device=dev)
def nerr(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
def merr(a, b): return float((np.abs(a - b).max(axis=(1, 2)) / np.abs(b).max(axis=(1, 2))).max())
def both(fn):
    out = {}
    for tc in ('1', '0'):
        os.environ['CC_B200_TC'] = tc
        out[tc] = fn()
    os.environ['CC_B200_TC'] = '1'
    return out['1'], out['0']
mode = sys.argv[1] if len(sys.argv) > 1 else 'all'
mv, cv = model_variants(), controller_variants()
if mode in ('all', 'parity'):
    for mname in ['compact_s50_d0']:
        ms = mv[mname]
        for B, T in [(300, 200), (128, 1000)]:
            theta = O.init_params(ms, 1, stable=True).astype(np.float32)
            u = np.random.default_rng(0).uniform(-1, 1, (B, T, ms.input_dim)).astype(np.float32)
            ref = CO.rollout(ms, theta, u, dtype=np.float64)
            ref32 = CO.rollout(ms, theta, u, dtype=np.float32)
            (y1, _, _), (y0, _, _) = both(lambda: ops.rollout_fwd(conv(ms), T_(theta), T_(u)))
            print('open fwd %s B=%d T=%d: tc err %.2e  ffma err %.2e  fp32-oracle err %.2e' % (mname, B, T, merr(y1.cpu().numpy(), ref['y']), merr(y0.cpu().numpy(), ref['y']), merr(ref32['y'], ref['y'])), flush=True)
        for cname in ['compact_s5_d0']:
            cs = cv[cname]
            for B, Tr in [(300, 200), (128, 1001)]:
                thc = O.init_params(cs, 2).astype(np.float32); thm = O.init_params(ms, 1, stable=True).astype(np.float32)
                refs = (3 * np.random.default_rng(0).uniform(-1, 1, (B, 1, 1)) * np.ones((B, Tr, 1))).astype(np.float32)
                ybar = np.random.default_rng(6).normal(size=(B, Tr, 1)).astype(np.float32)
                ref = CO.closedloop(cs, ms, thc, thm, refs, ybar=ybar, dtype=np.float64)
                def run():
                    yh, _, tape = ops.closedloop_fwd(conv(cs), conv(ms), T_(thc), T_(thm), T_(refs), want_tape=True)
                    tb = ops.closedloop_bwd(conv(cs), conv(ms), T_(thc), T_(thm), T_(refs), tape, T_(ybar))
                    return yh.cpu().numpy(), tb.cpu().numpy()
                (yh1, tb1), (yh0, tb0) = both(run)
                print('closed %s+%s B=%d Tr=%d: yhat tc %.2e ffma %.2e | theta_c_bar tc %.2e ffma %.2e' % (mname, cname, B, Tr, merr(yh1, ref['yhat']), merr(yh0, ref['yhat']), nerr(tb1, ref['theta_c_bar']), nerr(tb0, ref['theta_c_bar'])), flush=True)
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
if mode in ('all', 'time'):
    ms_ = mv['compact_s50_d0']; cs_ = cv['compact_s5_d0']
    spec = conv(ms_); cs = conv(cs_)
    theta = T_(O.init_params(ms_, 1, stable=True)); thc = T_(O.init_params(cs_, 2))
    for B in [int(x) for x in os.environ.get('BS', '65536').split(',')]:
        T = 1000
        u = torch.rand(B, T, 1, device=dev) * 2 - 1
        refs = torch.rand(B, T + 1, 1, device=dev) * 2 - 1
        ybc = torch.randn(B, T + 1, 1, device=dev) / (B * T)
        for tc in os.environ.get('TCS', '1,0').split(','):
            os.environ['CC_B200_TC'] = tc
            f0 = timeit(lambda: ops.rollout_fwd(spec, theta, u))
            c0 = timeit(lambda: ops.closedloop_fwd(cs, spec, thc, theta, refs))
            yh, _, tape = ops.closedloop_fwd(cs, spec, thc, theta, refs, want_tape=True)
            c1 = timeit(lambda: ops.closedloop_fwd(cs, spec, thc, theta, refs, want_tape=True))
            cb = timeit(lambda: ops.closedloop_bwd(cs, spec, thc, theta, refs, tape, ybc))
            Tr = T + 1
            print('TC=%s B=%d open fwd %.2f ms (%.3e steps/s, %.1f TFLOP/s alg) | closed fwd %.2f ms, fwd+tape %.2f ms, bwd %.2f ms -> fwd+bwd %.3e steps/s (%.1f TFLOP/s alg)' % (tc, B, f0, B*T/f0*1e3, B*T*20500/f0/1e9, c0, c1, cb, B*Tr/(c1+cb)*1e3, B*Tr*41870/(c1+cb)/1e9), flush=True)
            del tape
