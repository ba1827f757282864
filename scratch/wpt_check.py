"""GPU: latency kernels (warp per trajectory) vs the fp64 oracle and the other families; small-batch timing (scratch)."""
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import np_oracle as O, c_oracle as CO
from ccspecs import model_variants, controller_variants, tc_model_variants, tc_controller_variants
from cchelpers import conv
from chain_control_b200 import ops
dev = 'cuda'
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device=dev)
def nerr(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
def merr(a, b): return float((np.abs(a - b).max(axis=(1, 2)) / np.abs(b).max(axis=(1, 2))).max())
FAM = {'latency': ('1', str(1 << 30)), 'tcgen05': ('1', '0'), 'ffma': ('0', '0')}
def setfam(f):
    os.environ['CC_B200_TC'], os.environ['CC_B200_WPT_MAXB'] = FAM[f]
for mname, ms in tc_model_variants().items():
    for cname, cs in tc_controller_variants().items():
        B, Tr = 70, 150
        thc = O.init_params(cs, 2).astype(np.float32); thm = O.init_params(ms, 1).astype(np.float32)
        refs = (3 * np.random.default_rng(0).uniform(-1, 1, (B, Tr, 1))).astype(np.float32)
        ybar = np.random.default_rng(6).normal(size=(B, Tr, 1)).astype(np.float32)
        ref = CO.closedloop(cs, ms, thc, thm, refs, ybar=ybar, dtype=np.float64)
        setfam('latency')
        yh, us, tape = ops.closedloop_fwd(conv(cs), conv(ms), T_(thc), T_(thm), T_(refs), want_tape=True, want_u=True)
        tb = ops.closedloop_bwd(conv(cs), conv(ms), T_(thc), T_(thm), T_(refs), tape, T_(ybar))
        print('closed %-20s %-14s yhat %.2e u %.2e grad %.2e' % (mname, cname, merr(yh.cpu().numpy(), ref['yhat']), merr(us.cpu().numpy(), ref['u']), nerr(tb.cpu().numpy(), ref['theta_c_bar'])), flush=True)
    u = np.random.default_rng(0).uniform(-1, 1, (33, 120, 1)).astype(np.float32)
    ref = CO.rollout(ms, thm, u, dtype=np.float64)
    y, _, _ = ops.rollout_fwd(conv(ms), T_(thm), T_(u))
    print('open   %-20s y %.2e' % (mname, merr(y.cpu().numpy(), ref['y'])), flush=True)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
ms_ = model_variants()['compact_s50_d0']; cs_ = controller_variants()['compact_s5_d0']
spec = conv(ms_); cs = conv(cs_)
theta = T_(O.init_params(ms_, 1, stable=True)); thc = T_(O.init_params(cs_, 2))
for B in (8, 64, 256, 1024, 2048, 4096):
    Tr = 1001
    refs = torch.rand(B, Tr, 1, device=dev) * 2 - 1
    ybc = torch.randn(B, Tr, 1, device=dev) / (B * Tr)
    for fam in ('latency', 'tcgen05'):
        setfam(fam)
        c0 = timeit(lambda: ops.closedloop_fwd(cs, spec, thc, theta, refs, want_tape=True))
        yh, _, tape = ops.closedloop_fwd(cs, spec, thc, theta, refs, want_tape=True)
        cb = timeit(lambda: ops.closedloop_bwd(cs, spec, thc, theta, refs, tape, ybc))
        print('%-8s B=%5d closed fwd+tape %.2f ms, bwd %.2f ms -> %.3e traj-steps/s' % (fam, B, c0, cb, B * Tr / (c0 + cb) * 1e3), flush=True)
