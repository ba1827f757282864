"""scratch: observed parity errors of the deep-MLP tcgen05 family (forward) and of the CUDA path against the reference-executed golden
fixtures -> gpurun_out/r02b_parity_floors.json (summarised in profiles/r02b_parity_floors.md)."""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import np_oracle as O
from oracle import c_oracle as CO
from ccspecs import mlp_model_variants
from cchelpers import conv, normwise, l2rel
from chain_control_b200 import ops
import test_golden as TG
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device="cuda")
rows = []
for name, spec in mlp_model_variants().items():
    for B, T in ((2500, 60), (512, 1000)):
        theta = O.init_params(spec, 1).astype(np.float32)
        if T == 1000:  # long horizon: contractive scaling of f's weights keeps the random-init dynamics bounded
            theta[: spec.f.n_params()] *= 0.5
        u = np.random.default_rng(0).uniform(-1, 1, (B, T, spec.input_dim)).astype(np.float32)
        x0 = (0.1 * np.random.default_rng(3).normal(size=(B, spec.state_dim))).astype(np.float32)
        r64 = CO.rollout(spec, theta, u, x0=x0, dtype=np.float64); r32 = CO.rollout(spec, theta, u, x0=x0, dtype=np.float32)
        y, _, _ = ops.rollout_fwd(conv(spec), T_(theta), T_(u), T_(x0))
        rows.append(dict(case="family 5: " + name, B=B, T=T, y_kernel=normwise(y.cpu().numpy(), r64["y"]), y_fp32_oracle=normwise(r32["y"], r64["y"]), g_kernel=None, g_fp32_oracle=None))
        print(rows[-1], flush=True)
for name in TG.OPEN:
    g = TG.load(name)
    spec = conv(TG.spec_of(str(g["spec"])))
    theta, u, tgt = g["theta"], g["u"], g["targets"]
    x0 = TG._x0_of(g, u.shape[0]); x0 = None if x0 is None else T_(x0)
    y, _, tape = ops.rollout_fwd(spec, T_(theta), T_(u), x0, want_tape=True)
    ge = None
    try:
        tb, _ = ops.rollout_bwd(spec, T_(theta), T_(u), tape, (2.0 / y.numel()) * (y - T_(tgt)), x0)
        ge = l2rel(tb.cpu().numpy().astype(np.float64) + TG.regu_grad(theta, float(g["l1"]), float(g["l2"])), g["theta_bar"])
    except Exception:
        pass
    rows.append(dict(case="golden (reference-executed): " + name, B=u.shape[0], T=u.shape[1], y_kernel=normwise(y.cpu().numpy(), g["y"]), y_fp32_oracle=None, g_kernel=ge, g_fp32_oracle=None))
    print(rows[-1], flush=True)
for name in TG.CLOSED:
    g = TG.load(name)
    cspec = conv(TG.spec_of(str(g["cspec"])))
    mspecs = {k: conv(O.make_node(**v)) for k, v in json.loads(str(g["mspecs"])).items()}
    names = json.loads(str(g["model_names"]))
    refs, (thc, chain) = g["refs"], TG.ctrl_theta_and_chain(g)
    tb = np.zeros_like(thc); ye = 0.0
    for n in names:
        thm = g[f"theta_m__{n}"]
        noise = T_(g[f"noise__{n}"]) if f"noise__{n}" in g.files else None
        yhat, _, tape = ops.closedloop_fwd(cspec, mspecs[n], T_(thc), T_(thm), T_(refs), noise=noise, want_tape=True)
        ye = max(ye, normwise(yhat.cpu().numpy(), g[f"yhat__{n}"]))
        ybar = (2.0 / (yhat.numel() * len(names))) * (yhat - T_(refs))
        tb += ops.closedloop_bwd(cspec, mspecs[n], T_(thc), T_(thm), T_(refs), tape, ybar, noise=noise).cpu().numpy()
    rows.append(dict(case="golden (reference-executed): " + name, B=refs.shape[0], T=refs.shape[1], y_kernel=ye, y_fp32_oracle=None, g_kernel=l2rel(chain(tb), g["theta_c_bar"]), g_fp32_oracle=None))
    print(rows[-1], flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02b_parity_floors.json"), "w"), indent=1)
