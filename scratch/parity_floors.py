"""scratch: observed parity errors of the kernels vs the fp64 oracle next to the fp32 restatement's own error (the noise floor
of each problem) -> gpurun_out/r02_parity_floors.json (summarised in profiles/r02_parity_floors.md)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import np_oracle as O
from oracle import c_oracle as CO
from ccspecs import model_variants, controller_variants, tc_model_variants
from cchelpers import conv, normwise, l2rel
from chain_control_b200 import ops
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device="cuda")
rows = []
def closed(tag, mspec, cspec, B, Tr, stable_c=False, noise_on=False, env=None, K=1):
    env = env or {}
    for k, v in env.items(): os.environ[k] = v
    thc = O.init_params(cspec, 2, stable=stable_c).astype(np.float32); thm = O.init_params(mspec, 1, stable=True).astype(np.float32)
    rng = np.random.default_rng(0)
    refs = (np.sign(rng.normal(size=(B, 1, 1))) * rng.uniform(0, 3, size=(B, 1, 1)) * np.ones((B, Tr, 1))).astype(np.float32)
    noise = (0.02 * rng.normal(size=(B, Tr, 1))).astype(np.float32) if noise_on else None
    r64 = CO.closedloop(cspec, mspec, thc, thm, refs, noise, mse_loss=True, dtype=np.float64)
    r32 = CO.closedloop(cspec, mspec, thc, thm, refs, noise, mse_loss=True, dtype=np.float32)
    yh, _, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), None if noise is None else T_(noise), want_tape=True, ckpt_every=K)
    ybar = O.mse_cotangent(refs, yh.cpu().numpy())
    g = ops.closedloop_bwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), tape, T_(ybar), None if noise is None else T_(noise), ckpt_every=K).cpu().numpy()
    rows.append(dict(case=tag, B=B, T=Tr, K=K, env=env, y_kernel=normwise(yh.cpu().numpy(), r64["yhat"]), y_fp32_oracle=normwise(r32["yhat"], r64["yhat"]),
                     g_kernel=l2rel(g, r64["theta_c_bar"]), g_fp32_oracle=l2rel(r32["theta_c_bar"], r64["theta_c_bar"])))
    print(rows[-1], flush=True)
    for k in env: os.environ.pop(k)
def opn(tag, spec, B, T, env=None, stable=True):
    env = env or {}
    for k, v in env.items(): os.environ[k] = v
    th = O.init_params(spec, 1, stable=stable).astype(np.float32)
    u = np.random.default_rng(0).uniform(-1, 1, (B, T, spec.input_dim)).astype(np.float32)
    tg = np.zeros((B, T + 1, spec.output_dim), np.float32)
    r64 = CO.rollout(spec, th, u, targets=tg, dtype=np.float64); r32 = CO.rollout(spec, th, u, targets=tg, dtype=np.float32)
    y, _, tape = ops.rollout_fwd(conv(spec), T_(th), T_(u), want_tape=True)
    ybar = O.mse_cotangent(tg, y.cpu().numpy())
    g, _ = ops.rollout_bwd(conv(spec), T_(th), T_(u), tape, T_(ybar))
    rows.append(dict(case=tag, B=B, T=T, K=1, env=env, y_kernel=normwise(y.cpu().numpy(), r64["y"]), y_fp32_oracle=normwise(r32["y"], r64["y"]),
                     g_kernel=l2rel(g.cpu().numpy(), r64["theta_bar"]), g_fp32_oracle=l2rel(r32["theta_bar"], r64["theta_bar"])))
    print(rows[-1], flush=True)
    for k in env: os.environ.pop(k)
m50, c5 = model_variants()["compact_s50_d0"], controller_variants()["compact_s5_d0"]
OLD = {"CC_B200_LIN": "0"}
closed("cfg2 (64 x 1001), blocked linear", m50, c5, 64, 1001)
closed("cfg2 (64 x 1001), round-1 kernels", m50, c5, 64, 1001, env=OLD)
closed("cfg2x sample (2048 x 1001), blocked linear", m50, c5, 2048, 1001)
closed("cfg2x sample (2048 x 1001), generic controller path", m50, c5, 2048, 1001, env={"CC_B200_LIN_LC": "0"})
closed("cfg2x sample (2048 x 1001), round-1 kernels", m50, c5, 2048, 1001, env=OLD)
closed("cfg2 with noise", m50, c5, 290, 1001, noise_on=True)
closed("full S=20 post-step model + feed-through controller", tc_model_variants()["full_s20_d0"], O.make_node(6, 2, 1, f_depth=0, g_feedthrough_u=True), 290, 1001)
for K in (1, 32, 100, 128):
    closed("cfg5 (4096 x 10001), contractive controller", m50, c5, 4096, 10001, stable_c=True, K=K)
closed("cfg5 (4096 x 10001), eqx-style random controller (near-marginal trajectories)", m50, c5, 4096, 10001, K=100)
closed("deep model (5,10,2) + CLI controller (25,10,2), FFMA", model_variants()["full_s5_w10_d2"], controller_variants()["full_s25_w10_d2"], 290, 1001)
opn("cfg1 (8 x 1000), blocked linear", m50, 8, 1000)
opn("cfg1 (8 x 1000), round-1 kernels", m50, 8, 1000, env=OLD)
opn("cfg4 sample (2048 x 1000), blocked linear", m50, 2048, 1000)
opn("cfg4 sample (2048 x 1000), round-1 kernels", m50, 2048, 1000, env=OLD)
opn("cfg1 with eqx-style (unstable) init", m50, 8, 1000, stable=False)
opn("sweep (32,32,2) tanh", O.make_node(32, 1, 1, f_width=32, f_depth=2, f_act_final=O.ACT_TANH), 257, 1000, stable=False)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02_parity_floors.json"), "w"), indent=1)
