import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import np_oracle as O
from oracle import c_oracle as CO
from ccspecs import model_variants, controller_variants
from cchelpers import conv, normwise, l2rel
from chain_control_b200 import ops
mspec, cspec = model_variants()["compact_s50_d0"], controller_variants()["compact_s5_d0"]
thc = O.init_params(cspec, 2).astype(np.float32); thm = O.init_params(mspec, 1, stable=True).astype(np.float32)
B, Tr = 4096, 10001
rng = np.random.default_rng(0)
refs = (np.sign(rng.normal(size=(B, 1, 1))) * rng.uniform(0, 3, size=(B, 1, 1)) * np.ones((B, Tr, 1))).astype(np.float32)
T_ = lambda a: torch.tensor(np.asarray(a, np.float32), device="cuda")
ref = CO.closedloop(cspec, mspec, thc, thm, refs, mse_loss=True, dtype=np.float64)
r32 = CO.closedloop(cspec, mspec, thc, thm, refs, dtype=np.float32)
def per(y): return np.abs(y - ref["yhat"]).max(axis=(1, 2)) / np.abs(ref["yhat"]).max(axis=(1, 2))
e32 = per(r32["yhat"]); i = e32.argmax()
print("fp32 oracle: max", e32.max(), "at traj", i, "amp", refs[i, 0, 0], "max|y|", np.abs(ref["yhat"][i]).max(), "median", np.median(e32))
for env, K in [({}, 1), ({}, 32), ({}, 100), ({"CC_B200_LIN": "0"}, 1), ({"CC_B200_LIN_LC": "0"}, 1)]:
    for k, v in env.items(): os.environ[k] = v
    yh, _, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), want_tape=True, ckpt_every=K)
    e = per(yh.cpu().numpy()); j = e.argmax()
    kk = np.abs(yh.cpu().numpy()[j] - ref["yhat"][j]).argmax()
    print(env, "K", K, ": max", e.max(), "traj", j, "amp", refs[j, 0, 0], "max|y|", np.abs(ref["yhat"][j]).max(), "at step", kk, "median", np.median(e), "n>1e-5:", int((e > 1e-5).sum()))
    for k in env: os.environ.pop(k)
