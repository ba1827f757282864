"""The oracle and the CUDA path against golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE FILES
(tests/golden/make_golden.py + refshim.py; fixtures tests/golden/*.npz are committed, nothing here reads /root/reference).

CPU part (-m "not gpu"): oracle/np_oracle.py and oracle/cc_oracle.c in fp64 must reproduce the reference's outputs,
loss and gradient to round-off (1e-9): this is what pins the oracle.
GPU part (-m gpu): the product through the C ABI against the same fixtures at BASELINE.json's tolerances
(outputs 1e-5 normwise, gradients 1e-4 L2-normwise, fp32 vs the fp64 fixture).
"""
import glob
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as O
from cchelpers import TOL_GRAD, TOL_STATE, conv, l2rel, normwise

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OPEN = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "open_*.npz")))
CLOSED = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "closed_*.npz")))
PIN = 1e-9  # fp64 oracle vs fp64 reference execution


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def spec_of(js):
    return O.make_node(**json.loads(js))


def regu_grad(theta, l1, l2):
    """cotangent of step_fn.py:101-114 with l1_l2_regularisers (cc/utils/utils.py:40-47): l1*|θ|_1 + l2*|θ|_2"""
    g = np.zeros_like(theta)
    if l1:
        g += l1 * np.sign(theta)
    if l2:
        g += l2 * theta / np.linalg.norm(theta)
    return g


def regu_value(theta, l1, l2):
    return l1 * np.abs(theta).sum() + l2 * np.linalg.norm(theta)


def _x0_of(g, B):
    """LinearModel fixtures carry the (possibly random) initial state x0[S] (linear_model.py:41-44)"""
    if "x0" not in g.files or not np.any(g["x0"]):
        return None
    return np.broadcast_to(g["x0"], (B, g["x0"].size)).copy()


def pid_gain_grad(tb, dt):
    """chain rule from the PID node's kernel-order cotangent (tests/golden/make_golden.py: pid_kernel_theta; f rows are
    constants, g = [-d/dt, i, dc, -dc] with dc = p + i dt + d/dt) to the reference's flat order (d_gain, i_gain, p_gain)"""
    gb = tb[8:12]
    dcb = gb[2] - gb[3]
    return np.array([-gb[0] / dt + dcb / dt, gb[1] + dt * dcb, dcb])


def ctrl_theta_and_chain(g):
    """kernel-order controller parameters of a fixture and the map from the kernel-order cotangent back to the reference's
    flat order.  Plain neural-ODE controllers: identity.  PID: pid_gain_grad.  Compact controller + PID secondary
    (superposition): the product's own composition (chain_control_b200/examples/superposition.py), fp64 on the CPU."""
    if "superposed_dims" in g.files:
        import torch

        from chain_control_b200.examples.superposition import superposed_theta

        S, R, O_ = (int(v) for v in g["superposed_dims"])
        flat = torch.tensor(g["theta_c_ref_order"], dtype=torch.float64, requires_grad=True)
        th = superposed_theta(flat, float(g["superposed_alpha"]), S, R, O_, float(g["superposed_dt"]), False)

        def chain(tb):
            (gr,) = torch.autograd.grad(th, flat, grad_outputs=torch.tensor(np.asarray(tb, np.float64)), retain_graph=True)
            return gr.numpy()

        return th.detach().numpy(), chain
    if "pid_gains_dip" in g.files:
        return g["theta_c"], lambda tb: pid_gain_grad(tb, float(g["pid_dt"]))
    return g["theta_c"], lambda tb: tb + regu_grad(g["theta_c"], float(g["l1"]), float(g["l2"]))


def test_fixtures_present():
    assert len(OPEN) >= 6 and len(CLOSED) >= 4, (OPEN, CLOSED)


@pytest.mark.parametrize("name", OPEN)
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_open_loop_pinned(name, impl):
    g = load(name)
    spec = spec_of(str(g["spec"]))
    theta, u, tgt = g["theta"], g["u"], g["targets"]
    assert theta.size == spec.n_params()  # flatten_module order and count (module_utils.py:20-24)
    x0 = _x0_of(g, u.shape[0])
    if impl == "numpy":
        y, _ = O.rollout_fwd(spec, theta, u, np.float64, x0=x0)
        tb, _ = O.rollout_bwd(spec, theta, u, O.mse_cotangent(tgt, y), np.float64, x0=x0)
    else:
        r = CO.rollout(spec, theta, u, x0=x0, dtype=np.float64)
        y = r["y"]
        tb = CO.rollout(spec, theta, u, x0=x0, ybar=O.mse_cotangent(tgt, y), dtype=np.float64)["theta_bar"]
    assert normwise(y, g["y"]) < PIN
    assert abs(O.mse(tgt, y) - float(g["train_mse"])) < PIN * max(1.0, float(g["train_mse"]))
    l1, l2 = float(g["l1"]), float(g["l2"])
    assert abs(O.mse(tgt, y) + regu_value(theta, l1, l2) - float(g["loss"])) < PIN * max(1.0, float(g["loss"]))
    assert l2rel(tb + regu_grad(theta, l1, l2), g["theta_bar"]) < PIN


def _closed_oracle(g, impl):
    cspec = spec_of(str(g["cspec"]))
    mspecs = {k: O.make_node(**v) for k, v in json.loads(str(g["mspecs"])).items()}
    names = json.loads(str(g["model_names"]))
    refs, thc = g["refs"], ctrl_theta_and_chain(g)[0]
    assert thc.size == cspec.n_params()
    tb, losses, yh = np.zeros_like(thc), [], {}
    for n in names:
        thm = g[f"theta_m__{n}"]
        noise = g[f"noise__{n}"] if f"noise__{n}" in g.files else None
        if impl == "numpy":
            y, _ = O.closedloop_fwd(cspec, mspecs[n], thc, thm, refs, noise, np.float64)
            # mean over models of the per-model mse (step_fn.py:42-46,256-271)
            tb += O.closedloop_bwd(cspec, mspecs[n], thc, thm, refs, O.mse_cotangent(refs, y) / len(names), noise,
                                   np.float64)
        else:
            y = CO.closedloop(cspec, mspecs[n], thc, thm, refs, noise=noise, dtype=np.float64)["yhat"]
            tb += CO.closedloop(cspec, mspecs[n], thc, thm, refs, noise=noise,
                                ybar=O.mse_cotangent(refs, y) / len(names), dtype=np.float64)["theta_c_bar"]
        yh[n] = y
        losses.append(O.mse(refs, y))
    return thc, names, yh, losses, tb


@pytest.mark.parametrize("name", CLOSED)
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_closed_loop_pinned(name, impl):
    g = load(name)
    thc, names, yh, losses, tb = _closed_oracle(g, impl)
    for n, l in zip(names, losses):
        assert normwise(yh[n], g[f"yhat__{n}"]) < PIN
        assert abs(l - float(g[f"{n}_train_mse"])) < PIN * max(1.0, l)
    l1, l2 = float(g["l1"]), float(g["l2"])
    assert abs(np.mean(losses) + regu_value(thc, l1, l2) - float(g["loss"])) < PIN * max(1.0, float(g["loss"]))
    special = "pid_gains_dip" in g.files or "superposed_dims" in g.files  # (the d-term divides by dt twice: 1e-7)
    assert l2rel(ctrl_theta_and_chain(g)[1](tb), g["theta_c_bar"]) < (1e-7 if special else PIN)


def _adam_clip_step(theta, grad, lr, clip, b1=0.9, b2=0.999, eps=1e-8):
    """first step of optax.chain(clip_by_global_norm(clip), adam(lr)) (cc/train/test_trainer.py:55)"""
    g = grad * (clip / max(np.linalg.norm(grad), clip))
    m, v = (1 - b1) * g, (1 - b2) * g * g
    return theta - lr * (m / (1 - b1)) / (np.sqrt(v / (1 - b2)) + eps)


@pytest.mark.parametrize("name", OPEN + CLOSED)
def test_fixture_optimizer_step_consistent(name):
    g = load(name)
    k = "theta" if name.startswith("open") else "theta_c"
    th = g["theta_c_ref_order"] if "theta_c_ref_order" in g.files else g[k]
    after = _adam_clip_step(th, g[k + "_bar"], float(g["lr"]), float(g["clip"]))
    assert l2rel(after, g[k + "_after"]) < 1e-12


# ------------------------------------------------------------------- the product kernels on the CPU emulation


@pytest.mark.parametrize("name", [n for n in OPEN if "s64" not in n])
def test_emulated_kernels_open_loop_vs_reference_golden(name):
    """the product's device code (compiled for the CPU fiber emulation, tests/emu) against the reference's outputs/gradients"""
    import emu_lib as E

    g = load(name)
    spec = conv(spec_of(str(g["spec"])))
    theta, u, tgt = g["theta"], g["u"], g["targets"]
    x0 = _x0_of(g, u.shape[0])
    y, _, _ = E.np_rollout_fwd(E.emu(), spec, theta, u, x0)
    assert normwise(y, g["y"]) < TOL_STATE
    ybar = (2.0 / y.size) * (y.astype(np.float64) - tgt)
    _, tb, _ = E.np_rollout_bwd(E.emu(), spec, theta, u, ybar, x0, want_ubar=False)
    assert l2rel(tb.astype(np.float64) + regu_grad(theta, float(g["l1"]), float(g["l2"])), g["theta_bar"]) < TOL_GRAD


@pytest.mark.parametrize("name", CLOSED)
def test_emulated_kernels_closed_loop_vs_reference_golden(name):
    import emu_lib as E

    g = load(name)
    cspec = conv(spec_of(str(g["cspec"])))
    mspecs = {k: conv(O.make_node(**v)) for k, v in json.loads(str(g["mspecs"])).items()}
    names = json.loads(str(g["model_names"]))
    refs, (thc, chain) = g["refs"], ctrl_theta_and_chain(g)
    tb = np.zeros_like(thc)
    for n in names:
        thm = g[f"theta_m__{n}"]
        noise = g[f"noise__{n}"] if f"noise__{n}" in g.files else None
        yhat, _, _ = E.np_closedloop_fwd(E.emu(), cspec, mspecs[n], thc, thm, refs, noise)
        assert normwise(yhat, g[f"yhat__{n}"]) < TOL_STATE
        ybar = (2.0 / (yhat.size * len(names))) * (yhat.astype(np.float64) - refs)
        tb += E.np_closedloop_bwd(E.emu(), cspec, mspecs[n], thc, thm, refs, ybar, noise)[1]
    assert l2rel(chain(tb), g["theta_c_bar"]) < TOL_GRAD


# ----------------------------------------------------------------------------------------------- GPU


def _t(a, dev="cuda:0"):
    import torch

    return torch.tensor(np.asarray(a, np.float32), device=dev)


@pytest.mark.gpu
@pytest.mark.parametrize("name", OPEN)
def test_cuda_open_loop_vs_reference_golden(name):
    from chain_control_b200 import ops

    g = load(name)
    spec = conv(spec_of(str(g["spec"])))
    theta, u, tgt = g["theta"], g["u"], g["targets"]
    x0 = _x0_of(g, u.shape[0])
    x0 = None if x0 is None else _t(x0)
    y, _, tape = ops.rollout_fwd(spec, _t(theta), _t(u), x0, want_tape=True)
    assert normwise(y.cpu().numpy(), g["y"]) < TOL_STATE
    ybar = (2.0 / y.numel()) * (y - _t(tgt))
    # (64-wide deep MLPs: the reverse pass runs on the tcgen05 family 6, cc_mlpb_kernels.cuh)
    tb, _ = ops.rollout_bwd(spec, _t(theta), _t(u), tape, ybar, x0)
    tb = tb.cpu().numpy().astype(np.float64) + regu_grad(theta, float(g["l1"]), float(g["l2"]))
    assert l2rel(tb, g["theta_bar"]) < TOL_GRAD


@pytest.mark.gpu
@pytest.mark.parametrize("name", CLOSED)
def test_cuda_closed_loop_vs_reference_golden(name):
    from chain_control_b200 import ops

    g = load(name)
    cspec = conv(spec_of(str(g["cspec"])))
    mspecs = {k: conv(O.make_node(**v)) for k, v in json.loads(str(g["mspecs"])).items()}
    names = json.loads(str(g["model_names"]))
    refs, (thc, chain) = g["refs"], ctrl_theta_and_chain(g)
    tb = np.zeros_like(thc)
    for n in names:
        thm = g[f"theta_m__{n}"]
        noise = _t(g[f"noise__{n}"]) if f"noise__{n}" in g.files else None
        yhat, _, tape = ops.closedloop_fwd(cspec, mspecs[n], _t(thc), _t(thm), _t(refs), noise=noise, want_tape=True)
        assert normwise(yhat.cpu().numpy(), g[f"yhat__{n}"]) < TOL_STATE
        ybar = (2.0 / (yhat.numel() * len(names))) * (yhat - _t(refs))
        tb += ops.closedloop_bwd(cspec, mspecs[n], _t(thc), _t(thm), _t(refs), tape, ybar, noise=noise).cpu().numpy()
    assert l2rel(chain(tb), g["theta_c_bar"]) < TOL_GRAD
