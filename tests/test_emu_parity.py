"""CPU (no GPU): the device code of chain_control_b200/csrc compiled against the fiber emulation in
tests/emu, called through the same C ABI, compared with the oracle.  Catches indexing / barrier /
layout errors before any GPU time is spent.  Sizes are tiny because every CUDA thread is a fiber."""
import numpy as np
import pytest

import emu_lib as E
from cchelpers import TOL_GRAD, TOL_STATE, conv, l2rel, normwise
from ccspecs import controller_variants, model_variants
from oracle import np_oracle as O


def _u(B, T, I, seed=0):
    return np.random.default_rng(seed).uniform(-1, 1, size=(B, T, I)).astype(np.float32)


@pytest.mark.parametrize("name", list(model_variants()))
def test_emu_rollout_fwd_bwd(name):
    spec = model_variants()[name]
    B, T = 7, 21  # ragged: not a multiple of the 4-trajectory thread tile nor of the 16-step I/O chunk
    theta = O.init_params(spec, 1).astype(np.float32)
    u = _u(B, T, spec.input_dim)
    x0 = (0.1 * np.random.default_rng(3).normal(size=(B, spec.state_dim))).astype(np.float32)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    y64, xs = O.rollout_fwd(spec, theta, u, x0=x0)
    tb64, ub64 = O.rollout_bwd(spec, theta, u, ybar, x0=x0)
    y, tb, ub = E.np_rollout_bwd(E.emu(), conv(spec), theta, u, ybar, x0)
    assert normwise(y, y64) < TOL_STATE
    assert l2rel(tb, tb64) < TOL_GRAD
    assert l2rel(ub, ub64) < TOL_GRAD
    _, xf, _ = E.np_rollout_fwd(E.emu(), conv(spec), theta, u, x0)
    assert normwise(xf, xs[:, -1]) < TOL_STATE


@pytest.mark.parametrize("cname", list(controller_variants()))
@pytest.mark.parametrize("mname", ["compact_s50_d0", "full_s5_w10_d2", "full_s1_tv"])
def test_emu_closedloop_fwd_bwd(cname, mname):
    mspec, cspec = model_variants()[mname], controller_variants()[cname]
    B, Tr = 6, 19
    thc, thm = O.init_params(cspec, 2).astype(np.float32), O.init_params(mspec, 1).astype(np.float32)
    refs = 3 * _u(B, Tr, 1)
    noise = (0.02 * np.random.default_rng(1).normal(size=(B, Tr, 1))).astype(np.float32)
    ybar = np.random.default_rng(6).normal(size=(B, Tr, 1)).astype(np.float32)
    y64, us64 = O.closedloop_fwd(cspec, mspec, thc, thm, refs, noise)
    tb64 = O.closedloop_bwd(cspec, mspec, thc, thm, refs, ybar, noise)
    yhat, us, _ = E.np_closedloop_fwd(E.emu(), conv(cspec), conv(mspec), thc, thm, refs, noise)
    assert normwise(yhat, y64) < TOL_STATE
    assert normwise(us, us64) < TOL_STATE
    _, tb = E.np_closedloop_bwd(E.emu(), conv(cspec), conv(mspec), thc, thm, refs, ybar, noise)
    assert l2rel(tb, tb64) < TOL_GRAD


def test_emu_edge_cases():
    L = E.emu()
    spec = model_variants()["full_s5_w10_d2"]
    theta = O.init_params(spec, 1).astype(np.float32)
    # T = 0: only y0
    y, xf, _ = E.np_rollout_fwd(L, conv(spec), theta, np.zeros((3, 0, 1), np.float32))
    y64, _ = O.rollout_fwd(spec, theta, np.zeros((3, 0, 1)))
    np.testing.assert_allclose(y, y64, rtol=1e-6, atol=1e-7)
    # B = 1, T = exactly one I/O chunk, and one chunk + 1
    for T in (16, 17):
        u = _u(1, T, 1, seed=T)
        y, _, _ = E.np_rollout_fwd(L, conv(spec), theta, u)
        assert normwise(y, O.rollout_fwd(spec, theta, u)[0]) < TOL_STATE
    # three warp tiles with a ragged tail (B = 70 = 32 + 32 + 6) as ONE CTA of three warps and as THREE CTAs of one warp
    # (CC_B200_WPC is the warps-per-CTA override the plan really reads): multi-CTA indexing and the gradient reduction
    import os
    u = _u(70, 9, 1, seed=4)
    ybar = np.random.default_rng(5).normal(size=(70, 10, 1)).astype(np.float32)
    tb64, ub64 = O.rollout_bwd(spec, theta, u, ybar)
    for wpc in ("3", "1"):
        os.environ["CC_B200_WPC"] = wpc
        try:
            y, tb, ub = E.np_rollout_bwd(L, conv(spec), theta, u, ybar)
        finally:
            del os.environ["CC_B200_WPC"]
        assert normwise(y, O.rollout_fwd(spec, theta, u)[0]) < TOL_STATE
        assert l2rel(tb, tb64) < TOL_GRAD and l2rel(ub, ub64) < TOL_GRAD


def test_emu_mse_value_and_cotangent():
    import ctypes as C

    L = E.emu()
    rng = np.random.default_rng(0)
    y, yhat = rng.normal(size=5000).astype(np.float32), rng.normal(size=5000).astype(np.float32)
    ybar = np.empty_like(y)
    part = np.zeros(512, np.float32)
    rc = L.cc_mse_value_and_cotangent(y.ctypes.data, yhat.ctypes.data, ybar.ctypes.data, part.ctypes.data, y.size,
                                      C.c_float(1.0 / y.size), None)
    L.check(rc, "mse")
    assert abs(part[0] - O.mse(y.astype(np.float64), yhat.astype(np.float64))) < 1e-6
    np.testing.assert_allclose(ybar, O.mse_cotangent(y, yhat), rtol=1e-6, atol=1e-9)


# ---- blocked linear family (cc_lin_kernels.cuh, cc_lin_open.cuh): the tensor-memory / MMA layer is emulated with the same
# tf32 truncation of the operands (cc_lin_tm.cuh), so these runs also predict the rounding error of the 3xTF32 products ----
LIN_MODELS = {
    "compact_s50": dict(state_dim=50, input_dim=1, output_dim=1, readout_pre_step=True, u_transform=O.UT_ARCTAN),
    "full_s20_ktanh": dict(state_dim=20, input_dim=1, output_dim=1, u_transform=O.UT_KTANH, u_scale=6.0),
    "compact_s56": dict(state_dim=56, input_dim=1, output_dim=1, readout_pre_step=True, u_transform=O.UT_ARCTAN),
    "euler_s60": dict(state_dim=60, input_dim=1, output_dim=1, integrator=O.INT_EE),
    "full_s51": dict(state_dim=51, input_dim=1, output_dim=1),
}
LIN_OPEN_MODELS = dict(LIN_MODELS)
LIN_OPEN_MODELS.update({
    "compact_s40_i2_o3": dict(state_dim=40, input_dim=2, output_dim=3, readout_pre_step=True, u_transform=O.UT_ARCTAN),
    "full_s51_i4_o4": dict(state_dim=51, input_dim=4, output_dim=4),
    "euler_s18_i3_o2": dict(state_dim=18, input_dim=3, output_dim=2, integrator=O.INT_EE),
    "compact_s44_o2_nobias": dict(state_dim=44, input_dim=1, output_dim=2, readout_pre_step=True, f_use_bias=False, g_use_bias=False),
    "noint_s6_i2_o2": dict(state_dim=6, input_dim=2, output_dim=2, integrator=O.INT_NONE),
})


def _lin_node(kw):
    kw = dict(kw)
    return O.make_node(kw.pop("state_dim"), kw.pop("input_dim"), kw.pop("output_dim"), f_depth=0, **kw)


@pytest.mark.parametrize("mname", list(LIN_MODELS))
@pytest.mark.parametrize("cname", ["compact_s5_d0", "full_s6_d0_ft", "euler_s3_d0"])
@pytest.mark.parametrize("stage", ["0", "1"])
def test_emu_lin_closedloop(mname, cname, stage, monkeypatch):
    import ctypes as C

    from ccspecs import tc_controller_variants

    monkeypatch.setenv("CC_B200_LIN_STAGE", stage)
    mspec, cspec = _lin_node(LIN_MODELS[mname]), tc_controller_variants()[cname]
    L = E.emu()
    assert L.cc_kernel_family_at(C.byref(conv(cspec).to_c()), C.byref(conv(mspec).to_c()), 1, 64) == 4
    B, Tr = 37, 45  # ragged: 2 warps (one partial), last block partial, 1.4 staging chunks
    thc, thm = O.init_params(cspec, 2).astype(np.float32), O.init_params(mspec, 1, stable=True).astype(np.float32)
    refs = 3 * _u(B, Tr, 1)
    noise = (0.02 * np.random.default_rng(1).normal(size=(B, Tr, 1))).astype(np.float32)
    ybar = np.random.default_rng(6).normal(size=(B, Tr, 1)).astype(np.float32)
    y64, us64 = O.closedloop_fwd(cspec, mspec, thc, thm, refs, noise)
    tb64 = O.closedloop_bwd(cspec, mspec, thc, thm, refs, ybar, noise)
    yhat, us, _ = E.np_closedloop_fwd(L, conv(cspec), conv(mspec), thc, thm, refs, noise)
    assert normwise(yhat, y64) < TOL_STATE
    assert normwise(us, us64) < TOL_STATE
    _, tb = E.np_closedloop_bwd(L, conv(cspec), conv(mspec), thc, thm, refs, ybar, noise)
    assert l2rel(tb, tb64) < TOL_GRAD


@pytest.mark.parametrize("mname", list(LIN_OPEN_MODELS))
def test_emu_lin_rollout(mname):
    import ctypes as C

    spec = _lin_node(LIN_OPEN_MODELS[mname])
    L = E.emu()
    assert L.cc_kernel_family_at(None, C.byref(conv(spec).to_c()), 1, 64) == 4
    B, T = 37, 45
    theta = O.init_params(spec, 1, stable=spec.integrator != O.INT_NONE).astype(np.float32)
    if spec.integrator == O.INT_NONE:
        theta = (0.5 * theta).astype(np.float32)
    u = _u(B, T, spec.input_dim)
    x0 = (0.1 * np.random.default_rng(3).normal(size=(B, spec.state_dim))).astype(np.float32)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    y64, xs = O.rollout_fwd(spec, theta, u, x0=x0)
    tb64, ub64 = O.rollout_bwd(spec, theta, u, ybar, x0=x0)
    y, tb, ub = E.np_rollout_bwd(L, conv(spec), theta, u, ybar, x0)
    assert normwise(y, y64) < TOL_STATE
    assert l2rel(tb, tb64) < TOL_GRAD
    assert l2rel(ub, ub64) < TOL_GRAD
    _, xf, _ = E.np_rollout_fwd(L, conv(spec), theta, u, x0)  # no tape: the block length divides T, x_final is exact
    assert normwise(xf, xs[:, -1]) < TOL_STATE


# ---- checkpointed recomputation (ckpt_every > 1): the host replays the forward pass of one segment from its checkpoint into a
# segment-local tape and walks it backwards (cc_abi_bwd.inc); adjoints and gradient sums carry over between the launches ----
@pytest.mark.parametrize("case", ["lin_compact", "lin_post_generic_ctrl", "ffma_deep_tile_ctrl", "ffma_tv_tiny_ctrl"])
def test_emu_closedloop_checkpointed(case):
    from ccspecs import tc_model_variants

    mspec, cspec, B, Tr, Ks = {
        "lin_compact": (model_variants()["compact_s50_d0"], controller_variants()["compact_s5_d0"], 37, 61, (8, 20, 32)),
        "lin_post_generic_ctrl": (tc_model_variants()["full_s20_d0"], controller_variants()["full_s6_ft_tv"], 20, 45, (10,)),
        "ffma_deep_tile_ctrl": (model_variants()["full_s5_w10_d2"], controller_variants()["full_s25_w10_d2"], 9, 30, (8, 7)),
        "ffma_tv_tiny_ctrl": (model_variants()["full_s1_tv"], controller_variants()["compact_s5_d0"], 9, 30, (8,)),
    }[case]
    L = E.emu()
    rng = np.random.default_rng(0)
    thc, thm = O.init_params(cspec, 2).astype(np.float32), O.init_params(mspec, 1, stable=True).astype(np.float32)
    refs = (3 * rng.uniform(-1, 1, (B, Tr, 1))).astype(np.float32)
    noise = (0.02 * rng.normal(size=(B, Tr, 1))).astype(np.float32)
    ybar = rng.normal(size=(B, Tr, 1)).astype(np.float32)
    tb64 = O.closedloop_bwd(cspec, mspec, thc, thm, refs, ybar, noise)
    _, tb1 = E.np_closedloop_bwd(L, conv(cspec), conv(mspec), thc, thm, refs, ybar, noise)
    for K in Ks:
        _, tb = E.np_closedloop_bwd(L, conv(cspec), conv(mspec), thc, thm, refs, ybar, noise, ckpt_every=K)
        assert l2rel(tb, tb64) < TOL_GRAD
        assert l2rel(tb, tb1) < 1e-6  # theta_bar(K) == theta_bar(1) (bit-identical on the FFMA kernels: the replay is exact)


@pytest.mark.parametrize("name", ["full_s5_w10_d2", "full_s12_w25_d1_tanh", "compact_s50_d0"])
def test_emu_rollout_checkpointed(name):
    spec = model_variants()[name]
    L = E.emu()
    rng = np.random.default_rng(1)
    B, T = 9, 30
    theta = O.init_params(spec, 1).astype(np.float32)
    u = rng.uniform(-1, 1, (B, T, spec.input_dim)).astype(np.float32)
    x0 = (0.1 * rng.normal(size=(B, spec.state_dim))).astype(np.float32)
    ybar = rng.normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    tb64, ub64 = O.rollout_bwd(spec, theta, u, ybar, x0=x0)
    _, tb1, ub1 = E.np_rollout_bwd(L, conv(spec), theta, u, ybar, x0)
    for K in (8, 7):
        _, tb, ub = E.np_rollout_bwd(L, conv(spec), theta, u, ybar, x0, ckpt_every=K)
        assert l2rel(tb, tb64) < TOL_GRAD and l2rel(ub, ub64) < TOL_GRAD
        assert l2rel(tb, tb1) < 1e-6 and l2rel(ub, ub1) < 1e-6


def _np_opt_step(theta, mu, nu, grad, t, lr, b1, b2, eps, cn, cv, l1, l2):
    """numpy restatement of optax.chain(clip_by_global_norm, clip, adam) + the reference's regulariser gradients"""
    g = grad.astype(np.float64).copy()
    th = theta.astype(np.float64)
    if l1:
        g += l1 * np.sign(th)
    if l2:
        g += l2 * th / np.sqrt((th * th).sum())
    n = np.sqrt((g * g).sum())
    if cn > 0 and n > cn:
        g *= cn / n
    if cv > 0:
        g = np.clip(g, -cv, cv)
    mu = b1 * mu + (1 - b1) * g
    nu = b2 * nu + (1 - b2) * g * g
    th = th - lr * (mu / (1 - b1 ** t)) / (np.sqrt(nu / (1 - b2 ** t)) + eps)
    return th, mu, nu


@pytest.mark.parametrize("hp", [(1e-3, 0.0, 0.0, 0.0, 0.0), (3e-3, 1.0, 0.0, 0.0, 0.0), (1e-3, 0.5, 0.01, 0.02, 0.05)])
def test_emu_fused_optimiser_step(hp):
    import ctypes as C

    lr, cn, cv, l1, l2 = hp
    L = E.emu()
    rng = np.random.default_rng(0)
    P = 2651
    theta = rng.normal(size=P).astype(np.float32)
    mu, nu = np.zeros(P, np.float32), np.zeros(P, np.float32)
    count = np.zeros(1, np.int32)
    out = np.zeros(3, np.float32)
    th64, mu64, nu64 = theta.astype(np.float64), np.zeros(P), np.zeros(P)
    for t in range(1, 4):
        grad = (rng.normal(size=P) * (3.0 if t == 2 else 0.3)).astype(np.float32)
        th64, mu64, nu64 = _np_opt_step(th64, mu64, nu64, grad, t, lr, 0.9, 0.999, 1e-8, cn, cv, l1, l2)
        rc = L.cc_opt_step(theta.ctypes.data, mu.ctypes.data, nu.ctypes.data, grad.ctypes.data, count.ctypes.data, out.ctypes.data,
                           P, *[C.c_float(v) for v in (lr, 0.9, 0.999, 1e-8, cn, cv, l1, l2, 1.0)], None)
        L.check(rc, "cc_opt_step")
        assert count[0] == t
        np.testing.assert_allclose(theta, th64, rtol=2e-6, atol=2e-7)
        np.testing.assert_allclose(mu, mu64, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("name", ["full_s5_w10_d2", "euler_s9", "full_s32_w32_d2_i2_o3", "noint_s7_w20_d1", "compact_s40_w64_d3_tanh_nobias",
                                  "full_s128_w128_d2", "compact_s70_w100_d1_i2_o2"])
@pytest.mark.parametrize("nh", ["1", "2"])
def test_emu_mlp_family_fwd(name, nh, monkeypatch):
    """deep-MLP tcgen05 forward family (cc_mlp_kernels.cuh) on the CPU emulation of its TMEM / MMA layer: B = 133 (two tiles,
    one ragged), outputs + final state + the tape against the fp64 oracle"""
    import ctypes as C

    from ccspecs import mlp_model_variants

    monkeypatch.setenv("CC_B200_MLP", "1")
    monkeypatch.setenv("CC_B200_MLP_NH", nh)  # one / two threads per trajectory where both kernels exist (padded widths 32, 64)
    spec = mlp_model_variants()[name]
    if nh == "2" and (spec.state_dim > 64 or max(spec.f.sizes[1:]) > 64 or max(spec.state_dim, max(spec.f.sizes[1:])) <= 16):
        pytest.skip("one variant only")
    L = E.emu()
    B, T = 133, (4 if spec.state_dim > 64 else 7)
    assert L.cc_kernel_family_at(None, C.byref(conv(spec).to_c()), 0, B) == 5
    theta = O.init_params(spec, 1).astype(np.float32)
    u = 1.5 * _u(B, T, spec.input_dim)
    x0 = (0.1 * np.random.default_rng(3).normal(size=(B, spec.state_dim))).astype(np.float32)
    y64, xs = O.rollout_fwd(spec, theta, u, x0=x0)
    # trajectories whose outputs nearly cancel put the fp32 oracle itself above 1e-5 in the normwise metric (wide layers):
    # the kernel has to stay within 3x of that floor
    floor = normwise(O.rollout_fwd(spec, theta, u, x0=x0, dtype=np.float32)[0], y64)
    y, xf, tape = E.np_rollout_fwd(L, conv(spec), theta, u, x0, tape_every=3)
    assert normwise(y, y64) < max(TOL_STATE, 3 * floor), floor
    assert normwise(xf, xs[:, -1]) < TOL_STATE
    Bpad, S = (B + 31) // 32 * 32, spec.state_dim
    nck = (T + 2) // 3
    rec = tape[: nck * S * Bpad].reshape(nck, S, Bpad)[:, :, :B]  # checkpoints at steps 0, 3, 6: the pre-step states
    for c, k in enumerate(range(0, T, 3)):
        assert normwise(rec[c].T, xs[:, k]) < TOL_STATE


def test_emu_mlp_family_edge_cases(monkeypatch):
    """deep-MLP family: T = 0 (only y0), a single trajectory, four inputs / four outputs, no initial state"""
    import ctypes as C

    monkeypatch.setenv("CC_B200_MLP", "1")
    L = E.emu()
    for spec in (O.make_node(6, 4, 4, f_width=12, f_depth=2, u_transform=O.UT_KTANH, u_scale=2.0),
                 O.make_node(66, 4, 4, f_width=70, f_depth=1, readout_pre_step=True)):
        assert L.cc_kernel_family_at(None, C.byref(conv(spec).to_c()), 0, 1) == 5
        theta = O.init_params(spec, 4).astype(np.float32)
        for B, T in ((1, 3), (2, 0), (129, 1)):
            u = _u(B, T, spec.input_dim, seed=B)
            y64, xs = O.rollout_fwd(spec, theta, u)
            y, xf, _ = E.np_rollout_fwd(L, conv(spec), theta, u)
            assert y.shape == (B, T + 1, 4)
            assert normwise(y, y64) < TOL_STATE, (spec.state_dim, B, T)
            assert normwise(xf, xs[:, -1]) < TOL_STATE or T == 0


MLPB_CASES = {
    # 3 Linear layers, relu, post-step readout, RK4; two CTAs with a ragged second tile
    "s40_w48_d2": (lambda: O.make_node(40, 1, 1, f_width=48, f_depth=2), 130, 3),
    # 2 Linear layers, tanh + final tanh, pre-step readout, two inputs / outputs, arctan input transform, full 64 columns
    "s64_w64_d1_tanh_pre": (lambda: O.make_node(64, 2, 2, f_width=64, f_depth=1, f_act=O.ACT_TANH, f_act_final=O.ACT_TANH,
                                                readout_pre_step=True, u_transform=O.UT_ARCTAN), 70, 3),
    # no integrator, no biases, three outputs
    "s20_w40_d1_noint": (lambda: O.make_node(20, 1, 3, f_width=40, f_depth=1, integrator=O.INT_NONE, f_use_bias=False,
                                             g_use_bias=False, f_act_final=O.ACT_TANH), 33, 4),
}


@pytest.mark.parametrize("nh", ["2", "4"])
@pytest.mark.parametrize("name", list(MLPB_CASES))
def test_emu_mlpb_reverse(name, nh, monkeypatch):
    """deep-MLP tcgen05 reverse pass (cc_mlpb_kernels.cuh) through the emulated TMEM / MMA layer: round table, operand layouts,
    hi + lo weight-gradient rounds, per-CTA sums and their reduction against the fp64 oracle"""
    import ctypes as C

    monkeypatch.setenv("CC_B200_MLP", "1")
    monkeypatch.setenv("CC_B200_MLPB_NH", nh)  # worker threads per trajectory (2 = default, 4 = the 16-column variant)
    mk, B, T = MLPB_CASES[name]
    spec = mk()
    assert E.emu().cc_kernel_family_at(None, C.byref(conv(spec).to_c()), 1, B) == 6
    theta = O.init_params(spec, 1).astype(np.float32)
    u = _u(B, T, spec.input_dim)
    x0 = (0.1 * np.random.default_rng(3).normal(size=(B, spec.state_dim))).astype(np.float32)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    y64, _ = O.rollout_fwd(spec, theta, u, x0=x0)
    tb64, _ = O.rollout_bwd(spec, theta, u, ybar, x0=x0)
    y, tb, _ = E.np_rollout_bwd(E.emu(), conv(spec), theta, u, ybar, x0, want_ubar=False)
    assert normwise(y, y64) < TOL_STATE
    assert l2rel(tb, tb64) < 1e-5  # (3xTF32 chains + split weight-gradient operands: 1e-7 observed)


def test_emu_mlpb_flush_interval_and_narrow(monkeypatch):
    """flush interval of the weight-gradient accumulators does not change the result beyond rounding; a narrow (width <= 32)
    architecture takes the family only when forced"""
    import ctypes as C

    monkeypatch.setenv("CC_B200_MLP", "1")
    spec = O.make_node(5, 1, 1, f_width=10, f_depth=2)
    m = conv(spec).to_c()
    assert E.emu().cc_kernel_family_at(None, C.byref(m), 1, 64) != 6
    monkeypatch.setenv("CC_B200_MLPB", "1")
    assert E.emu().cc_kernel_family_at(None, C.byref(m), 1, 64) == 6
    B, T = 9, 5
    theta = O.init_params(spec, 1).astype(np.float32)
    u = _u(B, T, 1)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, 1)).astype(np.float32)
    tb64, _ = O.rollout_bwd(spec, theta, u, ybar)
    res = []
    for fl in ("1", "2", "7"):
        monkeypatch.setenv("CC_B200_MLPB_FLUSH", fl)
        _, tb, _ = E.np_rollout_bwd(E.emu(), conv(spec), theta, u, ybar, None, want_ubar=False)
        assert l2rel(tb, tb64) < 1e-5
        res.append(tb)
    assert l2rel(res[0], res[2]) < 1e-6


@pytest.mark.parametrize("name", ["s40_w48_d2", "s64_w64_d1_tanh_pre"])
def test_emu_mlpb_checkpointed_segment_replay(name, monkeypatch):
    """family 6 with ckpt_every > 1: the forward family writes the state every K steps only; the reverse entry point replays each
    segment (family-5 kernel from the checkpoint, no outputs) and walks it backwards on the emulated tensor cores, adjoint and
    per-CTA gradient sums carried between the launches -- same gradient as the unsegmented pass"""
    monkeypatch.setenv("CC_B200_MLP", "1")
    mk, B, _ = MLPB_CASES[name]
    spec = mk()
    T = 5
    theta = O.init_params(spec, 1).astype(np.float32)
    u = _u(B, T, spec.input_dim)
    x0 = (0.1 * np.random.default_rng(3).normal(size=(B, spec.state_dim))).astype(np.float32)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    tb64, _ = O.rollout_bwd(spec, theta, u, ybar, x0=x0)
    _, tb1, _ = E.np_rollout_bwd(E.emu(), conv(spec), theta, u, ybar, x0, want_ubar=False)
    for K in (2, 4, 7):  # 3 segments (2 + 2 + 1), 2 segments (4 + 1), one segment longer than T
        y, tbk, _ = E.np_rollout_bwd(E.emu(), conv(spec), theta, u, ybar, x0, want_ubar=False, ckpt_every=K)
        assert l2rel(tbk, tb64) < 1e-5, K
        assert l2rel(tbk, tb1) < 2e-6, K  # (flush points of the TMEM accumulators move with the segment boundaries)
