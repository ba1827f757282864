"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI (chain_control_b200.ops is a
thin ctypes wrapper), against the oracle on the same seeded inputs, plus size-independent properties at
BASELINE.json's full sizes.  Tolerances (written here, from BASELINE.json north_star):
  outputs/states: normwise 1e-5 (per-trajectory max-abs error / max-abs value), fp32 vs the fp64 oracle
  gradients:      L2-normwise 1e-4
"""
import os

import numpy as np
import pytest

from cchelpers import TOL_GRAD, TOL_STATE, conv, l2rel, normwise
from ccspecs import (controller_variants, mlp_model_variants, model_variants, tc_controller_variants, tc_model_variants,
                     tc_open_model_variants)
from oracle import c_oracle as CO
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T_():
    import torch

    return lambda a: torch.tensor(np.asarray(a, np.float32), device="cuda")


FAMILY_ENV = {
    # kernel family -> (CC_B200_TC, CC_B200_WPT_MAXB, CC_B200_LIN): which kernels the plan may pick where the architecture allows it
    "lin": ("1", "0", "1"),                 # blocked linear kernels (the default for linear depth-0 models; others fall through)
    "tcgen05": ("1", "0", "0"),             # tensor-core kernels at every batch size
    "latency": ("1", str(1 << 40), "0"),    # one warp per trajectory at every batch size (default: B <= 1184; ModelTrainer reverse pass <= 12288)
    "ffma": ("0", "0", "0"),                # register-tiled FP32 kernels
}
FAMILY_KEYS = ("CC_B200_TC", "CC_B200_WPT_MAXB", "CC_B200_LIN")


@pytest.fixture(params=list(FAMILY_ENV))
def path(request):
    """run the test once per kernel family (where the architecture is not eligible the plan falls through to FFMA)"""
    old = {k: os.environ.get(k) for k in FAMILY_KEYS}
    for k, v in zip(FAMILY_KEYS, FAMILY_ENV[request.param]):
        os.environ[k] = v
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _u(B, T, I, seed=0):
    return np.random.default_rng(seed).uniform(-1, 1, size=(B, T, I)).astype(np.float32)


@pytest.mark.parametrize("name", list(model_variants()))
def test_rollout_fwd_bwd_vs_oracle(name, T_):
    from chain_control_b200 import ops

    spec = model_variants()[name]
    B, T = 301, 203
    theta = O.init_params(spec, 1).astype(np.float32)
    u = _u(B, T, spec.input_dim)
    x0 = (0.1 * np.random.default_rng(3).normal(size=(B, spec.state_dim))).astype(np.float32)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    ref = CO.rollout(spec, theta, u, x0=x0, ybar=ybar, want_ubar=True, dtype=np.float64)
    y, xf, tape = ops.rollout_fwd(conv(spec), T_(theta), T_(u), T_(x0), want_tape=True, want_x_final=True)
    tb, ub = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar), T_(x0), want_ubar=True)
    assert normwise(y.cpu().numpy(), ref["y"]) < TOL_STATE
    assert l2rel(tb.cpu().numpy(), ref["theta_bar"]) < TOL_GRAD
    # relu kinks make the input cotangent of deep relu nets sensitive to fp32 mask flips: 5x slack
    assert l2rel(ub.cpu().numpy(), ref["ubar"]) < 5 * TOL_GRAD


@pytest.mark.parametrize("cname", list(controller_variants()))
@pytest.mark.parametrize("mname", ["compact_s50_d0", "full_s5_w10_d2", "full_s1_tv"])
def test_closedloop_fwd_bwd_vs_oracle(cname, mname, T_):
    from chain_control_b200 import ops

    mspec, cspec = model_variants()[mname], controller_variants()[cname]
    B, Tr = 290, 201
    thc, thm = O.init_params(cspec, 2).astype(np.float32), O.init_params(mspec, 1).astype(np.float32)
    refs = 3 * _u(B, Tr, 1)
    noise = (0.02 * np.random.default_rng(1).normal(size=(B, Tr, 1))).astype(np.float32)
    ybar = np.random.default_rng(6).normal(size=(B, Tr, 1)).astype(np.float32)
    ref = CO.closedloop(cspec, mspec, thc, thm, refs, noise, ybar=ybar, dtype=np.float64)
    # fp32 restatement = "what the reference's fp32 path computes"; its distance to fp64 is the noise floor of the
    # problem itself (relu masks flip between fp32 and fp64 when noise pushes pre-activations across 0)
    ref32 = CO.closedloop(cspec, mspec, thc, thm, refs, noise, ybar=ybar, dtype=np.float32)
    floor = l2rel(ref32["theta_c_bar"], ref["theta_c_bar"])
    yh, us, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), T_(noise), want_tape=True,
                                      want_u=True)
    tb = ops.closedloop_bwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), tape, T_(ybar), T_(noise))
    assert normwise(yh.cpu().numpy(), ref["yhat"]) < TOL_STATE
    assert normwise(us.cpu().numpy(), ref["u"]) < TOL_STATE
    assert l2rel(tb.cpu().numpy(), ref["theta_c_bar"]) < max(TOL_GRAD, 2 * floor), (floor,)


@pytest.mark.parametrize("name", list(tc_open_model_variants()))
def test_tc_rollout_vs_oracle(name, path, T_):
    """depth-0 models, open loop: forward on tcgen05 (3xTF32), reverse pass on the FFMA kernels from the tape it wrote"""
    from chain_control_b200 import ops

    spec = tc_open_model_variants()[name]
    B, T = 301, 203
    theta = O.init_params(spec, 1).astype(np.float32)
    u = _u(B, T, spec.input_dim)
    x0 = (0.1 * np.random.default_rng(3).normal(size=(B, spec.state_dim))).astype(np.float32)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    ref = CO.rollout(spec, theta, u, x0=x0, ybar=ybar, want_ubar=True, dtype=np.float64)
    y, xf, tape = ops.rollout_fwd(conv(spec), T_(theta), T_(u), T_(x0), want_tape=True, want_x_final=True)
    tb, ub = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar), T_(x0), want_ubar=True)
    assert normwise(y.cpu().numpy(), ref["y"]) < TOL_STATE
    assert bool(np.isfinite(xf.cpu().numpy()).all())
    assert l2rel(tb.cpu().numpy(), ref["theta_bar"]) < TOL_GRAD
    assert l2rel(ub.cpu().numpy(), ref["ubar"]) < TOL_GRAD


@pytest.mark.parametrize("cname", list(tc_controller_variants()))
@pytest.mark.parametrize("mname", list(tc_model_variants()))
def test_tc_closedloop_vs_oracle(cname, mname, path, T_):
    """depth-0 models in closed loop with the tiny controllers: forward and (linear models) reverse pass on tcgen05"""
    from chain_control_b200 import ops

    mspec, cspec = tc_model_variants()[mname], tc_controller_variants()[cname]
    B, Tr = 290, 201
    thc, thm = O.init_params(cspec, 2).astype(np.float32), O.init_params(mspec, 1).astype(np.float32)
    refs = 3 * _u(B, Tr, 1)
    noise = (0.02 * np.random.default_rng(1).normal(size=(B, Tr, 1))).astype(np.float32)
    ybar = np.random.default_rng(6).normal(size=(B, Tr, 1)).astype(np.float32)
    ref = CO.closedloop(cspec, mspec, thc, thm, refs, noise, ybar=ybar, dtype=np.float64)
    yh, us, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), T_(noise), want_tape=True,
                                      want_u=True)
    tb = ops.closedloop_bwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), tape, T_(ybar), T_(noise))
    assert normwise(yh.cpu().numpy(), ref["yhat"]) < TOL_STATE
    assert normwise(us.cpu().numpy(), ref["u"]) < TOL_STATE
    assert l2rel(tb.cpu().numpy(), ref["theta_c_bar"]) < TOL_GRAD


def test_tc_partial_tile_and_empty(T_):
    """B not a multiple of the 128-trajectory tile, B = 1, T = 0 (only y0) on the tcgen05 path"""
    import torch

    from chain_control_b200 import ops

    spec = tc_model_variants()["compact_s50_d0"]
    theta = O.init_params(spec, 1).astype(np.float32)
    for B, T in [(1, 17), (129, 5), (3, 0)]:
        u = _u(B, T, 1)
        ref = CO.rollout(spec, theta, u, dtype=np.float64)
        y, _, _ = ops.rollout_fwd(conv(spec), T_(theta), torch.tensor(u, device="cuda").reshape(B, T, 1))
        assert tuple(y.shape) == (B, T + 1, 1)
        assert normwise(y.cpu().numpy(), ref["y"]) < TOL_STATE


def test_cfg1_model_trainer_shapes(path, T_):
    """configs[0]: 8 trajectories x 1000 steps (outputs 1001), notebook-3 model, T=1000 horizon."""
    from chain_control_b200 import ops

    spec = model_variants()["compact_s50_d0"]
    theta = O.init_params(spec, 1).astype(np.float32)
    u = _u(8, 1000, 1)
    targets = np.random.default_rng(2).normal(size=(8, 1001, 1)).astype(np.float32)
    ref = CO.rollout(spec, theta, u, targets=targets, dtype=np.float64)
    # eqx-style random init is mildly unstable over 1000 steps: the fp32 restatement's own distance to fp64 is the
    # noise floor of the problem (0.6-5e-6, more for unlucky seeds); the kernel must stay within it
    floor = normwise(CO.rollout(spec, theta, u, dtype=np.float32)["y"], ref["y"])
    y, _, tape = ops.rollout_fwd(conv(spec), T_(theta), T_(u), want_tape=True)
    ybar = O.mse_cotangent(targets, y.cpu().numpy())
    tb, _ = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar))
    assert tuple(y.shape) == (8, 1001, 1)
    assert normwise(y.cpu().numpy(), ref["y"]) < max(TOL_STATE, 3 * floor), (floor,)
    assert l2rel(tb.cpu().numpy(), ref["theta_bar"]) < TOL_GRAD


def test_cfg2_controller_trainer_64x1001(path, T_):
    """configs[1]: closed-loop BPTT over 64 reference trajectories x 1001 steps (mse(refs, yhat))."""
    from chain_control_b200 import ops

    mspec, cspec = model_variants()["compact_s50_d0"], controller_variants()["compact_s5_d0"]
    thc = O.init_params(cspec, 2).astype(np.float32)
    thm = O.init_params(mspec, 1, stable=True).astype(np.float32)
    rng = np.random.default_rng(0)
    refs = (np.sign(rng.normal(size=(64, 1, 1))) * rng.uniform(0, 3, size=(64, 1, 1)) * np.ones((64, 1001, 1))).astype(np.float32)
    ref = CO.closedloop(cspec, mspec, thc, thm, refs, mse_loss=True, dtype=np.float64)
    yh, _, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), want_tape=True)
    ybar = O.mse_cotangent(refs, yh.cpu().numpy())
    tb = ops.closedloop_bwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), tape, T_(ybar))
    assert normwise(yh.cpu().numpy(), ref["yhat"]) < TOL_STATE
    assert l2rel(tb.cpu().numpy(), ref["theta_c_bar"]) < TOL_GRAD


def test_full_size_properties(path, T_):
    """B = 65,536 (configs[3]) is too big for the oracle; check size-independent properties instead."""
    import torch

    from chain_control_b200 import ops

    spec = conv(model_variants()["compact_s50_d0"])
    theta = T_(O.init_params(model_variants()["compact_s50_d0"], 1, stable=True))
    B, T = 65536, 128
    torch.manual_seed(0)
    u = torch.rand(B, T, 1, device="cuda") * 2 - 1
    y, _, tape = ops.rollout_fwd(spec, theta, u, want_tape=True)
    # 1. batch independence: any sub-batch alone reproduces its rows (different tiling of the batch)
    idx = torch.arange(1000, 1300, device="cuda")
    ys, _, _ = ops.rollout_fwd(spec, theta, u[idx].contiguous())
    assert normwise(ys.cpu().numpy(), y[idx].cpu().numpy()) < 2e-6
    # 2. determinism: reruns are bit-identical (fixed reduction order)
    ybar = torch.randn(B, T + 1, 1, device="cuda") / B
    tb1, _ = ops.rollout_bwd(spec, theta, u, tape, ybar)
    tb2, _ = ops.rollout_bwd(spec, theta, u, tape, ybar)
    assert torch.equal(tb1, tb2)
    # 3. the VJP is linear in the cotangent, and additive over batch shards (what the NCCL allreduce relies on)
    tb3, _ = ops.rollout_bwd(spec, theta, u, tape, 2 * ybar)
    assert l2rel(tb3.cpu().numpy(), 2 * tb1.cpu().numpy()) < 1e-6
    h = B // 2
    ya, _, ta = ops.rollout_fwd(spec, theta, u[:h].contiguous(), want_tape=True)
    yb, _, tb_ = ops.rollout_fwd(spec, theta, u[h:].contiguous(), want_tape=True)
    ga, _ = ops.rollout_bwd(spec, theta, u[:h].contiguous(), ta, ybar[:h].contiguous())
    gb, _ = ops.rollout_bwd(spec, theta, u[h:].contiguous(), tb_, ybar[h:].contiguous())
    assert l2rel((ga + gb).cpu().numpy(), tb1.cpu().numpy()) < 1e-5
    # 4. decomposition independence: a sub-batch that starts in the middle of a tile and ends with a partial one
    y64, _, _ = ops.rollout_fwd(spec, theta, u[77:4096 + 13].contiguous())
    assert normwise(y64.cpu().numpy(), y[77:4096 + 13].cpu().numpy()) < 2e-6


def test_full_size_closed_loop_properties(path, T_):
    """closed loop at the bench size (65,536 trajectories; T_r shortened to keep the test fast): size-independent
    properties of the forward rollout and of the controller gradient, per kernel family"""
    import torch

    from chain_control_b200 import ops

    mspec, cspec = model_variants()["compact_s50_d0"], controller_variants()["compact_s5_d0"]
    model, ctrl = conv(mspec), conv(cspec)
    thm = T_(O.init_params(mspec, 1, stable=True))
    thc = T_(O.init_params(cspec, 2))
    B, Tr = 65536, 96
    amp = torch.sign(torch.randn(B, 1, 1, device="cuda")) * torch.rand(B, 1, 1, device="cuda") * 3
    refs = amp.expand(B, Tr, 1).contiguous()
    yh, _, tape = ops.closedloop_fwd(ctrl, model, thc, thm, refs, want_tape=True)
    # 1. batch independence: a sub-batch alone (different tiles, partial tile) reproduces its rows
    idx = torch.arange(777, 1100, device="cuda")
    ys, _, _ = ops.closedloop_fwd(ctrl, model, thc, thm, refs[idx].contiguous())
    assert normwise(ys.cpu().numpy(), yh[idx].cpu().numpy()) < 2e-6
    # 2. a sub-batch against the fp64 oracle (the only part small enough for it)
    ref = CO.closedloop(cspec, mspec, thc.cpu().numpy(), thm.cpu().numpy(), refs[idx].cpu().numpy(), dtype=np.float64)
    assert normwise(ys.cpu().numpy(), ref["yhat"]) < TOL_STATE
    # 3. determinism of the gradient (fixed reduction order), linearity in the cotangent, additivity over batch shards
    ybar = torch.randn(B, Tr, 1, device="cuda") / B
    g1 = ops.closedloop_bwd(ctrl, model, thc, thm, refs, tape, ybar)
    g2 = ops.closedloop_bwd(ctrl, model, thc, thm, refs, tape, ybar)
    assert torch.equal(g1, g2)
    g3 = ops.closedloop_bwd(ctrl, model, thc, thm, refs, tape, 2 * ybar)
    assert l2rel(g3.cpu().numpy(), 2 * g1.cpu().numpy()) < 1e-6
    h = B // 2
    _, _, ta = ops.closedloop_fwd(ctrl, model, thc, thm, refs[:h].contiguous(), want_tape=True)
    _, _, tb = ops.closedloop_fwd(ctrl, model, thc, thm, refs[h:].contiguous(), want_tape=True)
    ga = ops.closedloop_bwd(ctrl, model, thc, thm, refs[:h].contiguous(), ta, ybar[:h].contiguous())
    gb = ops.closedloop_bwd(ctrl, model, thc, thm, refs[h:].contiguous(), tb, ybar[h:].contiguous())
    assert l2rel((ga + gb).cpu().numpy(), g1.cpu().numpy()) < 1e-5
    # 4. the kernel families agree with each other at full size (outputs and gradient): compare with the FFMA kernels
    saved = {k: os.environ.get(k) for k in FAMILY_KEYS}
    for k, v in zip(FAMILY_KEYS, FAMILY_ENV["ffma" if path != "ffma" else "tcgen05"]):
        os.environ[k] = v
    try:
        yo, _, tapeo = ops.closedloop_fwd(ctrl, model, thc, thm, refs, want_tape=True)
        go = ops.closedloop_bwd(ctrl, model, thc, thm, refs, tapeo, ybar)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert normwise(yo.cpu().numpy(), yh.cpu().numpy()) < 5e-6
    assert l2rel(go.cpu().numpy(), g1.cpu().numpy()) < 2e-5


def test_zero_controller_equals_open_loop(T_):
    import torch

    from chain_control_b200 import ops

    mspec, cspec = model_variants()["full_s5_w10_d2"], controller_variants()["compact_s5_d0"]
    thm = T_(O.init_params(mspec, 1))
    thc = torch.zeros(cspec.n_params(), device="cuda")
    refs = T_(_u(33, 77, 1))
    yh, us, _ = ops.closedloop_fwd(conv(cspec), conv(mspec), thc, thm, refs, want_u=True)
    y, _, _ = ops.rollout_fwd(conv(mspec), thm, torch.zeros(33, 77, 1, device="cuda"))
    assert float(us.abs().max()) == 0.0
    np.testing.assert_allclose(yh.cpu().numpy(), y[:, 1:].cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_product_path_has_no_cpu_fallback():
    import torch

    from chain_control_b200 import _lib, ops

    spec = conv(model_variants()["full_s5_w10_d2"])
    with pytest.raises(_lib.CcError):
        ops.rollout_fwd(spec, torch.zeros(spec.n_params()), torch.zeros(2, 3, 1))


def test_host_feed_delivers_batches_in_order():
    """parallel.HostFeed: double-buffered pinned-host -> device copies on a side stream keep order and contents"""
    import torch

    from chain_control_b200 import parallel

    feed = parallel.HostFeed("cuda")
    host = [torch.full((257, 33, 1), float(i)).pin_memory() for i in range(7)]
    feed.push(host[0])
    seen = []
    for i in range(7):
        x = feed.pop()
        if i + 1 < 7:
            feed.push(host[i + 1])
        seen.append((x * 2).sum())  # a consumer on the compute stream
        feed.release()
    torch.cuda.synchronize()
    assert [float(s) for s in seen] == [2.0 * i * 257 * 33 for i in range(7)]


# ---- configs[4]: long-horizon closed-loop BPTT with checkpointed recomputation ------------------------------------------
@pytest.mark.parametrize("B", [64, 4096])
def test_cfg5_long_horizon_checkpointed(B, T_):
    """T_r = 10,001: yhat and theta_c_bar against the fp64 oracle for ckpt_every in {1, 32, 100, 128}, and
    theta_bar(K) == theta_bar(1) within 1e-6 (the segments are replayed from their checkpoints; the blocked linear kernels change
    their block length with K, which moves the rounding at the 1e-7 level).
    Contractive synthetic parameters for model AND controller: with the eqx-style random controller the closed
    loop has near-marginal trajectories (|ref| ~ 0.347) on which the fp32 restatement itself is 1e-4 (yhat) / 7 % (gradient) off
    fp64.  Even so, over 10,001 steps the fp32 restatement's gradient sits 2-4e-4 from fp64: the kernels are held to the
    tolerance or 3x that floor, whichever is larger (floors recorded in profiles/r02_parity_floors.md)."""
    from chain_control_b200 import ops

    mspec, cspec = model_variants()["compact_s50_d0"], controller_variants()["compact_s5_d0"]
    thc = O.init_params(cspec, 2, stable=True).astype(np.float32)
    thm = O.init_params(mspec, 1, stable=True).astype(np.float32)
    Tr = 10001
    rng = np.random.default_rng(0)
    refs = (np.sign(rng.normal(size=(B, 1, 1))) * rng.uniform(0, 3, size=(B, 1, 1)) * np.ones((B, Tr, 1))).astype(np.float32)
    ref = CO.closedloop(cspec, mspec, thc, thm, refs, mse_loss=True, dtype=np.float64)
    r32 = CO.closedloop(cspec, mspec, thc, thm, refs, mse_loss=True, dtype=np.float32)
    floor_y, floor_g = normwise(r32["yhat"], ref["yhat"]), l2rel(r32["theta_c_bar"], ref["theta_c_bar"])
    def run(K, m_block=None):
        if m_block:
            os.environ["CC_B200_LIN_M"] = str(m_block)
        try:
            yh, _, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), want_tape=True, ckpt_every=K)
            ybar = O.mse_cotangent(refs, yh.cpu().numpy())
            tb = ops.closedloop_bwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), tape, T_(ybar), ckpt_every=K).cpu().numpy()
        finally:
            os.environ.pop("CC_B200_LIN_M", None)
        return yh.cpu().numpy(), tb

    yh1, g1 = run(1)
    for K in (1, 32, 100, 128):
        yh, tb = run(K)
        assert normwise(yh, ref["yhat"]) < max(TOL_STATE, 3 * floor_y), (K, floor_y)
        assert l2rel(tb, ref["theta_c_bar"]) < max(TOL_GRAD, 3 * floor_g), (K, floor_g)
        if K > 1:
            # against ckpt_every = 1: bit-identical on the FFMA kernels (exact replay, test_checkpointed_general_architectures);
            # here every checkpoint materialises x_m in fp32 where the unsegmented run keeps the 3xTF32 operand chain, and the
            # block length follows K (8 or 10 steps instead of 13): 1e-5 after 10,001 steps and up to 1,250 checkpoints, 20x
            # below the fp32 restatement's own distance to fp64 at this horizon (2-4e-4); short horizons: 2e-7 (test_emu_parity)
            assert l2rel(tb, g1) < 3e-5, K


@pytest.mark.parametrize("mname,cname", [("full_s5_w10_d2", "full_s25_w10_d2"), ("euler_s30_d0_tanh", "compact_s5_d0")])
def test_checkpointed_general_architectures(mname, cname, T_):
    """ckpt_every > 1 on the FFMA kernels (deep MLPs, tile-sized controller, non-linear depth-0 model): exact replay"""
    from chain_control_b200 import ops

    mspec = dict(model_variants(), **tc_model_variants())[mname]
    cspec = controller_variants()[cname]
    B, Tr = 290, 1001
    thc, thm = O.init_params(cspec, 2).astype(np.float32), O.init_params(mspec, 1).astype(np.float32)
    refs = 3 * _u(B, Tr, 1)
    noise = (0.02 * np.random.default_rng(1).normal(size=(B, Tr, 1))).astype(np.float32)
    ybar = (np.random.default_rng(6).normal(size=(B, Tr, 1)) / Tr).astype(np.float32)
    yh1, _, tape1 = ops.closedloop_fwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), T_(noise), want_tape=True)
    g1 = ops.closedloop_bwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), tape1, T_(ybar), T_(noise)).cpu().numpy()
    for K in (8, 64, 100):
        yh, _, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), T_(noise), want_tape=True, ckpt_every=K)
        g = ops.closedloop_bwd(conv(cspec), conv(mspec), T_(thc), T_(thm), T_(refs), tape, T_(ybar), T_(noise), ckpt_every=K).cpu().numpy()
        assert tape.numel() < tape1.numel() / 2
        # (K > 1 selects the FFMA kernels for both passes; K = 1 may run the forward pass on another family)
        assert normwise(yh.cpu().numpy(), yh1.cpu().numpy()) < 2e-5
        assert l2rel(g, g1) < 2e-5, K


def test_model_trainer_long_horizon_runs(T_):
    """ModelTrainer at 65,536 x 10,000 (131 GB of residuals with a store-everything scan): the blocked linear family keeps
    the state at the block starts only (10 GB); gradient finite, linear in the cotangent, additive over batch shards"""
    import torch

    from chain_control_b200 import ops

    spec = conv(model_variants()["compact_s50_d0"])
    theta = T_(O.init_params(model_variants()["compact_s50_d0"], 1, stable=True))
    B, T = 65536, 10000
    torch.manual_seed(0)
    u = torch.rand(B, T, 1, device="cuda") * 2 - 1
    y, _, tape = ops.rollout_fwd(spec, theta, u, want_tape=True, ckpt_every=100)
    assert tape.numel() * 4 < 12e9
    ybar = (2.0 / y.numel()) * y
    g, _ = ops.rollout_bwd(spec, theta, u, tape, ybar, ckpt_every=100)
    assert bool(torch.isfinite(g).all()) and float(g.norm()) > 0
    # a sub-batch against the fp64 oracle (forward) and its gradient share by additivity
    idx = slice(1000, 1064)
    ref = CO.rollout(model_variants()["compact_s50_d0"], theta.cpu().numpy(), u[idx].cpu().numpy(), dtype=np.float64)
    assert normwise(y[idx].cpu().numpy(), ref["y"]) < TOL_STATE
    h = B // 2
    ya, _, ta = ops.rollout_fwd(spec, theta, u[:h].contiguous(), want_tape=True)
    ga, _ = ops.rollout_bwd(spec, theta, u[:h].contiguous(), ta, ybar[:h].contiguous())
    del ta
    yb, _, tb_ = ops.rollout_fwd(spec, theta, u[h:].contiguous(), want_tape=True)
    gb, _ = ops.rollout_bwd(spec, theta, u[h:].contiguous(), tb_, ybar[h:].contiguous())
    assert l2rel((ga + gb).cpu().numpy(), g.cpu().numpy()) < 1e-5


def test_mse_value_and_cotangent_gpu(T_):
    """K6 on the GPU (it runs inside bench's timed step): value and cotangent against numpy, odd sizes and unaligned views"""
    import torch

    from chain_control_b200 import ops

    rng = np.random.default_rng(0)
    for n, off in [(5000, 0), (65536 * 33 + 7, 0), (4097, 1)]:
        y = rng.normal(size=n + off).astype(np.float32)
        yhat = rng.normal(size=n + off).astype(np.float32)
        yd, hd = T_(y)[off:], T_(yhat)[off:]
        loss, ybar = ops.mse_value_and_cotangent(yd, hd)
        assert abs(float(loss) - O.mse(y[off:].astype(np.float64), yhat[off:].astype(np.float64))) < 2e-6
        np.testing.assert_allclose(ybar.cpu().numpy(), O.mse_cotangent(y[off:], yhat[off:]), rtol=1e-6, atol=1e-9)
    # sharded mean: the partial of a shard is scaled by the GLOBAL count
    loss, ybar = ops.mse_value_and_cotangent(yd, hd, n_total=3 * n)
    assert abs(float(loss) * 3 - O.mse(y[off:].astype(np.float64), yhat[off:].astype(np.float64))) < 2e-6


@pytest.mark.parametrize("hp", [(1e-3, 0.0, 0.0, 0.0, 0.0), (3e-3, 1.0, 0.0, 0.0, 0.0), (1e-3, 0.5, 0.01, 0.02, 0.05)])
def test_fused_optimiser_step_gpu(hp, T_):
    """cc_opt_step (regulariser gradients -> clip_by_global_norm -> clip -> adam, one launch) against the numpy restatement of
    optax.chain(clip_by_global_norm, clip, adam) (cc/utils/high_level/defaults.py:43-49)"""
    import torch

    from chain_control_b200 import ops
    from test_emu_parity import _np_opt_step

    lr, cn, cv, l1, l2 = hp
    rng = np.random.default_rng(0)
    P = 2651
    theta = rng.normal(size=P).astype(np.float32)
    opt = ops.FusedAdam(T_(theta), lr, clip_global_norm=cn, clip_value=cv, l1=l1, l2=l2)
    th64, mu64, nu64 = theta.astype(np.float64), np.zeros(P), np.zeros(P)
    for t in range(1, 6):
        grad = (rng.normal(size=P) * (3.0 if t % 2 == 0 else 0.3)).astype(np.float32)
        th64, mu64, nu64 = _np_opt_step(th64, mu64, nu64, grad, t, lr, 0.9, 0.999, 1e-8, cn, cv, l1, l2)
        opt.step(T_(grad))
        assert int(opt.count.item()) == t
        np.testing.assert_allclose(opt.theta.cpu().numpy(), th64, rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize("S,W", [(32, 32), (64, 64)])
def test_sweep_architectures_depth2(S, W, T_):
    """configs[2]'s deep sweep architectures (S = W, depth 2), forward and forward + BPTT against the fp64 oracle"""
    from chain_control_b200 import ops

    spec = O.make_node(S, 1, 1, f_width=W, f_depth=2, f_act_final=O.ACT_TANH)
    B, T = 257, 120
    theta = O.init_params(spec, 1).astype(np.float32)
    u = _u(B, T, 1)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, 1)).astype(np.float32)
    ref = CO.rollout(spec, theta, u, ybar=ybar, want_ubar=True, dtype=np.float64)
    r32 = CO.rollout(spec, theta, u, ybar=ybar, want_ubar=True, dtype=np.float32)
    floor = l2rel(r32["theta_bar"], ref["theta_bar"])
    y, _, tape = ops.rollout_fwd(conv(spec), T_(theta), T_(u), want_tape=True)
    assert normwise(y.cpu().numpy(), ref["y"]) < TOL_STATE
    if W > 32:  # 64-wide deep MLPs: reverse pass on the tcgen05 family 6 (no input cotangent there: never a silent fallback)
        from chain_control_b200 import _lib

        with pytest.raises(_lib.CcError):
            ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar), want_ubar=True)
        tb, _ = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar))
        assert l2rel(tb.cpu().numpy(), ref["theta_bar"]) < max(TOL_GRAD, 3 * floor), floor
        return
    tb, ub = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar), want_ubar=True)
    assert l2rel(tb.cpu().numpy(), ref["theta_bar"]) < max(TOL_GRAD, 3 * floor), floor


@pytest.mark.parametrize("closed", [True, False])
def test_full_size_gradient_vs_oracle(closed, T_):
    """theta_bar at the size the bench runs (65,536 trajectories, full horizon: every CTA slot busy, several waves): the
    cotangent is non-zero on 512 scattered trajectories only, so the fp64 oracle can compute the SAME gradient from that
    sub-batch (trajectories are independent)"""
    import torch

    from chain_control_b200 import ops

    mspec, cspec = model_variants()["compact_s50_d0"], controller_variants()["compact_s5_d0"]
    thm = O.init_params(mspec, 1, stable=True).astype(np.float32)
    thc = O.init_params(cspec, 2).astype(np.float32)
    B, T = 65536, 1001 if closed else 1000
    rng = np.random.default_rng(11)
    idx = np.sort(rng.choice(B, 512, replace=False))
    torch.manual_seed(1)
    if closed:
        amp = torch.sign(torch.randn(B, 1, 1, device="cuda")) * torch.rand(B, 1, 1, device="cuda") * 3
        x = amp.expand(B, T, 1).contiguous()
    else:
        x = torch.rand(B, T, 1, device="cuda") * 2 - 1
    rows = T if closed else T + 1
    ybar = torch.zeros(B, rows, 1, device="cuda")
    yb_sub = rng.normal(size=(512, rows, 1)).astype(np.float32) / 512
    ybar[torch.tensor(idx, device="cuda")] = T_(yb_sub)
    x_sub = x[torch.tensor(idx, device="cuda")].cpu().numpy()
    if closed:
        yh, _, tape = ops.closedloop_fwd(conv(cspec), conv(mspec), T_(thc), T_(thm), x, want_tape=True)
        g = ops.closedloop_bwd(conv(cspec), conv(mspec), T_(thc), T_(thm), x, tape, ybar).cpu().numpy()
        ref = CO.closedloop(cspec, mspec, thc, thm, x_sub, ybar=yb_sub, dtype=np.float64)
        assert normwise(yh[torch.tensor(idx, device="cuda")].cpu().numpy(), ref["yhat"]) < TOL_STATE
        assert l2rel(g, ref["theta_c_bar"]) < TOL_GRAD
    else:
        y, _, tape = ops.rollout_fwd(conv(mspec), T_(thm), x, want_tape=True)
        g, _ = ops.rollout_bwd(conv(mspec), T_(thm), x, tape, ybar)
        ref = CO.rollout(mspec, thm, x_sub, ybar=yb_sub, dtype=np.float64)
        assert normwise(y[torch.tensor(idx, device="cuda")].cpu().numpy(), ref["y"]) < TOL_STATE
        assert l2rel(g.cpu().numpy(), ref["theta_bar"]) < TOL_GRAD


@pytest.mark.parametrize("name", list(mlp_model_variants()))
@pytest.mark.parametrize("B,force,nh", [(301, "1", "1"), (2500, None, None), (700, "1", "2")])
def test_mlp_family_rollout_vs_oracle(name, B, force, nh, T_, monkeypatch):
    """deep-MLP tcgen05 forward family (cc_mlp_kernels.cuh): outputs, final state and the tape it leaves for the FFMA
    reverse pass against the fp64 oracle; B = 301 (3 tiles, ragged) forced, B = 2500 as the default plan picks it"""
    import ctypes as C

    from chain_control_b200 import _lib, ops

    if force:
        monkeypatch.setenv("CC_B200_MLP", force)
    if nh:  # one / two threads per trajectory (cc_mlp_kernels.cuh / cc_mlp2_kernels.cuh); default: the plan's choice
        monkeypatch.setenv("CC_B200_MLP_NH", nh)
    spec = mlp_model_variants()[name]
    T = 60
    assert _lib.lib().cc_kernel_family_at(None, C.byref(conv(spec).to_c()), 0, B) == 5
    theta = O.init_params(spec, 1).astype(np.float32)
    u = _u(B, T, spec.input_dim)
    x0 = (0.1 * np.random.default_rng(3).normal(size=(B, spec.state_dim))).astype(np.float32)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    wide = max(spec.f.sizes[1:]) > 32  # the FFMA reverse pass does not fit 64-wide deep MLPs
    ref = CO.rollout(spec, theta, u, x0=x0, ybar=None if wide else ybar, dtype=np.float64)
    # trajectories whose outputs nearly cancel put the fp32 oracle itself near / above 1e-5 in the normwise metric for the
    # wide layers: the kernel has to stay within 3x of that floor (observed floors: profiles/r02_parity_floors.md)
    floor = normwise(CO.rollout(spec, theta, u, x0=x0, dtype=np.float32)["y"], ref["y"])
    y, xf, tape = ops.rollout_fwd(conv(spec), T_(theta), T_(u), T_(x0), want_tape=not wide, want_x_final=True)
    assert normwise(y.cpu().numpy(), ref["y"]) < max(TOL_STATE, 3 * floor), floor
    xs = O.rollout_fwd(spec, theta, u[:16], x0=x0[:16])[1]
    assert normwise(xf.cpu().numpy()[:16], xs[:, -1]) < TOL_STATE
    # the same call on the FFMA kernels: two independent implementations agree
    monkeypatch.setenv("CC_B200_MLP", "0")
    if max(spec.f.sizes[1:]) <= 64:  # (beyond 64 no other family exists)
        y2, _, _ = ops.rollout_fwd(conv(spec), T_(theta), T_(u), T_(x0))
        assert normwise(y2.cpu().numpy(), y.cpu().numpy()) < max(TOL_STATE, 3 * floor), floor
    if not wide:  # BPTT (FFMA reverse kernel) from the tape the tcgen05 forward wrote
        r32 = CO.rollout(spec, theta, u, x0=x0, ybar=ybar, dtype=np.float32)
        floor = l2rel(r32["theta_bar"], ref["theta_bar"])
        tb, _ = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar), T_(x0))
        assert l2rel(tb.cpu().numpy(), ref["theta_bar"]) < max(TOL_GRAD, 3 * floor), floor


@pytest.mark.parametrize("B", [3000, 8000])
@pytest.mark.parametrize("mname", ["compact_s50_d0", "euler_s30_d0_tanh"])
def test_default_plan_forward_reverse_pairs(B, mname, T_):
    """default environment (no family forced): the forward and reverse plans pick their families independently with different
    batch thresholds (latency / tcgen05 / FFMA / blocked linear); B = 3000 and 8000 sit in the ranges where the two passes of
    non-linear depth-0 models come from different families, which the forced-family fixtures never pair"""
    import ctypes as C

    from chain_control_b200 import _lib, ops

    for k in ("CC_B200_TC", "CC_B200_WPT_MAXB", "CC_B200_LIN", "CC_B200_MLP", "CC_B200_TCW"):
        assert k not in os.environ
    spec = tc_model_variants()[mname]
    T = 40
    theta = O.init_params(spec, 1, stable=True).astype(np.float32)
    u = _u(B, T, spec.input_dim)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    fams = tuple(_lib.lib().cc_kernel_family_at(None, C.byref(conv(spec).to_c()), g, B) for g in (0, 1))
    ref = CO.rollout(spec, theta, u, ybar=ybar, dtype=np.float64)
    y, _, tape = ops.rollout_fwd(conv(spec), T_(theta), T_(u), want_tape=True)
    tb, _ = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar))
    assert normwise(y.cpu().numpy(), ref["y"]) < TOL_STATE, fams
    assert l2rel(tb.cpu().numpy(), ref["theta_bar"]) < TOL_GRAD, fams


MLPB_GPU_CASES = {
    "s40_w48_d2": (lambda: O.make_node(40, 1, 1, f_width=48, f_depth=2), 301, 57),
    "s64_w64_d1_tanh_pre": (lambda: O.make_node(64, 2, 2, f_width=64, f_depth=1, f_act=O.ACT_TANH, f_act_final=O.ACT_TANH,
                                                readout_pre_step=True, u_transform=O.UT_ARCTAN), 257, 40),
    "s64_w64_d2_tanh": (lambda: O.make_node(64, 1, 1, f_width=64, f_depth=2, f_act_final=O.ACT_TANH), 257, 120),
    "s20_w33_d2_euler": (lambda: O.make_node(20, 1, 1, f_width=33, f_depth=2, integrator=O.INT_EE, f_act=O.ACT_TANH), 130, 33),
    "s20_w40_d1_noint": (lambda: O.make_node(20, 1, 3, f_width=40, f_depth=1, integrator=O.INT_NONE, f_use_bias=False,
                                             g_use_bias=False, f_act_final=O.ACT_TANH), 33, 21),
    "s33_w64_d2_one_step": (lambda: O.make_node(33, 4, 4, f_width=64, f_depth=2, readout_pre_step=True), 5, 1),
    # widths 17..32 take the family from 1024 trajectories on (padded to the 64-wide rounds)
    "s20_w25_d2_b1100": (lambda: O.make_node(20, 1, 1, f_width=25, f_depth=2, f_act_final=O.ACT_TANH), 1100, 12),
}


@pytest.mark.parametrize("flush", ["1", "2", "5", "3/nh4"])
@pytest.mark.parametrize("name", list(MLPB_GPU_CASES))
def test_mlpb_reverse_vs_oracle(name, flush, T_, monkeypatch):
    """deep-MLP tcgen05 reverse pass (family 6, cc_mlpb_kernels.cuh) against the fp64 C oracle: theta_bar of f and g"""
    import ctypes as C

    from chain_control_b200 import _lib, ops

    if flush.endswith("/nh4"):  # the four-threads-per-trajectory variant (16 columns per thread)
        flush = flush[:-4]
        monkeypatch.setenv("CC_B200_MLPB_NH", "4")
    monkeypatch.setenv("CC_B200_MLPB_FLUSH", flush)
    mk, B, T = MLPB_GPU_CASES[name]
    spec = mk()
    assert _lib.lib().cc_kernel_family_at(None, C.byref(conv(spec).to_c()), 1, B) == 6
    theta = O.init_params(spec, 1).astype(np.float32)
    u = _u(B, T, spec.input_dim)
    x0 = (0.1 * np.random.default_rng(3).normal(size=(B, spec.state_dim))).astype(np.float32)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    ref = CO.rollout(spec, theta, u, x0=x0, ybar=ybar, want_ubar=False, dtype=np.float64)
    y, _, tape = ops.rollout_fwd(conv(spec), T_(theta), T_(u), T_(x0), want_tape=True)
    tb, _ = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar), T_(x0))
    # (forward: same floor convention as test_mlp_family_rollout_vs_oracle -- the fp32 oracle itself sits near 1e-5 on wide layers)
    floor = normwise(CO.rollout(spec, theta, u, x0=x0, dtype=np.float32)["y"], ref["y"])
    assert normwise(y.cpu().numpy(), ref["y"]) < max(TOL_STATE, 3 * floor), floor
    assert l2rel(tb.cpu().numpy(), ref["theta_bar"]) < 2e-5  # (1e-6 observed; TOL_GRAD is 1e-4)
    tb2, _ = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape, T_(ybar), T_(x0))
    assert (tb2 == tb).all()  # deterministic: fixed-order reduction of the per-CTA sums


@pytest.mark.gpu
@pytest.mark.parametrize("name,K", [("s40_w48_d2", 8), ("s64_w64_d1_tanh_pre", 7), ("s64_w64_d2_tanh", 50), ("s20_w25_d2_b1100", 5)])
def test_mlpb_checkpointed_vs_oracle(name, K, T_):
    """family 6 with ckpt_every = K > 1 (north_star (3) for the deep-MLP family): checkpoints every K steps from the family-5
    forward kernel, segment replay + tensor-core reverse pass per segment; against the fp64 oracle and the unsegmented pass"""
    from chain_control_b200 import ops

    mk, B, T = MLPB_GPU_CASES[name]
    spec = mk()
    theta = O.init_params(spec, 1).astype(np.float32)
    u = _u(B, T, spec.input_dim)
    x0 = (0.1 * np.random.default_rng(3).normal(size=(B, spec.state_dim))).astype(np.float32)
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim)).astype(np.float32)
    ref = CO.rollout(spec, theta, u, x0=x0, ybar=ybar, want_ubar=False, dtype=np.float64)
    y1, _, tape1 = ops.rollout_fwd(conv(spec), T_(theta), T_(u), T_(x0), want_tape=True)
    tb1, _ = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tape1, T_(ybar), T_(x0))
    yk, _, tapek = ops.rollout_fwd(conv(spec), T_(theta), T_(u), T_(x0), want_tape=True, ckpt_every=K)
    assert tapek.numel() < tape1.numel() / min(K, T) * 1.6 + 4 * T + 64  # the tape really shrinks
    assert (yk == y1).all()
    tbk, _ = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tapek, T_(ybar), T_(x0), ckpt_every=K)
    assert l2rel(tbk.cpu().numpy(), ref["theta_bar"]) < 2e-5
    assert l2rel(tbk.cpu().numpy(), tb1.cpu().numpy()) < 5e-6
    tbk2, _ = ops.rollout_bwd(conv(spec), T_(theta), T_(u), tapek, T_(ybar), T_(x0), ckpt_every=K)
    assert (tbk2 == tbk).all()


@pytest.mark.gpu
def test_mlpb_long_horizon_checkpointed_runs(T_):
    """(64,64,2) ModelTrainer gradient at T = 4,000 x 4,736 trajectories with ckpt_every = 100: the tape is 40 checkpoints
    (48 MB) instead of 4,000 states (4.8 GB); finite, deterministic, and equal to the sum over two half batches"""
    import torch

    from chain_control_b200 import ops

    spec = O.make_node(64, 1, 1, f_width=64, f_depth=2, f_act_final=O.ACT_TANH)
    B, T, K = 4736, 4000, 100
    theta = T_(O.init_params(spec, 1).astype(np.float32))
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.rand((B, T, 1), device="cuda", generator=g) * 2 - 1
    ybar = torch.randn((B, T + 1, 1), device="cuda", generator=g) / (B * T)
    y, _, tape = ops.rollout_fwd(conv(spec), theta, u, want_tape=True, ckpt_every=K)
    assert tape.numel() * 4 < 60e6
    tb, _ = ops.rollout_bwd(conv(spec), theta, u, tape, ybar, ckpt_every=K)
    assert torch.isfinite(tb).all() and float(tb.abs().max()) > 0
    h = B // 2
    parts = []
    for sl in (slice(0, h), slice(h, B)):
        _, _, tp = ops.rollout_fwd(conv(spec), theta, u[sl].contiguous(), want_tape=True, ckpt_every=K)
        parts.append(ops.rollout_bwd(conv(spec), theta, u[sl].contiguous(), tp, ybar[sl].contiguous(), ckpt_every=K)[0])
    assert l2rel((parts[0] + parts[1]).cpu().numpy(), tb.cpu().numpy()) < 1e-5


def test_mlpb_full_size_gradient_vs_oracle(T_):
    """(64,64,2) at 65,536 trajectories (512 tiles on 148 SMs: every CTA slot busy, four waves): ybar is non-zero on 384
    scattered trajectories, so the fp64 oracle computes the same gradient on that sub-batch; plus additivity over the batch"""
    from chain_control_b200 import ops

    spec = O.make_node(64, 1, 1, f_width=64, f_depth=2, f_act_final=O.ACT_TANH)
    B, T = 65536, 48
    rng = np.random.default_rng(11)
    theta = O.init_params(spec, 1).astype(np.float32)
    u = rng.uniform(-1, 1, size=(B, T, 1)).astype(np.float32)
    idx = np.sort(rng.choice(B, size=384, replace=False))
    ybar = np.zeros((B, T + 1, 1), np.float32)
    ybar[idx] = rng.normal(size=(len(idx), T + 1, 1)).astype(np.float32)
    ref = CO.rollout(spec, theta, u[idx], ybar=ybar[idx], want_ubar=False, dtype=np.float64)
    th, ud, yb = T_(theta), T_(u), T_(ybar)
    y, _, tape = ops.rollout_fwd(conv(spec), th, ud, want_tape=True)
    tb, _ = ops.rollout_bwd(conv(spec), th, ud, tape, yb)
    floor = normwise(CO.rollout(spec, theta, u[idx], dtype=np.float32)["y"], ref["y"])
    assert normwise(y.cpu().numpy()[idx], ref["y"]) < max(TOL_STATE, 3 * floor), floor
    assert l2rel(tb.cpu().numpy(), ref["theta_bar"]) < TOL_GRAD
    # additivity: the gradient of the whole batch = the sum over two halves run as separate calls
    half = B // 2
    parts = []
    for sl in (slice(0, half), slice(half, B)):
        _, _, tp = ops.rollout_fwd(conv(spec), th, ud[sl].contiguous(), want_tape=True)
        parts.append(ops.rollout_bwd(conv(spec), th, ud[sl].contiguous(), tp, yb[sl].contiguous())[0])
    assert l2rel((parts[0] + parts[1]).cpu().numpy(), tb.cpu().numpy()) < 1e-5


def test_mlpb_autograd_model_trainer_step(T_):
    """the public differentiable entry (ops.batched_unroll, what ModelTrainer's loss calls) on a 64-wide deep model: loss
    gradient through family 5 forward + family 6 reverse against the fp64 oracle"""
    import torch

    from chain_control_b200 import ops

    spec = O.make_node(48, 1, 1, f_width=64, f_depth=2)
    B, T = 200, 30
    rng = np.random.default_rng(2)
    theta = O.init_params(spec, 1).astype(np.float32)
    u = rng.uniform(-1, 1, size=(B, T, 1)).astype(np.float32)
    ytgt = rng.normal(size=(B, T + 1, 1)).astype(np.float32)
    th = T_(theta).requires_grad_(True)
    y = ops.batched_unroll(conv(spec), th, T_(u))
    loss = torch.mean((y - T_(ytgt)) ** 2)
    loss.backward()
    y64, _ = O.rollout_fwd(spec, theta, u)
    ybar = 2.0 * (y64 - ytgt) / y64.size
    tb64, _ = O.rollout_bwd(spec, theta, u, ybar)
    lv = float(loss.detach())
    assert abs(lv - float(np.mean((y64 - ytgt) ** 2))) < 1e-5 * max(1.0, lv)
    assert l2rel(th.grad.cpu().numpy(), tb64) < TOL_GRAD
