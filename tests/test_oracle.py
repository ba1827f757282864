"""CPU tests of the oracle itself (no GPU): the numpy restatement against an independent torch
autograd statement of the reference forward, and against central finite differences."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
import ccref_torch as TR
from ccspecs import controller_variants, model_variants

torch.set_num_threads(2)


def _u(B, T, I, seed=0):
    return np.random.default_rng(seed).uniform(-1, 1, size=(B, T, I))


@pytest.mark.parametrize("name", list(model_variants()))
def test_rollout_fwd_bwd_vs_torch(name):
    spec = model_variants()[name]
    B, T = 3, 17
    theta = O.init_params(spec, 1)
    u = _u(B, T, spec.input_dim)
    y, xs = O.rollout_fwd(spec, theta, u)
    assert y.shape == (B, T + 1, spec.output_dim)

    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    ut = torch.tensor(u, dtype=torch.float64, requires_grad=True)
    yt = TR.rollout(spec, th, ut)
    np.testing.assert_allclose(y, yt.detach().numpy(), rtol=1e-12, atol=1e-13)

    ybar = np.random.default_rng(5).normal(size=y.shape)
    (yt * torch.tensor(ybar)).sum().backward()
    tb, ub = O.rollout_bwd(spec, theta, u, ybar)
    np.testing.assert_allclose(tb, th.grad.numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(ub, ut.grad.numpy(), rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("cname", list(controller_variants()))
@pytest.mark.parametrize("mname", ["compact_s50_d0", "full_s5_w10_d2", "full_s1_tv"])
def test_closedloop_fwd_bwd_vs_torch(cname, mname):
    mspec = model_variants()[mname]
    cspec = controller_variants()[cname]
    B, Tr = 3, 13
    theta_c = O.init_params(cspec, 2)
    theta_m = O.init_params(mspec, 1)
    refs = _u(B, Tr, 1, seed=3) * 3
    noise = 0.02 * np.random.default_rng(9).normal(size=(B, Tr, 1))
    yhat, us = O.closedloop_fwd(cspec, mspec, theta_c, theta_m, refs, noise)

    tc = torch.tensor(theta_c, dtype=torch.float64, requires_grad=True)
    tm = torch.tensor(theta_m, dtype=torch.float64, requires_grad=True)
    yt = TR.closedloop(cspec, mspec, tc, tm, torch.tensor(refs), torch.tensor(noise))
    np.testing.assert_allclose(yhat, yt.detach().numpy(), rtol=1e-12, atol=1e-13)

    ybar = np.random.default_rng(6).normal(size=yhat.shape)
    (yt * torch.tensor(ybar)).sum().backward()
    tb, tmb = O.closedloop_bwd(cspec, mspec, theta_c, theta_m, refs, ybar, noise, with_model_grad=True)
    np.testing.assert_allclose(tb, tc.grad.numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(tmb, tm.grad.numpy(), rtol=1e-9, atol=1e-11)


def test_finite_differences_mse_loss():
    """d mse / d theta by central differences on a handful of parameters (fp64)."""
    spec = model_variants()["full_s12_w25_d1_tanh"]
    B, T = 2, 25
    theta = O.init_params(spec, 4)
    u = _u(B, T, spec.input_dim, 7)
    target = np.random.default_rng(8).normal(size=(B, T + 1, spec.output_dim))
    y, _ = O.rollout_fwd(spec, theta, u)
    tb, _ = O.rollout_bwd(spec, theta, u, O.mse_cotangent(target, y))
    rng = np.random.default_rng(0)
    for idx in rng.choice(theta.size, 12, replace=False):
        e = np.zeros_like(theta)
        e[idx] = 1e-6
        lp = O.mse(target, O.rollout_fwd(spec, theta + e, u)[0])
        lm = O.mse(target, O.rollout_fwd(spec, theta - e, u)[0])
        fd = (lp - lm) / 2e-6
        assert abs(fd - tb[idx]) <= 1e-6 * max(1.0, abs(fd)), (idx, fd, tb[idx])


def test_unroll_equals_successive_steps():
    """AbstractModel.unroll == T successive step() calls (abstract.py:54-70) incl. include_y0."""
    spec = model_variants()["full_s5_w10_d2"]
    theta = O.init_params(spec, 1)
    u = _u(2, 9, 1)
    y, xs = O.rollout_fwd(spec, theta, u)
    fl, gl = O.split_params(spec, theta)
    x = np.zeros((2, spec.state_dim))
    for k in range(9):
        x, out, _ = O.node_step(spec, fl, gl, x, np.float64(0), u[:, k])
        np.testing.assert_array_equal(out, y[:, k + 1])
        np.testing.assert_array_equal(x, xs[:, k + 1])
    np.testing.assert_array_equal(y[:, 0], O.node_y0(spec, gl, 2, np.float64)[0])


def test_compact_readout_is_pre_step():
    """compact model emits [g(x0), g(x0), g(x1), ...] (neural_ode_model_compact_example.py:103,121)."""
    spec = model_variants()["compact_s50_d0"]
    theta = O.init_params(spec, 1)
    y, xs = O.rollout_fwd(spec, theta, _u(2, 6, 1))
    np.testing.assert_array_equal(y[:, 0], y[:, 1])
    fl, gl = O.split_params(spec, theta)
    W, b = gl[0]
    np.testing.assert_allclose(y[:, 3], xs[:, 2] @ W.T + b, rtol=1e-13)


def test_zero_controller_closed_loop_equals_open_loop_zero_input():
    mspec = model_variants()["full_s5_w10_d2"]
    cspec = controller_variants()["compact_s5_d0"]
    theta_m = O.init_params(mspec, 1)
    theta_c = np.zeros(cspec.n_params())
    refs = _u(2, 11, 1)
    yhat, us = O.closedloop_fwd(cspec, mspec, theta_c, theta_m, refs)
    assert np.all(us == 0)
    y, _ = O.rollout_fwd(mspec, theta_m, np.zeros((2, 11, 1)))
    np.testing.assert_allclose(yhat, y[:, 1:], rtol=1e-13, atol=1e-15)


def test_fp32_noise_floor():
    """fp32 restatement vs fp64 truth at T=1000: normwise 1e-5 on outputs is attainable."""
    spec = model_variants()["compact_s50_d0"]
    theta = O.init_params(spec, 1)
    u = _u(4, 1000, 1)
    y64, _ = O.rollout_fwd(spec, theta, u)
    y32, _ = O.rollout_fwd(spec, theta, u, dtype=np.float32)
    err = np.abs(y32 - y64).max(axis=(1, 2)) / np.abs(y64).max(axis=(1, 2))
    assert err.max() < 1e-5, err
