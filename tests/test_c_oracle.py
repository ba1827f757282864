"""The C port of the oracle against the numpy restatement (fp64 tight; fp32 at the fp32 noise floor)."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import np_oracle as O
from ccspecs import controller_variants, model_variants


def _u(B, T, I, seed=0):
    return np.random.default_rng(seed).uniform(-1, 1, size=(B, T, I))


@pytest.mark.parametrize("name", list(model_variants()))
def test_c_rollout_vs_numpy_f64(name):
    spec = model_variants()[name]
    B, T = 19, 23  # not a multiple of the 16-lane block
    theta = O.init_params(spec, 1)
    u = _u(B, T, spec.input_dim)
    x0 = 0.1 * np.random.default_rng(2).normal(size=(B, spec.state_dim))
    ybar = np.random.default_rng(5).normal(size=(B, T + 1, spec.output_dim))
    y, _ = O.rollout_fwd(spec, theta, u, x0=x0)
    tb, ub = O.rollout_bwd(spec, theta, u, ybar, x0=x0)
    r = CO.rollout(spec, theta, u, x0=x0, ybar=ybar, want_ubar=True, dtype=np.float64, nthreads=3)
    np.testing.assert_allclose(r["y"], y, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(r["theta_bar"], tb, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(r["ubar"], ub, rtol=1e-9, atol=1e-11)
    r0 = CO.rollout(spec, theta, u, x0=x0, dtype=np.float64)
    np.testing.assert_array_equal(r0["y"], r["y"])


@pytest.mark.parametrize("cname", list(controller_variants()))
@pytest.mark.parametrize("mname", ["compact_s50_d0", "full_s5_w10_d2", "full_s1_tv"])
def test_c_closedloop_vs_numpy_f64(cname, mname):
    mspec, cspec = model_variants()[mname], controller_variants()[cname]
    B, Tr = 18, 14
    theta_c, theta_m = O.init_params(cspec, 2), O.init_params(mspec, 1)
    refs = _u(B, Tr, 1, 3) * 3
    noise = 0.02 * np.random.default_rng(9).normal(size=(B, Tr, 1))
    ybar = np.random.default_rng(6).normal(size=(B, Tr, 1))
    yhat, us = O.closedloop_fwd(cspec, mspec, theta_c, theta_m, refs, noise)
    tb, tmb = O.closedloop_bwd(cspec, mspec, theta_c, theta_m, refs, ybar, noise, with_model_grad=True)
    r = CO.closedloop(cspec, mspec, theta_c, theta_m, refs, noise, ybar=ybar, with_model_grad=True,
                      dtype=np.float64, nthreads=2)
    np.testing.assert_allclose(r["yhat"], yhat, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(r["u"], us, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(r["theta_c_bar"], tb, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(r["theta_m_bar"], tmb, rtol=1e-9, atol=1e-11)


def test_c_mse_value_and_grad():
    spec = model_variants()["compact_s50_d0"]
    B, T = 8, 50
    theta = O.init_params(spec, 1)
    u = _u(B, T, 1)
    targets = np.random.default_rng(4).normal(size=(B, T + 1, 1))
    y, _ = O.rollout_fwd(spec, theta, u)
    tb, _ = O.rollout_bwd(spec, theta, u, O.mse_cotangent(targets, y))
    r = CO.rollout(spec, theta, u, targets=targets, dtype=np.float64)
    assert abs(r["loss"] - O.mse(targets, y)) < 1e-12
    np.testing.assert_allclose(r["theta_bar"], tb, rtol=1e-9, atol=1e-12)
    # closed loop: mse(refs, yhat)
    cspec = controller_variants()["compact_s5_d0"]
    theta_c = O.init_params(cspec, 2)
    refs = _u(B, T, 1, 3)
    yhat, _ = O.closedloop_fwd(cspec, spec, theta_c, theta, refs)
    tcb = O.closedloop_bwd(cspec, spec, theta_c, theta, refs, O.mse_cotangent(refs, yhat))
    r = CO.closedloop(cspec, spec, theta_c, theta, refs, mse_loss=True, dtype=np.float64)
    assert abs(r["loss"] - O.mse(refs, yhat)) < 1e-12
    np.testing.assert_allclose(r["theta_c_bar"], tcb, rtol=1e-9, atol=1e-12)


def test_c_f32_close_to_f64():
    spec = model_variants()["compact_s50_d0"]
    theta = O.init_params(spec, 1)
    u = _u(8, 1000, 1)
    y64 = CO.rollout(spec, theta, u, dtype=np.float64)["y"]
    y32 = CO.rollout(spec, theta, u, dtype=np.float32)["y"]
    err = np.abs(y32 - y64).max(axis=(1, 2)) / np.abs(y64).max(axis=(1, 2))
    assert err.max() < 1e-5, err
