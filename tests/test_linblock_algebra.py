"""CPU: the algebra behind the blocked linear kernel family (chain_control_b200/csrc/cc_lin.h) -- m integrator steps of a
linear depth-0 model collapsed into one matrix product, Markov-parameter sums for the in-block outputs, the transposed
recursion for the reverse pass and the gradient of a trainable model from Q = sum Lam (x) X -- stated in NumPy
(tests/linblock_ref.py) and checked against the fp64 oracle to machine precision."""
import numpy as np
import pytest

import linblock_ref as LB
from oracle import np_oracle as O


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


CASES = [
    (50, 1, 1, True, O.INT_RK4, True, O.UT_ARCTAN),    # notebook-3/4 compact model
    (20, 1, 1, False, O.INT_RK4, True, O.UT_KTANH),    # full model (post-step readout)
    (7, 2, 3, True, O.INT_EE, False, O.UT_IDENTITY),   # explicit Euler, no bias, several inputs / outputs
    (6, 2, 2, False, O.INT_NONE, True, O.UT_ARCTAN),   # "no-integrate"
]


@pytest.mark.parametrize("S,I,Od,pre,integ,bias,ut", CASES)
@pytest.mark.parametrize("m", [1, 4, 10])
def test_blocked_open_loop_equals_oracle(S, I, Od, pre, integ, bias, ut, m):
    rng = np.random.default_rng(0)
    spec = O.make_node(S, I, Od, f_depth=0, readout_pre_step=pre, integrator=integ, f_use_bias=bias, g_use_bias=bias,
                       u_transform=ut, u_scale=6.0)
    th = O.init_params(spec, 1, stable=(integ != O.INT_NONE))
    if integ == O.INT_NONE:
        th = th * 0.5
    B, T = 5, 37
    u = rng.uniform(-1, 1, (B, T, I))
    x0 = 0.1 * rng.normal(size=(B, S))
    ybar = rng.normal(size=(B, T + 1, Od))
    y, _ = O.rollout_fwd(spec, th, u, x0=x0)
    tb, ub = O.rollout_bwd(spec, th, u, ybar, x0=x0)
    y2, _ = LB.rollout_fwd(spec, th, u, m, x0)
    tb2, ub2, _ = LB.rollout_bwd(spec, th, u, ybar, m, x0)
    assert rel(y2, y) < 1e-12
    assert rel(tb2, tb) < 1e-11
    assert rel(ub2, ub) < 1e-12


@pytest.mark.parametrize("S,pre,integ", [(50, True, O.INT_RK4), (20, False, O.INT_RK4), (9, True, O.INT_EE)])
@pytest.mark.parametrize("m", [1, 5, 13])
def test_blocked_closed_loop_equals_oracle(S, pre, integ, m):
    rng = np.random.default_rng(1)
    spec = O.make_node(S, 1, 1, f_depth=0, readout_pre_step=pre, integrator=integ, u_transform=O.UT_ARCTAN)
    th = O.init_params(spec, 1, stable=True)
    ctrls = [O.make_node(5, 2, 1, f_depth=0, readout_pre_step=True, out_scale=0.5),
             O.make_node(6, 2, 1, f_depth=0, g_feedthrough_u=True, f_time_variant=True, g_time_variant=True)]
    B, Tr = 5, 41
    for cs in ctrls:
        thc = O.init_params(cs, 2)
        refs = 3 * rng.uniform(-1, 1, (B, Tr, 1))
        noise = 0.02 * rng.normal(size=(B, Tr, 1))
        yb = rng.normal(size=(B, Tr, 1))
        yh, us = O.closedloop_fwd(cs, spec, thc, th, refs, noise)
        g = O.closedloop_bwd(cs, spec, thc, th, refs, yb, noise)
        yh2, us2, tape = LB.closedloop_fwd(cs, spec, thc, th, refs, m, noise)
        g2 = LB.closedloop_bwd(cs, spec, thc, th, refs, yb, m, tape)
        assert rel(yh2, yh) < 1e-12 and rel(us2, us) < 1e-12
        assert rel(g2, g) < 1e-11
