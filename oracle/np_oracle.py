"""NumPy restatement of chain_control's neural-ODE unroll hot path (TEST INFRASTRUCTURE).

This module is the *oracle*: a CPU restatement of the reference algorithm, used only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs as the checker.  The product path (``chain_control_b200``) never imports it.

PARITY PINNED BY REFERENCE EXECUTION (round 2): jax / equinox==0.9.0 / optax / imt-tree-utils (reference
``setup.py:40-59``) are not installable here and the reference's own tests pin no numeric value on this path
(``cc/train/test_trainer.py:25-147`` only checks that a step runs), so the golden vectors are GENERATED: the reference's
own source files (cc/core/abstract.py, module_utils.py, cc/examples/*.py, nn_lib/*, cc/train/step_fn.py,
cc/utils/utils.py) are imported from /root/reference and executed unmodified on a NumPy stand-in of that third-party
layer (``tests/golden/refshim.py``; gradients = complex-step derivatives of the reference's ``loss_fn_model`` /
``loss_fn_controller``), by the committed script ``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``.
``tests/test_golden.py`` requires this module and the C oracle (fp64) to reproduce those outputs, losses and
gradients to 1e-9 (11 fixtures: compact/full models, time-variant, feed-through, Euler / no-integrate, two output leaves,
observation noise, two-model mean, l1/l2 regularisers).  What the stand-in cannot pin: jax's threefry streams (weight init,
the noise VALUES; the noise semantics are pinned with recorded draws) and fp32 rounding order inside XLA.
Cross-checks that predate the fixtures: torch autograd and central finite differences (``tests/test_oracle.py``).

Reference lines restated (all relative to /root/reference):
  * open-loop unroll      cc/core/abstract.py:54-70,  cc/core/module_utils.py:8-17
  * closed-loop unroll    cc/core/abstract.py:97-127
  * model step (full)     cc/examples/neural_ode_model.py:106-149, y0 :160-175
  * model step (compact)  cc/examples/neural_ode_model_compact_example.py:94-110, y0 :121-122
  * controller (full)     cc/examples/neural_ode_controller.py:108-149
  * controller (compact)  cc/examples/neural_ode_controller_compact_example.py:106-129
  * integrators           cc/examples/nn_lib/integrate.py:1-28
  * MLP                   cc/examples/nn_lib/mlp_network.py:5-35  (eqx.nn.Linear: y = W x + b, W[out,in])
  * flat parameter order  cc/core/module_utils.py:20-24
  * default loss          cc/utils/utils.py:53-55 (mse), cc/train/step_fn.py:133-146, 249-283

Conventions: everything is batched over a leading axis B (the reference's ``eqx.filter_vmap``,
cc/train/step_fn.py:136,257).  ``dtype`` selects fp64 (truth) or fp32 (reference-equivalent).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# activation / transform / integrator enums (values shared with include/cc_b200.h)
ACT_IDENTITY, ACT_RELU, ACT_TANH = 0, 1, 2
UT_IDENTITY, UT_ARCTAN, UT_KTANH = 0, 1, 2
INT_RK4, INT_EE, INT_NONE = 0, 1, 2


@dataclass
class MlpSpec:
    """sizes = [in] + depth*[width] + [out]  (mlp_network.py:18)."""

    sizes: List[int]
    act: int = ACT_RELU
    act_final: int = ACT_IDENTITY
    use_bias: bool = True

    @property
    def n_layers(self) -> int:
        return len(self.sizes) - 1

    def n_params(self) -> int:
        return sum(o * i + (o if self.use_bias else 0) for i, o in zip(self.sizes[:-1], self.sizes[1:]))


@dataclass
class NodeSpec:
    """One neural-ODE module (model or controller)."""

    state_dim: int
    input_dim: int
    output_dim: int
    f: MlpSpec
    g: MlpSpec
    dt: float = 0.01
    integrator: int = INT_RK4
    f_time_variant: bool = False
    g_time_variant: bool = False
    g_feedthrough_u: bool = False
    readout_pre_step: bool = False  # compact variants read out g(x) BEFORE the update
    u_transform: int = UT_IDENTITY
    u_scale: float = 1.0  # k in k*tanh(u/k)
    out_scale: float = 1.0  # compact controller: sigmoid(alpha)=0.5 with a zero secondary

    def n_params(self) -> int:
        return self.f.n_params() + self.g.n_params()

    @property
    def has_time(self) -> bool:
        return self.f_time_variant or self.g_time_variant


def make_node(
    state_dim,
    input_dim,
    output_dim,
    dt=0.01,
    f_width=10,
    f_depth=2,
    g_width=10,
    g_depth=0,
    f_act=ACT_RELU,
    f_act_final=ACT_IDENTITY,
    g_act=ACT_RELU,
    g_act_final=ACT_IDENTITY,
    f_use_bias=True,
    g_use_bias=True,
    integrator=INT_RK4,
    f_time_variant=False,
    g_time_variant=False,
    g_feedthrough_u=False,
    readout_pre_step=False,
    u_transform=UT_IDENTITY,
    u_scale=1.0,
    out_scale=1.0,
) -> NodeSpec:
    """Mirror of the size bookkeeping in neural_ode_model.py:45-59 / neural_ode_controller.py:45-60."""
    f_in = state_dim + input_dim + (1 if f_time_variant else 0)
    g_in = state_dim + (1 if g_time_variant else 0) + (input_dim if g_feedthrough_u else 0)
    f = MlpSpec([f_in] + f_depth * [f_width] + [state_dim], f_act, f_act_final, f_use_bias)
    g = MlpSpec([g_in] + g_depth * [g_width] + [output_dim], g_act, g_act_final, g_use_bias)
    return NodeSpec(
        state_dim, input_dim, output_dim, f, g, dt, integrator, f_time_variant, g_time_variant,
        g_feedthrough_u, readout_pre_step, u_transform, u_scale, out_scale,
    )


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
def init_params(spec: NodeSpec, seed: int, stable: bool = False) -> np.ndarray:
    """eqx.nn.Linear-style init U(-1/sqrt(in), 1/sqrt(in)) for weight and bias, flat fp64 vector in
    flatten_module order (f layers then g layers; weight then bias).  JAX's threefry streams are not
    reproducible here, so the draw is distribution-equivalent only (numpy default_rng).
    ``stable=True`` subtracts I from the state->state block of f's first layer when f has depth 0
    (contractive dynamics for long horizons)."""
    rng = np.random.default_rng(seed)
    out = []
    for mlp in (spec.f, spec.g):
        for li, (i, o) in enumerate(zip(mlp.sizes[:-1], mlp.sizes[1:])):
            lim = 1.0 / np.sqrt(i)
            W = rng.uniform(-lim, lim, size=(o, i))
            if stable and mlp is spec.f and mlp.n_layers == 1:
                W[:, : spec.state_dim] -= np.eye(spec.state_dim)
            out.append(W.ravel())
            if mlp.use_bias:
                out.append(rng.uniform(-lim, lim, size=(o,)))
    return np.concatenate(out)


def unflatten(mlp: MlpSpec, theta: np.ndarray, off: int):
    """-> list of (W[out,in], b[out] or None), new offset."""
    layers = []
    for i, o in zip(mlp.sizes[:-1], mlp.sizes[1:]):
        W = theta[off : off + o * i].reshape(o, i)
        off += o * i
        b = None
        if mlp.use_bias:
            b = theta[off : off + o]
            off += o
        layers.append((W, b))
    return layers, off


def split_params(spec: NodeSpec, theta: np.ndarray):
    f, off = unflatten(spec.f, theta, 0)
    g, off = unflatten(spec.g, theta, off)
    assert off == theta.size, (off, theta.size)
    return f, g


# ----------------------------------------------------------------------------------------------
# elementwise pieces
# ----------------------------------------------------------------------------------------------
def _act(kind, x):
    if kind == ACT_IDENTITY:
        return x
    if kind == ACT_RELU:
        return np.maximum(x, 0)
    if kind == ACT_TANH:
        return np.tanh(x)
    raise ValueError(kind)


def _act_grad_from_out(kind, y):
    """derivative expressed through the activation OUTPUT (relu: y>0, i.e. 0 at exactly 0 like JAX)."""
    if kind == ACT_IDENTITY:
        return np.ones_like(y)
    if kind == ACT_RELU:
        return (y > 0).astype(y.dtype)
    if kind == ACT_TANH:
        return 1 - y * y
    raise ValueError(kind)


def _ut(kind, scale, u):
    if kind == UT_IDENTITY:
        return u
    if kind == UT_ARCTAN:
        return np.arctan(u)
    if kind == UT_KTANH:
        s = u.dtype.type(scale)
        return s * np.tanh(u / s)
    raise ValueError(kind)


def _ut_grad(kind, scale, u):
    if kind == UT_IDENTITY:
        return np.ones_like(u)
    if kind == UT_ARCTAN:
        return 1 / (1 + u * u)
    if kind == UT_KTANH:
        s = u.dtype.type(scale)
        th = np.tanh(u / s)
        return 1 - th * th
    raise ValueError(kind)


def mlp_fwd(mlp: MlpSpec, layers, a):
    """a[B,in] -> out[B,out]; cache = list of layer inputs and outputs (for the reverse pass)."""
    ins, outs = [], []
    L = len(layers)
    for li, (W, b) in enumerate(layers):
        ins.append(a)
        pre = a @ W.T
        if b is not None:
            pre = pre + b
        a = _act(mlp.act if li < L - 1 else mlp.act_final, pre)
        outs.append(a)
    return a, (ins, outs)


def mlp_bwd(mlp: MlpSpec, layers, cache, out_bar, grads):
    """VJP of mlp_fwd.  grads: list of [Wbar, bbar] accumulated in place.  returns input cotangent."""
    ins, outs = cache
    L = len(layers)
    d = out_bar
    for li in range(L - 1, -1, -1):
        W, b = layers[li]
        d = d * _act_grad_from_out(mlp.act if li < L - 1 else mlp.act_final, outs[li])
        if grads is not None:
            grads[li][0] += d.T @ ins[li]
            if b is not None:
                grads[li][1] += d.sum(0)
        d = d @ W
    return d


# ----------------------------------------------------------------------------------------------
# one node step (model or controller) and its VJP
# ----------------------------------------------------------------------------------------------
def _f_in(spec, z, t, ut):
    parts = [z]
    if spec.f_time_variant:
        parts.append(np.broadcast_to(t, (z.shape[0], 1)).astype(z.dtype))
    parts.append(ut)
    return np.concatenate(parts, 1)


def _g_in(spec, x, t, v):
    parts = [x]
    if spec.g_time_variant:
        parts.append(np.broadcast_to(t, (x.shape[0], 1)).astype(x.dtype))
    if spec.g_feedthrough_u:
        parts.append(v)
    return np.concatenate(parts, 1)


def node_step(spec: NodeSpec, fl, gl, x, t, v):
    """x[B,S], t scalar (dtype), v[B,I] -> x_next, out[B,O], cache.
    neural_ode_model.py:106-149 / compact :94-110; integrate.py:6-13 (operation order kept)."""
    dt_ = x.dtype.type
    h = dt_(np.float32(spec.dt))  # the reference's fp32 path sees f32(control_timestep)
    ut = _ut(spec.u_transform, spec.u_scale, v)
    c = {"x": x, "t": t, "v": v, "stages": []}

    def rhs(tt, z):
        k, cache = mlp_fwd(spec.f, fl, _f_in(spec, z, tt, ut))
        c["stages"].append(cache)
        return k

    if spec.integrator == INT_RK4:
        k1 = rhs(t, x)
        k2 = rhs(t + h / dt_(2), x + h * k1 / dt_(2))
        k3 = rhs(t + h / dt_(2), x + h * k2 / dt_(2))
        k4 = rhs(t + h, x + h * k3)
        dx = dt_(1 / 6) * (k1 + dt_(2) * k2 + dt_(2) * k3 + k4)
        xn = x + h * dx
    elif spec.integrator == INT_EE:
        xn = x + h * rhs(t, x)
    elif spec.integrator == INT_NONE:
        xn = rhs(t, x)
    else:
        raise ValueError(spec.integrator)

    xr = x if spec.readout_pre_step else xn
    out, gc = mlp_fwd(spec.g, gl, _g_in(spec, xr, t, v))
    c["g"] = gc
    out = dt_(spec.out_scale) * out
    return xn, out, c


def node_step_bwd(spec: NodeSpec, fl, gl, c, out_bar, lam_next, fgrads, ggrads):
    """VJP of node_step: (out_bar[B,O], lam_next[B,S]) -> lam[B,S], v_bar[B,I].  SURVEY §9.4."""
    x, v = c["x"], c["v"]
    dt_ = x.dtype.type
    h = dt_(np.float32(spec.dt))
    S, I = spec.state_dim, spec.input_dim
    v_bar = np.zeros_like(v)
    ut_bar = np.zeros_like(v)

    # readout
    gin_bar = mlp_bwd(spec.g, gl, c["g"], dt_(spec.out_scale) * out_bar, ggrads)
    xr_bar = gin_bar[:, :S]
    if spec.g_feedthrough_u:
        o = S + (1 if spec.g_time_variant else 0)
        v_bar = v_bar + gin_bar[:, o : o + I]
    lam_n = lam_next if spec.readout_pre_step else lam_next + xr_bar

    def rhs_bwd(i, kbar):
        zin_bar = mlp_bwd(spec.f, fl, c["stages"][i], kbar, fgrads)
        nonlocal ut_bar
        ut_bar = ut_bar + zin_bar[:, -I:] if I > 0 else ut_bar
        return zin_bar[:, :S]

    if spec.integrator == INT_RK4:
        k4b = (h / dt_(6)) * lam_n
        z4b = rhs_bwd(3, k4b)
        k3b = (h / dt_(3)) * lam_n + h * z4b
        z3b = rhs_bwd(2, k3b)
        k2b = (h / dt_(3)) * lam_n + (h / dt_(2)) * z3b
        z2b = rhs_bwd(1, k2b)
        k1b = (h / dt_(6)) * lam_n + (h / dt_(2)) * z2b
        z1b = rhs_bwd(0, k1b)
        lam = lam_n + z1b + z2b + z3b + z4b
    elif spec.integrator == INT_EE:
        lam = lam_n + rhs_bwd(0, h * lam_n)
    else:
        lam = rhs_bwd(0, lam_n)

    if spec.readout_pre_step:
        lam = lam + xr_bar
    v_bar = v_bar + ut_bar * _ut_grad(spec.u_transform, spec.u_scale, v)
    return lam, v_bar


def node_y0(spec: NodeSpec, gl, B, dtype, x0=None):
    """y0(): g([x0;(0)]) (neural_ode_model.py:160-175).  A model with feed-through (LinearModel, y = C x + D u,
    cc/examples/linear_model.py:66-82) has y0 = C x0: the input columns see zeros."""
    x = np.zeros((B, spec.state_dim), dtype) if x0 is None else x0
    v0 = np.zeros((x.shape[0], spec.input_dim), dtype) if spec.g_feedthrough_u else None
    out, gc = mlp_fwd(spec.g, gl, _g_in(spec, x, dtype(0), v0))
    return dtype(spec.out_scale) * out, gc


def _zero_grads(layers):
    return [[np.zeros_like(W), None if b is None else np.zeros_like(b)] for W, b in layers]


def _flatten_grads(fg, gg):
    out = []
    for grads in (fg, gg):
        for Wb, bb in grads:
            out.append(Wb.ravel())
            if bb is not None:
                out.append(bb)
    return np.concatenate(out)


# ----------------------------------------------------------------------------------------------
# K1/K2: open-loop unroll (AbstractModel.unroll, abstract.py:54-70) and its VJP
# ----------------------------------------------------------------------------------------------
def rollout_fwd(spec: NodeSpec, theta, u, dtype=np.float64, x0=None, return_tape=False):
    """u[B,T,I] -> y[B,T+1,O] (include_y0=True).  Optionally the per-step caches for the VJP."""
    dtype = np.dtype(dtype).type
    theta = np.asarray(theta, dtype)
    u = np.asarray(u, dtype)
    B, T, _ = u.shape
    fl, gl = split_params(spec, theta)
    x = np.zeros((B, spec.state_dim), dtype) if x0 is None else np.asarray(x0, dtype)
    t = dtype(0)
    y = np.empty((B, T + 1, spec.output_dim), dtype)
    y[:, 0], g0 = node_y0(spec, gl, B, dtype, x)
    tape = []
    xs = np.empty((B, T + 1, spec.state_dim), dtype)
    xs[:, 0] = x
    for k in range(T):
        x, out, c = node_step(spec, fl, gl, x, t, u[:, k])
        y[:, k + 1] = out
        xs[:, k + 1] = x
        if return_tape:
            tape.append(c)
        if spec.has_time:
            t = t + dtype(np.float32(spec.dt))
    if return_tape:
        return y, xs, (tape, g0)
    return y, xs


def rollout_bwd(spec: NodeSpec, theta, u, ybar, dtype=np.float64, x0=None):
    """VJP of rollout_fwd w.r.t. theta (flat, flatten_module order) and u: ybar[B,T+1,O] -> theta_bar[P], ubar[B,T,I]."""
    dtype = np.dtype(dtype).type
    theta = np.asarray(theta, dtype)
    u = np.asarray(u, dtype)
    ybar = np.asarray(ybar, dtype)
    B, T, _ = u.shape
    fl, gl = split_params(spec, theta)
    y, xs, (tape, g0) = rollout_fwd(spec, theta, u, dtype, x0, return_tape=True)
    fg, gg = _zero_grads(fl), _zero_grads(gl)
    lam = np.zeros((B, spec.state_dim), dtype)
    ubar = np.zeros_like(u)
    for k in range(T - 1, -1, -1):
        lam, ubar[:, k] = node_step_bwd(spec, fl, gl, tape[k], ybar[:, k + 1], lam, fg, gg)
    # y0 = g(x0): only g's parameters see it (x0 is not trained, neural_ode_model.py:151-158)
    mlp_bwd(spec.g, gl, g0, dtype(spec.out_scale) * ybar[:, 0], gg)
    return _flatten_grads(fg, gg), ubar


# ----------------------------------------------------------------------------------------------
# K3/K4: closed-loop unroll (AbstractController.unroll, abstract.py:97-127) and its VJP
# ----------------------------------------------------------------------------------------------
def closedloop_fwd(cspec: NodeSpec, mspec: NodeSpec, theta_c, theta_m, refs, noise=None,
                   dtype=np.float64, return_tape=False):
    """refs[B,Tr,R] (+ optional additive noise[B,Tr,O], already scaled) -> yhat[B,Tr,O], u[B,Tr,I_m].
    Controller input is merge_x_y(ref, y) = [ref ; obs] (step_fn.py:187-192)."""
    dtype = np.dtype(dtype).type
    theta_c = np.asarray(theta_c, dtype)
    theta_m = np.asarray(theta_m, dtype)
    refs = np.asarray(refs, dtype)
    B, Tr, R = refs.shape
    assert cspec.input_dim == R + mspec.output_dim and cspec.output_dim == mspec.input_dim
    cfl, cgl = split_params(cspec, theta_c)
    mfl, mgl = split_params(mspec, theta_m)
    xc = np.zeros((B, cspec.state_dim), dtype)
    xm = np.zeros((B, mspec.state_dim), dtype)
    tc = dtype(0)
    tm = dtype(0)
    y, _ = node_y0(mspec, mgl, B, dtype)
    yhat = np.empty((B, Tr, mspec.output_dim), dtype)
    us = np.empty((B, Tr, mspec.input_dim), dtype)
    tape = []
    for k in range(Tr):
        v = np.concatenate([refs[:, k], y], 1)
        xc, a, cc_ = node_step(cspec, cfl, cgl, xc, tc, v)
        xm, y, mc_ = node_step(mspec, mfl, mgl, xm, tm, a)
        if noise is not None:
            y = y + np.asarray(noise[:, k], dtype)
        yhat[:, k] = y
        us[:, k] = a
        if return_tape:
            tape.append((cc_, mc_))
        if cspec.has_time:
            tc = tc + dtype(np.float32(cspec.dt))
        if mspec.has_time:
            tm = tm + dtype(np.float32(mspec.dt))
    if return_tape:
        return yhat, us, tape
    return yhat, us


def closedloop_bwd(cspec: NodeSpec, mspec: NodeSpec, theta_c, theta_m, refs, ybar, noise=None,
                   dtype=np.float64, with_model_grad=False):
    """VJP of closedloop_fwd w.r.t. the controller parameters (model frozen, step_fn.py:249-252).
    ``with_model_grad`` additionally returns the model-parameter cotangent (validation only)."""
    dtype = np.dtype(dtype).type
    theta_c = np.asarray(theta_c, dtype)
    theta_m = np.asarray(theta_m, dtype)
    refs = np.asarray(refs, dtype)
    ybar = np.asarray(ybar, dtype)
    B, Tr, R = refs.shape
    cfl, cgl = split_params(cspec, theta_c)
    mfl, mgl = split_params(mspec, theta_m)
    _, _, tape = closedloop_fwd(cspec, mspec, theta_c, theta_m, refs, noise, dtype, return_tape=True)
    cfg_, cgg = _zero_grads(cfl), _zero_grads(cgl)
    mfg = _zero_grads(mfl) if with_model_grad else None
    mgg = _zero_grads(mgl) if with_model_grad else None
    lam_c = np.zeros((B, cspec.state_dim), dtype)
    lam_m = np.zeros((B, mspec.state_dim), dtype)
    yfb = np.zeros((B, mspec.output_dim), dtype)  # cotangent arriving through the feedback path
    for k in range(Tr - 1, -1, -1):
        cc_, mc_ = tape[k]
        yb = ybar[:, k] + yfb
        lam_m, abar = node_step_bwd(mspec, mfl, mgl, mc_, yb, lam_m, mfg, mgg)
        lam_c, vbar = node_step_bwd(cspec, cfl, cgl, cc_, abar, lam_c, cfg_, cgg)
        yfb = vbar[:, R:]
    tb = _flatten_grads(cfg_, cgg)
    if with_model_grad:
        # y0 = g_m(x0) receives the last feedback cotangent
        _, g0 = node_y0(mspec, mgl, B, dtype)
        mlp_bwd(mspec.g, mgl, g0, dtype(mspec.out_scale) * yfb, mgg)
        return tb, _flatten_grads(mfg, mgg)
    return tb


# ----------------------------------------------------------------------------------------------
# losses (cc/utils/utils.py:53-55) and their cotangents
# ----------------------------------------------------------------------------------------------
def mse(y, yhat):
    return np.mean((y - yhat) ** 2)


def mse_cotangent(y, yhat, n_total=None):
    """d mse(y, yhat) / d yhat; n_total = global element count when the batch is sharded."""
    n = yhat.size if n_total is None else n_total
    return (2.0 / n) * (yhat - y)
