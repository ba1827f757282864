"""Independent torch (autograd) statement of the reference forward, used to cross-check the oracle's
hand-derived reverse pass.  Written directly from the reference files (neural_ode_model.py:106-149,
neural_ode_controller_compact_example.py:106-129, abstract.py:54-127, integrate.py:1-28); shares no
code with oracle/np_oracle.py except the spec dataclasses."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import np_oracle as O


def _split(mlp, theta, off):
    layers = []
    for i, o in zip(mlp.sizes[:-1], mlp.sizes[1:]):
        W = theta[off:off + o * i].view(o, i)
        off += o * i
        b = None
        if mlp.use_bias:
            b = theta[off:off + o]
            off += o
        layers.append((W, b))
    return layers, off


def _act(kind, x):
    return {O.ACT_IDENTITY: lambda v: v, O.ACT_RELU: torch.relu, O.ACT_TANH: torch.tanh}[kind](x)


def _mlp(mlp, layers, a):
    for li, (W, b) in enumerate(layers):
        a = F.linear(a, W, b)
        a = _act(mlp.act if li < len(layers) - 1 else mlp.act_final, a)
    return a


def _ut(spec, u):
    if spec.u_transform == O.UT_ARCTAN:
        return torch.atan(u)
    if spec.u_transform == O.UT_KTANH:
        return spec.u_scale * torch.tanh(u / spec.u_scale)
    return u


def _cat(parts):
    return torch.cat([p for p in parts if p is not None], 1)


def step(spec, fl, gl, x, t, v):
    B = x.shape[0]
    tcol = lambda tt: torch.full((B, 1), float(tt), dtype=x.dtype)
    ut = _ut(spec, v)
    rhs = lambda tt, z: _mlp(spec.f, fl, _cat([z, tcol(tt) if spec.f_time_variant else None, ut]))
    h = float(np.float32(spec.dt))
    if spec.integrator == O.INT_RK4:
        k1 = rhs(t, x)
        k2 = rhs(t + h / 2, x + h * k1 / 2)
        k3 = rhs(t + h / 2, x + h * k2 / 2)
        k4 = rhs(t + h, x + h * k3)
        xn = x + h * (1 / 6 * (k1 + 2 * k2 + 2 * k3 + k4))
    elif spec.integrator == O.INT_EE:
        xn = x + h * rhs(t, x)
    else:
        xn = rhs(t, x)
    xr = x if spec.readout_pre_step else xn
    gin = _cat([xr, tcol(t) if spec.g_time_variant else None, v if spec.g_feedthrough_u else None])
    return xn, spec.out_scale * _mlp(spec.g, gl, gin)


def y0(spec, gl, B, dtype):
    x = torch.zeros(B, spec.state_dim, dtype=dtype)
    gin = _cat([x, torch.zeros(B, 1, dtype=dtype) if spec.g_time_variant else None])
    return spec.out_scale * _mlp(spec.g, gl, gin)


def rollout(spec, theta, u):
    fl, off = _split(spec.f, theta, 0)
    gl, off = _split(spec.g, theta, off)
    B, T, _ = u.shape
    x = torch.zeros(B, spec.state_dim, dtype=u.dtype)
    ys = [y0(spec, gl, B, u.dtype)]
    t = 0.0
    for k in range(T):
        x, y = step(spec, fl, gl, x, t, u[:, k])
        ys.append(y)
        if spec.has_time:
            t = t + float(np.float32(spec.dt))
    return torch.stack(ys, 1)


def closedloop(cspec, mspec, theta_c, theta_m, refs, noise=None):
    cfl, off = _split(cspec.f, theta_c, 0)
    cgl, off = _split(cspec.g, theta_c, off)
    mfl, off = _split(mspec.f, theta_m, 0)
    mgl, off = _split(mspec.g, theta_m, off)
    B, Tr, _ = refs.shape
    xc = torch.zeros(B, cspec.state_dim, dtype=refs.dtype)
    xm = torch.zeros(B, mspec.state_dim, dtype=refs.dtype)
    y = y0(mspec, mgl, B, refs.dtype)
    tc = tm = 0.0
    ys = []
    for k in range(Tr):
        xc, a = step(cspec, cfl, cgl, xc, tc, torch.cat([refs[:, k], y], 1))
        xm, y = step(mspec, mfl, mgl, xm, tm, a)
        if noise is not None:
            y = y + noise[:, k]
        ys.append(y)
        if cspec.has_time:
            tc += float(np.float32(cspec.dt))
        if mspec.has_time:
            tm += float(np.float32(mspec.dt))
    return torch.stack(ys, 1)
