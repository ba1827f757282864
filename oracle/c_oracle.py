"""ctypes front-end of the C oracle (oracle/cc_oracle.c).  TEST INFRASTRUCTURE — see np_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CC_MAX_LAYERS = 4


class CcMlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("sizes", C.c_int32 * (CC_MAX_LAYERS + 1)), ("act", C.c_int32),
                ("act_final", C.c_int32), ("use_bias", C.c_int32)]


class CcModule(C.Structure):
    _fields_ = [("state_dim", C.c_int32), ("input_dim", C.c_int32), ("output_dim", C.c_int32),
                ("f", CcMlp), ("g", CcMlp), ("integrator", C.c_int32), ("f_time_variant", C.c_int32),
                ("g_time_variant", C.c_int32), ("g_feedthrough_u", C.c_int32),
                ("readout_pre_step", C.c_int32), ("u_transform", C.c_int32), ("u_scale", C.c_float),
                ("out_scale", C.c_float), ("dt", C.c_float), ("t0", C.c_float)]


def to_c_module(spec) -> CcModule:
    """duck-typed: anything with the NodeSpec attribute names (oracle or product spec)."""
    def mlp(m):
        s = CcMlp()
        s.n_layers = len(m.sizes) - 1
        assert s.n_layers <= CC_MAX_LAYERS
        for i, v in enumerate(m.sizes):
            s.sizes[i] = int(v)
        s.act, s.act_final, s.use_bias = int(m.act), int(m.act_final), int(bool(m.use_bias))
        return s

    c = CcModule()
    c.state_dim, c.input_dim, c.output_dim = spec.state_dim, spec.input_dim, spec.output_dim
    c.f, c.g = mlp(spec.f), mlp(spec.g)
    c.integrator = int(spec.integrator)
    c.f_time_variant, c.g_time_variant = int(spec.f_time_variant), int(spec.g_time_variant)
    c.g_feedthrough_u, c.readout_pre_step = int(spec.g_feedthrough_u), int(spec.readout_pre_step)
    c.u_transform, c.u_scale = int(spec.u_transform), float(spec.u_scale)
    c.out_scale, c.dt = float(spec.out_scale), float(spec.dt)
    c.t0 = float(getattr(spec, 't0', 0.0))
    return c


def build(march: str = "x86-64-v3", out: str | None = None, force: bool = False) -> str:
    """compile the C oracle; returns the .so path.  march='native' is used for the CPU baseline."""
    out = out or os.path.join(_HERE, "libcc_oracle.so" if march == "x86-64-v3" else f"libcc_oracle_{march}.so")
    src = [os.path.join(_HERE, f) for f in ("cc_oracle.c", "cc_oracle_impl.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in src):
        return out
    cmd = ["/usr/bin/gcc", "-O3", f"-march={march}", "-fopenmp", "-fPIC", "-std=gnu11", "-shared", "-o", out,
           src[0], "-lm"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return out


_libs = {}


def lib(march: str = "x86-64-v3"):
    if march not in _libs:
        L = C.CDLL(build(march))
        L.cco_param_count.restype = C.c_int64
        L.cco_max_threads.restype = C.c_int
        _libs[march] = L
    return _libs[march]


def max_threads() -> int:
    return int(lib().cco_max_threads())


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _prep(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def rollout(spec, theta, u, x0=None, ybar=None, targets=None, want_ubar=False, dtype=np.float32,
            nthreads=0, march="x86-64-v3"):
    """forward (ybar is None and targets is None) or forward+backward.
    returns dict(y=..., theta_bar=..., ubar=..., loss=...)."""
    dt = np.dtype(dtype)
    ct = C.c_float if dt == np.float32 else C.c_double
    fn = getattr(lib(march), "cco_rollout_f32" if dt == np.float32 else "cco_rollout_f64")
    theta, u, x0 = _prep(theta, dt), _prep(u, dt), _prep(x0, dt)
    B, T, I = u.shape
    m = to_c_module(spec)
    y = np.empty((B, T + 1, spec.output_dim), dt)
    mode = 0
    cot = None
    if ybar is not None:
        mode, cot = 1, _prep(ybar, dt)
    elif targets is not None:
        mode, cot = 2, _prep(targets, dt)
    P = int(lib(march).cco_param_count(C.byref(m)))
    assert theta.size == P, (theta.size, P)
    tb = np.zeros(P, dt) if mode else None
    ub = np.zeros_like(u) if (mode and want_ubar) else None
    loss = C.c_double(0.0)
    rc = fn(C.byref(m), _ptr(theta, ct), _ptr(u, ct), _ptr(x0, ct), _ptr(y, ct), C.c_int(mode), _ptr(cot, ct),
            _ptr(tb, ct), _ptr(ub, ct), C.byref(loss), C.c_int64(B), C.c_int32(T), C.c_int32(nthreads))
    assert rc == 0, rc
    return dict(y=y, theta_bar=tb, ubar=ub, loss=loss.value if mode == 2 else None)


def closedloop(cspec, mspec, theta_c, theta_m, refs, noise=None, ybar=None, mse_loss=False,
               with_model_grad=False, dtype=np.float32, nthreads=0, march="x86-64-v3"):
    dt = np.dtype(dtype)
    ct = C.c_float if dt == np.float32 else C.c_double
    fn = getattr(lib(march), "cco_closedloop_f32" if dt == np.float32 else "cco_closedloop_f64")
    theta_c, theta_m, refs, noise = _prep(theta_c, dt), _prep(theta_m, dt), _prep(refs, dt), _prep(noise, dt)
    B, Tr, R = refs.shape
    c, m = to_c_module(cspec), to_c_module(mspec)
    yhat = np.empty((B, Tr, mspec.output_dim), dt)
    us = np.empty((B, Tr, mspec.input_dim), dt)
    mode, cot = 0, None
    if ybar is not None:
        mode, cot = 1, _prep(ybar, dt)
    elif mse_loss:
        mode = 2
    tcb = np.zeros(cspec.n_params(), dt) if mode else None
    tmb = np.zeros(mspec.n_params(), dt) if (mode and with_model_grad) else None
    loss = C.c_double(0.0)
    rc = fn(C.byref(c), C.byref(m), _ptr(theta_c, ct), _ptr(theta_m, ct), _ptr(refs, ct), _ptr(noise, ct),
            _ptr(yhat, ct), _ptr(us, ct), C.c_int(mode), _ptr(cot, ct), _ptr(tcb, ct), _ptr(tmb, ct),
            C.byref(loss), C.c_int64(B), C.c_int32(Tr), C.c_int32(nthreads))
    assert rc == 0, rc
    return dict(yhat=yhat, u=us, theta_c_bar=tcb, theta_m_bar=tmb, loss=loss.value if mode == 2 else None)
