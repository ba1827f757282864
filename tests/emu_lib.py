"""Loads the CPU-emulated build of the kernels (tests/emu) behind the same ctypes wrapper as the
product library, and offers numpy-level helpers used by both the emulated (CPU) and the real (GPU)
parity tests."""
import ctypes as C
import os
import subprocess

import numpy as np

from chain_control_b200._lib import CcLib

_HERE = os.path.dirname(os.path.abspath(__file__))
_EMU = os.path.join(_HERE, "emu", "libccb200_emu.so")
_emu = None


def emu() -> CcLib:
    global _emu
    if _emu is None:
        srcdir = os.path.join(_HERE, "..", "chain_control_b200", "csrc")
        srcs = [os.path.join(srcdir, f) for f in os.listdir(srcdir) if f.endswith((".cu", ".cuh", ".h", ".inc"))]
        srcs.append(os.path.join(_HERE, "emu", "cuda_emu.h"))
        if not os.path.exists(_EMU) or any(os.path.getmtime(s) > os.path.getmtime(_EMU) for s in srcs):
            subprocess.run([os.path.join(_HERE, "emu", "build_emu.sh")], check=True)
        _emu = CcLib(_EMU)
    return _emu


def _p(a):
    return None if a is None else a.ctypes.data


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def np_rollout_fwd(L, spec, theta, u, x0=None, tape_every=0):
    """host-pointer call (emulated library only)."""
    theta, u, x0 = _f32(theta), _f32(u), _f32(x0)
    B, T, _ = u.shape
    m = spec.to_c() if hasattr(spec, "to_c") else spec
    y = np.full((B, T + 1, m.output_dim), np.nan, np.float32)
    xf = np.full((B, m.state_dim), np.nan, np.float32)
    tape = None
    if tape_every:
        nbytes = L.cc_rollout_tape_bytes(C.byref(m), B, T, tape_every)
        tape = np.full(nbytes // 4, np.nan, np.float32)
    nws = L.cc_rollout_fwd_workspace_bytes(C.byref(m), B, T)
    ws = np.full(nws // 4 + 1, np.nan, np.float32)
    rc = L.cc_rollout_fwd(C.byref(m), _p(theta), _p(u), _p(x0), _p(y), _p(xf), _p(tape), tape_every or 1, B, T,
                          _p(ws) if nws else None, nws, None)
    L.check(rc, "cc_rollout_fwd")
    return y, xf, tape


def np_closedloop_fwd(L, cspec, mspec, theta_c, theta_m, refs, noise=None, tape_every=0, want_u=True):
    theta_c, theta_m, refs, noise = _f32(theta_c), _f32(theta_m), _f32(refs), _f32(noise)
    B, Tr, _ = refs.shape
    c, m = cspec.to_c(), mspec.to_c()
    yhat = np.full((B, Tr, m.output_dim), np.nan, np.float32)
    us = np.full((B, Tr, m.input_dim), np.nan, np.float32)
    tape = None
    if tape_every:
        nbytes = L.cc_closedloop_tape_bytes(C.byref(c), C.byref(m), B, Tr, tape_every)
        tape = np.full(nbytes // 4, np.nan, np.float32)
    nws = L.cc_closedloop_fwd_workspace_bytes(C.byref(c), C.byref(m), B, Tr)
    ws = np.full(nws // 4 + 1, np.nan, np.float32)
    rc = L.cc_closedloop_fwd(C.byref(c), C.byref(m), _p(theta_c), _p(theta_m), _p(refs), _p(noise), _p(yhat),
                             _p(us) if want_u else None, _p(tape), tape_every or 1, B, Tr, _p(ws) if nws else None, nws, None)
    L.check(rc, "cc_closedloop_fwd")
    return yhat, us, tape


def np_rollout_bwd(L, spec, theta, u, ybar, x0=None, want_ubar=True, ckpt_every=1):
    theta, u, x0, ybar = _f32(theta), _f32(u), _f32(x0), _f32(ybar)
    B, T, _ = u.shape
    m = spec.to_c()
    y, xf, tape = np_rollout_fwd(L, spec, theta, u, x0, tape_every=ckpt_every)
    nws = L.cc_rollout_workspace_bytes(C.byref(m), B, T, ckpt_every)
    ws = np.full(max(nws // 4, 1), np.nan, np.float32)
    tb = np.full(spec.n_params(), np.nan, np.float32)
    ub = np.full(u.shape, np.nan, np.float32) if want_ubar else None
    rc = L.cc_rollout_bwd(C.byref(m), _p(theta), _p(u), _p(x0), _p(tape), ckpt_every, _p(ybar), _p(tb), _p(ub), B, T, _p(ws),
                          ws.nbytes, None)
    L.check(rc, "cc_rollout_bwd")
    return y, tb, ub


def np_closedloop_bwd(L, cspec, mspec, theta_c, theta_m, refs, ybar, noise=None, ckpt_every=1):
    theta_c, theta_m, refs, noise, ybar = _f32(theta_c), _f32(theta_m), _f32(refs), _f32(noise), _f32(ybar)
    B, Tr, _ = refs.shape
    c, m = cspec.to_c(), mspec.to_c()
    # (no control-series output: with no noise either, the forward kernels take their folded fast path)
    yhat, us, tape = np_closedloop_fwd(L, cspec, mspec, theta_c, theta_m, refs, noise, tape_every=ckpt_every, want_u=False)
    nws = L.cc_closedloop_workspace_bytes(C.byref(c), C.byref(m), B, Tr, ckpt_every)
    ws = np.full(max(nws // 4, 1), np.nan, np.float32)
    tb = np.full(cspec.n_params(), np.nan, np.float32)
    rc = L.cc_closedloop_bwd(C.byref(c), C.byref(m), _p(theta_c), _p(theta_m), _p(refs), _p(noise), _p(tape), ckpt_every,
                             _p(ybar), _p(tb), B, Tr, _p(ws), ws.nbytes, None)
    L.check(rc, "cc_closedloop_bwd")
    return yhat, tb
