"""Torch-tensor front-end of the C ABI: the batched unroll operators and their custom VJPs.

These are the drop-in replacements for
    eqx.filter_vmap(model.unroll)(inputs)                                  cc/train/step_fn.py:124,136
    eqx.filter_vmap(controller.unroll(model, merge_x_y, noise))(refss, keys)  cc/train/step_fn.py:234-236,257-259
and, through torch.autograd.Function, for what eqx.filter_value_and_grad derives from them
(step_fn.py:133,249-252).  torch is only plumbing here (device memory, streams, autograd graph);
all arithmetic happens in libccb200.so.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .arch import NodeSpec


def _check(t: Optional[torch.Tensor], name: str, shape=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.CcError(f"{name} must be a CUDA tensor: chain_control_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    t = t.contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def rollout_fwd(spec: NodeSpec, theta, u, x0=None, *, want_tape=False, ckpt_every=1, want_x_final=False):
    """u[B,T,I] -> y[B,T+1,O]  (AbstractModel.unroll with include_y0=True, vmapped over B)."""
    L = _lib.lib()
    m = spec.to_c()
    B, T, I = u.shape
    if I != spec.input_dim:
        raise ValueError(f"u has {I} input channels, the model expects {spec.input_dim}")
    theta = _check(theta, "theta", (spec.n_params(),))
    u = _check(u, "u")
    x0 = _check(x0, "x0", (B, spec.state_dim))
    y = torch.empty((B, T + 1, spec.output_dim), device=u.device, dtype=torch.float32)
    xf = torch.empty((B, spec.state_dim), device=u.device, dtype=torch.float32) if want_x_final else None
    tape = None
    if want_tape:
        nbytes = L.cc_rollout_tape_bytes(C.byref(m), B, T, ckpt_every)
        tape = torch.empty(max(nbytes // 4, 1), device=u.device, dtype=torch.float32)
    rc = L.cc_rollout_fwd(C.byref(m), _ptr(theta), _ptr(u), _ptr(x0), _ptr(y), _ptr(xf), _ptr(tape), ckpt_every, B, T,
                          _stream())
    L.check(rc, "cc_rollout_fwd")
    return y, xf, tape


def closedloop_fwd(cspec: NodeSpec, mspec: NodeSpec, theta_c, theta_m, refs, noise=None, *, want_tape=False,
                   ckpt_every=1, want_u=False):
    """refs[B,Tr,R] -> yhat[B,Tr,O]  (AbstractController.unroll closed loop, vmapped over B)."""
    L = _lib.lib()
    c, m = cspec.to_c(), mspec.to_c()
    B, Tr, R = refs.shape
    if R + mspec.output_dim != cspec.input_dim:
        raise ValueError("controller input_dim must equal ref_dim + model output_dim (merge_x_y = [ref; obs])")
    theta_c = _check(theta_c, "theta_c", (cspec.n_params(),))
    theta_m = _check(theta_m, "theta_m", (mspec.n_params(),))
    refs = _check(refs, "refs")
    noise = _check(noise, "noise", (B, Tr, mspec.output_dim))
    yhat = torch.empty((B, Tr, mspec.output_dim), device=refs.device, dtype=torch.float32)
    us = torch.empty((B, Tr, mspec.input_dim), device=refs.device, dtype=torch.float32) if want_u else None
    tape = None
    if want_tape:
        nbytes = L.cc_closedloop_tape_bytes(C.byref(c), C.byref(m), B, Tr, ckpt_every)
        tape = torch.empty(max(nbytes // 4, 1), device=refs.device, dtype=torch.float32)
    rc = L.cc_closedloop_fwd(C.byref(c), C.byref(m), _ptr(theta_c), _ptr(theta_m), _ptr(refs), _ptr(noise), _ptr(yhat),
                             _ptr(us), _ptr(tape), ckpt_every, B, Tr, _stream())
    L.check(rc, "cc_closedloop_fwd")
    return yhat, us, tape


def rollout_bwd(spec: NodeSpec, theta, u, tape, ybar, x0=None, *, ckpt_every=1, want_ubar=False):
    """VJP of rollout_fwd: ybar[B,T+1,O] -> theta_bar[P] (and ubar[B,T,I])."""
    L = _lib.lib()
    m = spec.to_c()
    B, T, I = u.shape
    theta = _check(theta, "theta", (spec.n_params(),))
    u = _check(u, "u")
    x0 = _check(x0, "x0", (B, spec.state_dim))
    ybar = _check(ybar, "ybar", (B, T + 1, spec.output_dim))
    nws = L.cc_rollout_workspace_bytes(C.byref(m), B, T, ckpt_every)
    ws = torch.empty(max(nws // 4, 1), device=u.device, dtype=torch.float32)
    tb = torch.empty(spec.n_params(), device=u.device, dtype=torch.float32)
    ub = torch.empty_like(u) if want_ubar else None
    rc = L.cc_rollout_bwd(C.byref(m), _ptr(theta), _ptr(u), _ptr(x0), _ptr(tape), ckpt_every, _ptr(ybar), _ptr(tb),
                          _ptr(ub), B, T, _ptr(ws), ws.numel() * 4, _stream())
    L.check(rc, "cc_rollout_bwd")
    return tb, ub


def closedloop_bwd(cspec: NodeSpec, mspec: NodeSpec, theta_c, theta_m, refs, tape, ybar, noise=None, *, ckpt_every=1):
    """VJP of closedloop_fwd w.r.t. the controller parameters: ybar[B,Tr,O] -> theta_c_bar[P_c]."""
    L = _lib.lib()
    c, m = cspec.to_c(), mspec.to_c()
    B, Tr, R = refs.shape
    theta_c = _check(theta_c, "theta_c", (cspec.n_params(),))
    theta_m = _check(theta_m, "theta_m", (mspec.n_params(),))
    refs = _check(refs, "refs")
    noise = _check(noise, "noise", (B, Tr, mspec.output_dim))
    ybar = _check(ybar, "ybar", (B, Tr, mspec.output_dim))
    nws = L.cc_closedloop_workspace_bytes(C.byref(c), C.byref(m), B, Tr, ckpt_every)
    ws = torch.empty(max(nws // 4, 1), device=refs.device, dtype=torch.float32)
    tb = torch.empty(cspec.n_params(), device=refs.device, dtype=torch.float32)
    rc = L.cc_closedloop_bwd(C.byref(c), C.byref(m), _ptr(theta_c), _ptr(theta_m), _ptr(refs), _ptr(noise), _ptr(tape),
                             ckpt_every, _ptr(ybar), _ptr(tb), B, Tr, _ptr(ws), ws.numel() * 4, _stream())
    L.check(rc, "cc_closedloop_bwd")
    return tb
