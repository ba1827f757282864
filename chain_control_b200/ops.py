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
import math
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


def _on_device(fn):
    """every tensor argument must live on ONE CUDA device; the call runs with that device current, so the kernels launch
    on its current stream (tensors on cuda:1 while cuda:0 is current would otherwise launch on GPU 0 with GPU 1 pointers)"""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        devs = {a.device for a in list(args) + list(kwargs.values()) if isinstance(a, torch.Tensor) and a.is_cuda}
        if len(devs) > 1:
            raise ValueError(f"{fn.__name__}: tensors on different devices {sorted(map(str, devs))}")
        if not devs:
            return fn(*args, **kwargs)  # _check raises the no-CPU-path error
        with torch.cuda.device(next(iter(devs))):
            return fn(*args, **kwargs)

    return wrapped


@_on_device
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
    nws = L.cc_rollout_fwd_workspace_bytes(C.byref(m), B, T)
    ws = torch.empty(nws // 4 + 1, device=u.device, dtype=torch.float32) if nws else None
    rc = L.cc_rollout_fwd(C.byref(m), _ptr(theta), _ptr(u), _ptr(x0), _ptr(y), _ptr(xf), _ptr(tape), ckpt_every, B, T,
                          _ptr(ws), nws, _stream())
    L.check(rc, "cc_rollout_fwd")
    return y, xf, tape


@_on_device
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
    nws = L.cc_closedloop_fwd_workspace_bytes(C.byref(c), C.byref(m), B, Tr)
    ws = torch.empty(nws // 4 + 1, device=refs.device, dtype=torch.float32) if nws else None
    rc = L.cc_closedloop_fwd(C.byref(c), C.byref(m), _ptr(theta_c), _ptr(theta_m), _ptr(refs), _ptr(noise), _ptr(yhat),
                             _ptr(us), _ptr(tape), ckpt_every, B, Tr, _ptr(ws), nws, _stream())
    L.check(rc, "cc_closedloop_fwd")
    return yhat, us, tape


@_on_device
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


@_on_device
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


@_on_device
def mse_value_and_cotangent(y, yhat, n_total: Optional[int] = None):
    """default loss of the trainers (cc/utils/utils.py:53-55) fused with its cotangent:
    returns (loss[1] = sum((yhat-y)^2)/n_total over THIS shard, ybar = 2 (yhat-y)/n_total)."""
    L = _lib.lib()
    y = _check(y, "y")
    yhat = _check(yhat, "yhat", tuple(y.shape))
    n = y.numel()
    inv_n = 1.0 / float(n_total if n_total is not None else n)
    ybar = torch.empty_like(yhat)
    part = torch.empty(1 + 296, device=y.device, dtype=torch.float32)
    rc = L.cc_mse_value_and_cotangent(_ptr(y), _ptr(yhat), _ptr(ybar), _ptr(part), n, C.c_float(inv_n), _stream())
    L.check(rc, "cc_mse_value_and_cotangent")
    return part[:1], ybar


class FusedAdam:
    """optax.chain(clip_by_global_norm(c), clip(v), adam(lr)) plus the reference's l1 / l2 regulariser gradients as ONE kernel
    on the flat parameter vector (cc/train/step_fn.py:101-114,166-169; cc/utils/high_level/defaults.py:43-49).  State and
    step counter live on the device: no host round trip, CUDA-graph capturable."""

    def __init__(self, theta: torch.Tensor, lr: float, *, b1=0.9, b2=0.999, eps=1e-8, clip_global_norm=0.0, clip_value=0.0,
                 l1=0.0, l2=0.0):
        self.theta = _check(theta, "theta")
        self.mu, self.nu = torch.zeros_like(self.theta), torch.zeros_like(self.theta)
        self.count = torch.zeros(1, dtype=torch.int32, device=theta.device)
        self.out = torch.zeros(3, dtype=torch.float32, device=theta.device)
        self.hp = (float(lr), float(b1), float(b2), float(eps), float(clip_global_norm), float(clip_value), float(l1), float(l2))

    def step(self, grad: torch.Tensor, grad_scale: float = 1.0):
        L = _lib.lib()
        if grad.device != self.theta.device:
            raise ValueError(f"grad on {grad.device}, parameters on {self.theta.device}")
        grad = _check(grad, "grad", tuple(self.theta.shape))
        with torch.cuda.device(self.theta.device):
            rc = L.cc_opt_step(_ptr(self.theta), _ptr(self.mu), _ptr(self.nu), _ptr(grad), _ptr(self.count), _ptr(self.out),
                               self.theta.numel(), *[C.c_float(v) for v in self.hp], C.c_float(grad_scale), _stream())
        L.check(rc, "cc_opt_step")
        return self.theta


# ---- checkpointing policy of the differentiable entry points ---------------------------------------------------------------
# The reference stores every per-step residual (lax.scan transpose, cc/core/module_utils.py:17); here the tape is the scan
# carry, and when even that would not fit comfortably the autograd functions switch to checkpointed recomputation
# (ckpt_every = K > 1: a checkpoint every K steps, segments replayed in the reverse pass) on their own.
_CKPT = {"every": None, "max_tape_bytes": None}


def set_checkpointing(every: Optional[int] = None, max_tape_bytes: Optional[int] = None):
    """every = K forces ckpt_every = K for batched_unroll / batched_closed_loop (None: automatic); max_tape_bytes = the tape
    size above which the automatic policy checkpoints (None: a third of the device's memory)."""
    if every is not None and every < 1:
        raise ValueError("every must be >= 1")
    _CKPT["every"], _CKPT["max_tape_bytes"] = every, max_tape_bytes


def _ckpt_every(tape_bytes_at_1: int, T: int, device) -> int:
    if _CKPT["every"] is not None:
        return int(_CKPT["every"])
    budget = _CKPT["max_tape_bytes"]
    if budget is None:
        budget = torch.cuda.get_device_properties(device).total_memory // 3
    if tape_bytes_at_1 <= budget or T < 4:
        return 1
    return max(2, math.isqrt(T - 1) + 1)  # ceil(sqrt(T)): balances the checkpoints (T / K) against the segment tape (K)


class _RolloutFn(torch.autograd.Function):
    """y = vmap(model.unroll)(u); backward = the BPTT kernel (custom VJP, generic cotangent)."""

    @staticmethod
    def forward(ctx, theta, u, x0, spec):
        need_grad = theta.requires_grad or u.requires_grad
        K = 1
        if need_grad:
            m = spec.to_c()
            K = _ckpt_every(_lib.lib().cc_rollout_tape_bytes(C.byref(m), u.shape[0], u.shape[1], 1), u.shape[1], u.device)
        y, _, tape = rollout_fwd(spec, theta, u, x0, want_tape=need_grad, ckpt_every=K)
        ctx.spec, ctx.tape, ctx.K = spec, tape, K
        ctx.save_for_backward(theta, u, x0 if x0 is not None else torch.empty(0, device=u.device))
        ctx.has_x0 = x0 is not None
        ctx.want_ubar = u.requires_grad
        return y

    @staticmethod
    def backward(ctx, ybar):
        theta, u, x0 = ctx.saved_tensors
        tb, ub = rollout_bwd(ctx.spec, theta, u, ctx.tape, ybar.contiguous(), x0 if ctx.has_x0 else None,
                             ckpt_every=ctx.K, want_ubar=ctx.want_ubar)
        return tb, ub, None, None


class _ClosedLoopFn(torch.autograd.Function):
    """yhat = vmap(controller.unroll(model, merge_x_y, noise))(refs); gradient w.r.t. the controller only."""

    @staticmethod
    def forward(ctx, theta_c, theta_m, refs, noise, cspec, mspec):
        K = 1
        if theta_c.requires_grad:
            c, m = cspec.to_c(), mspec.to_c()
            K = _ckpt_every(_lib.lib().cc_closedloop_tape_bytes(C.byref(c), C.byref(m), refs.shape[0], refs.shape[1], 1), refs.shape[1], refs.device)
        yhat, _, tape = closedloop_fwd(cspec, mspec, theta_c, theta_m, refs, noise, want_tape=theta_c.requires_grad, ckpt_every=K)
        ctx.cspec, ctx.mspec, ctx.tape, ctx.K = cspec, mspec, tape, K
        ctx.save_for_backward(theta_c, theta_m, refs, noise if noise is not None else torch.empty(0, device=refs.device))
        ctx.has_noise = noise is not None
        return yhat

    @staticmethod
    def backward(ctx, ybar):
        theta_c, theta_m, refs, noise = ctx.saved_tensors
        tb = closedloop_bwd(ctx.cspec, ctx.mspec, theta_c, theta_m, refs, ctx.tape, ybar.contiguous(),
                            noise if ctx.has_noise else None, ckpt_every=ctx.K)
        return tb, None, None, None, None, None


def batched_unroll(spec: NodeSpec, theta, u, x0=None):
    """differentiable drop-in for eqx.filter_vmap(model.unroll)(inputs)   (cc/train/step_fn.py:124,136)."""
    return _RolloutFn.apply(theta, u, x0, spec)


def batched_closed_loop(cspec: NodeSpec, mspec: NodeSpec, theta_c, theta_m, refs, noise=None):
    """differentiable drop-in for eqx.filter_vmap(controller.unroll(model, merge_x_y, noise))(refss, keys)
    (cc/train/step_fn.py:234-236,257-259); `noise` is the already-scaled additive output noise."""
    return _ClosedLoopFn.apply(theta_c, theta_m, refs, noise, cspec, mspec)
