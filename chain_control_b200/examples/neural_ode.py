"""Neural-ODE model / controller modules mirroring cc/examples/neural_ode_model.py:16-177,
neural_ode_controller.py:16-160 and the compact variants (neural_ode_model_compact_example.py:32-124,
neural_ode_controller_compact_example.py:33-181) on top of the CUDA rollout kernels.

A module stores its trainable parameters as ONE flat fp32 CUDA vector in the reference's
flatten_module order (cc/core/module_utils.py:20-24); `f` / `g` expose per-layer views of it.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Callable, Optional

import numpy as np
import torch

from .. import arch as A
from .. import ops
from ..core.abstract import AbstractController, AbstractModel
from ..utils import batch_concat, make_postprocess_fn, struct_of_specs


# ------------------------------------------------------------------------------------------------
# callables -> kernel enums.  The reference passes Python callables (jax.nn.relu, jnp.tanh, lambdas);
# the kernels support a closed set, identified by probing.  Anything else raises (no silent fallback).
# ------------------------------------------------------------------------------------------------
_PROBE = np.array([-2.5, -0.7, 0.0, 0.3, 1.9], np.float64)


def _apply(fn, x):
    try:
        return np.asarray(fn(x), np.float64)  # numpy-compatible callables (np.tanh, lambdas, ...)
    except Exception:
        return np.asarray(fn(torch.tensor(x)).numpy(), np.float64)  # torch callables


def classify_activation(fn) -> int:
    if isinstance(fn, str):
        return {"identity": A.ACT_IDENTITY, "relu": A.ACT_RELU, "tanh": A.ACT_TANH}[fn]
    y = _apply(fn, _PROBE)
    for kind, ref in ((A.ACT_IDENTITY, _PROBE), (A.ACT_RELU, np.maximum(_PROBE, 0)), (A.ACT_TANH, np.tanh(_PROBE))):
        if np.allclose(y, ref, rtol=1e-5, atol=1e-6):
            return kind
    raise NotImplementedError("activation must be one of identity / relu / tanh (the CUDA kernels have no generic callable path)")


def classify_u_transform(fn):
    if isinstance(fn, str):
        return {"identity": A.UT_IDENTITY, "arctan": A.UT_ARCTAN}[fn], 1.0
    y = _apply(fn, _PROBE)
    if np.allclose(y, _PROBE, rtol=1e-5, atol=1e-6):
        return A.UT_IDENTITY, 1.0
    if np.allclose(y, np.arctan(_PROBE), rtol=1e-5, atol=1e-6):
        return A.UT_ARCTAN, 1.0
    k = float(_apply(fn, np.array([1e9]))[0])  # k * tanh(u / k) saturates at k
    if k > 0 and np.allclose(y, k * np.tanh(_PROBE / k), rtol=1e-5, atol=1e-6):
        return A.UT_KTANH, k
    raise NotImplementedError("u_transform must be identity, arctan or k*tanh(u/k)")


def _key_to_seed(key) -> int:
    if key is None:
        return 1
    if isinstance(key, (int, np.integer)):
        return int(key)
    arr = np.asarray(key).ravel()
    return int(arr[-1]) + 7919 * int(arr[0]) if arr.size else 1


def init_flat_params(spec: A.NodeSpec, key) -> torch.Tensor:
    """eqx.nn.Linear init: weight and bias ~ U(-1/sqrt(in), 1/sqrt(in)).  (JAX's threefry stream cannot be
    reproduced without jax; the draw is distribution-equivalent, seeded by `key`.)"""
    g = torch.Generator().manual_seed(_key_to_seed(key))
    parts = []
    for mlp in (spec.f, spec.g):
        for i, o in zip(mlp.sizes[:-1], mlp.sizes[1:]):
            lim = 1.0 / math.sqrt(i)
            parts.append((torch.rand(o * i, generator=g) * 2 - 1) * lim)
            if mlp.use_bias:
                parts.append((torch.rand(o, generator=g) * 2 - 1) * lim)
    return torch.cat(parts).to(torch.float32)


class _Layer:
    """view of one eqx.nn.Linear inside the flat vector"""

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias


def _mlp_views(mlp: A.MlpSpec, theta, off):
    layers = []
    for i, o in zip(mlp.sizes[:-1], mlp.sizes[1:]):
        W = theta[off:off + o * i].view(o, i)
        off += o * i
        b = None
        if mlp.use_bias:
            b = theta[off:off + o]
            off += o
        layers.append(_Layer(W, b))
    return layers, off


class _NeuralOdeBase:
    def __init__(self, spec: A.NodeSpec, theta: torch.Tensor, state, init_state, out_struct):
        self.spec = spec
        self.theta = theta
        self.state = state            # (x[S] tensor, t float)
        self.init_state = init_state
        # decided once (host sync here, not on every train step): the kernels skip the x0 load for the all-zero initial state
        self._x0_nonzero = bool(torch.any(init_state[0] != 0))
        self._out_struct = out_struct
        self._post = make_postprocess_fn(out_struct)

    # ---- parameters --------------------------------------------------------------------------
    def flat_params(self) -> torch.Tensor:
        return self.theta

    def with_params(self, theta: torch.Tensor):
        return type(self)(self.spec, theta, self.state, self.init_state, self._out_struct)

    @property
    def f(self):
        return _mlp_views(self.spec.f, self.theta, 0)[0]

    @property
    def g(self):
        return _mlp_views(self.spec.g, self.theta, self.spec.f.n_params())[0]

    def grad_filter_spec(self):
        # neural_ode_model.py:151-158: f and g are trained, state / init_state are not
        return OrderedDict(f=True, g=True, state=False, init_state=False)

    def reset(self):
        return type(self)(self.spec, self.theta, self.init_state, self.init_state, self._out_struct)

    def _spec_at_state(self):
        x, t = self.state
        if t == self.spec.t0:
            return self.spec
        s = A.NodeSpec(**{**self.spec.__dict__})
        s.t0 = float(t)
        return s

    def _step_flat(self, u_flat):
        """one step from self.state through a B=1, T=1 kernel call -> (new state, flat y)"""
        x, t = self.state
        dev = self.theta.device
        u = torch.as_tensor(u_flat, dtype=torch.float32, device=dev).reshape(1, 1, self.spec.input_dim)
        y, xf, _ = ops.rollout_fwd(self._spec_at_state(), self.theta.detach(), u, x.reshape(1, -1).to(dev), want_x_final=True)
        t_next = float(np.float32(t) + np.float32(self.spec.dt)) if self.spec.has_time else t
        return (xf[0], t_next), y[0, 1]


class NeuralOdeModel(_NeuralOdeBase, AbstractModel):
    """NeuralOdeModel of neural_ode_model.py:97-177 (readout after the update) and of the compact
    example :85-124 (readout before the update, `readout_pre_step`)."""

    def step(self, u):
        state, y = self._step_flat(batch_concat(u, 0))
        return NeuralOdeModel(self.spec, self.theta, state, self.init_state, self._out_struct), self._post(y)

    def y0(self):
        x, t = self.init_state
        dev = self.theta.device
        s = A.NodeSpec(**{**self.spec.__dict__})
        s.t0 = float(t)
        y, _, _ = ops.rollout_fwd(s, self.theta.detach(), torch.zeros(1, 0, self.spec.input_dim, device=dev), x.reshape(1, -1).to(dev))
        return self._post(y[0, 0])

    def unroll(self, time_series_of_x, include_y0: bool = True):
        u = batch_concat(time_series_of_x, 1).to(self.theta.device)
        y = self.unroll_batched(u[None], include_y0, _flat=True)[0]
        return self._post(y)

    def unroll_batched(self, inputs, include_y0: bool = True, _flat: bool = False):
        """eqx.filter_vmap(model.unroll)(inputs): inputs [B,T,I] (or a pytree of such) -> [B,T+1,O].
        Differentiable w.r.t. the model's flat parameter vector (BPTT kernel)."""
        u = batch_concat(inputs, 2).to(self.theta.device, torch.float32)
        x0, t0 = self.init_state
        x0b = None if not self._x0_nonzero else x0.to(u.device).expand(u.shape[0], -1).contiguous()
        spec = self.spec if t0 == self.spec.t0 else self._spec_at_state()
        y = ops.batched_unroll(spec, self.theta, u.contiguous(), x0b)
        if not include_y0:
            y = y[:, 1:]
        return y if _flat else self._post(y)


class _ClosedLoopUnroll:
    """what controller.unroll(model, merge_x_y, noise_scale) returns (abstract.py:103-127)."""

    def __init__(self, controller, model, merge_x_y, noise_scale):
        # the model is never trained inside a closed loop (step_fn.py:249-252): a LinearModel drops its trainable-D bookkeeping
        model = model.frozen() if hasattr(model, "frozen") else model
        self.controller, self.model, self.merge_x_y, self.noise_scale = controller, model, merge_x_y, noise_scale
        probe = merge_x_y("ref", "obs")
        if list(probe.keys()) != ["ref", "obs"]:
            raise NotImplementedError("the closed-loop kernel implements merge_x_y = OrderedDict(ref=..., obs=...) (step_fn.py:187-192)")

    def _noise(self, shape, keys, device):
        if self.noise_scale is None:
            return None
        g = torch.Generator(device=device).manual_seed(_key_to_seed(keys))
        return self.noise_scale * torch.randn(shape, generator=g, device=device, dtype=torch.float32)

    def batched(self, refss, keys=None):
        c, m = self.controller, self.model
        refs = batch_concat(refss, 2).to(c.theta.device, torch.float32).contiguous()
        noise = self._noise((refs.shape[0], refs.shape[1], m.spec.output_dim), keys, refs.device)
        y = ops.batched_closed_loop(c.spec, m.spec, c.theta, m.theta.detach(), refs, noise)
        return m._post(y)

    def __call__(self, time_series_of_x, key=None):
        refs = batch_concat(time_series_of_x, 1)
        out = self.batched(refs[None], key)
        from ..utils import tree_map
        return tree_map(lambda a: a[0], out)


class NeuralOdeController(_NeuralOdeBase, AbstractController):
    """NeuralOdeController of neural_ode_controller.py:99-160 and of the compact example :88-181
    (pre-step readout, output scaled by sigmoid(alpha) = 0.5 with the zero secondary controller)."""

    def step(self, x):
        state, y = self._step_flat(batch_concat(x, 0))
        return NeuralOdeController(self.spec, self.theta, state, self.init_state, self._out_struct), self._post(y)

    def unroll(self, model: AbstractModel, merge_x_y: Callable, noise_scale: Optional[float] = None):
        return _ClosedLoopUnroll(self, model, merge_x_y, noise_scale)


def _build(cls, input_specs, output_specs, control_timestep, state_dim, key, device, **kw):
    I, _ = struct_of_specs(input_specs)
    O, out_struct = struct_of_specs(output_specs)
    spec = A.make_node_spec(state_dim, I, O, control_timestep, **kw)
    theta = init_flat_params(spec, key).to(device)
    init_state = (torch.zeros(state_dim, device=device), 0.0)
    return cls(spec, theta, init_state, init_state, out_struct)


def make_neural_ode_model(input_specs, output_specs, control_timestep: float, state_dim: int, key=1,
                          f_integrate_method: str = "RK4", f_use_bias: bool = True, f_time_invariant: bool = True,
                          f_use_dropout: bool = False, f_dropout_rate: float = 0.4, f_width_size: int = 10,
                          f_depth: int = 2, f_activation="relu", f_final_activation="identity",
                          g_use_bias: bool = True, g_time_invariant: bool = True, g_use_dropout: bool = False,
                          g_dropout_rate: float = 0.4, g_width_size: int = 10, g_depth: int = 0, g_activation="relu",
                          g_final_activation="identity", u_transform="identity", device="cuda") -> NeuralOdeModel:
    """same keyword surface as cc/examples/neural_ode_model.py:16-40"""
    if f_use_dropout or g_use_dropout:
        raise NotImplementedError  # as in the reference (:41-43)
    ut, k = classify_u_transform(u_transform)
    return _build(NeuralOdeModel, input_specs, output_specs, control_timestep, state_dim, key, device,
                  f_width=f_width_size, f_depth=f_depth, g_width=g_width_size, g_depth=g_depth,
                  f_act=classify_activation(f_activation), f_act_final=classify_activation(f_final_activation),
                  g_act=classify_activation(g_activation), g_act_final=classify_activation(g_final_activation),
                  f_use_bias=f_use_bias, g_use_bias=g_use_bias, integrator=A.INTEGRATORS[f_integrate_method],
                  f_time_variant=not f_time_invariant, g_time_variant=not g_time_invariant, u_transform=ut, u_scale=k)


def make_neural_ode_model_compact(input_specs, output_specs, control_timestep: float, state_dim: int, key=1,
                                  f_integrate_method: str = "RK4", f_width_size: int = 10, f_depth: int = 2,
                                  f_activation="relu", f_final_activation="identity", g_width_size: int = 10,
                                  g_depth: int = 0, g_activation="relu", g_final_activation="identity",
                                  u_transform="identity", device="cuda") -> NeuralOdeModel:
    """cc/examples/neural_ode_model_compact_example.py:32-49: time-invariant, readout BEFORE the update (:103)"""
    ut, k = classify_u_transform(u_transform)
    return _build(NeuralOdeModel, input_specs, output_specs, control_timestep, state_dim, key, device,
                  f_width=f_width_size, f_depth=f_depth, g_width=g_width_size, g_depth=g_depth,
                  f_act=classify_activation(f_activation), f_act_final=classify_activation(f_final_activation),
                  g_act=classify_activation(g_activation), g_act_final=classify_activation(g_final_activation),
                  integrator=A.INTEGRATORS[f_integrate_method], readout_pre_step=True, u_transform=ut, u_scale=k)


def make_neural_ode_controller(input_specs, output_specs, control_timestep: float, state_dim: int, key=1,
                               f_integrate_method: str = "RK4", f_use_bias: bool = True, f_time_invariant: bool = True,
                               f_use_dropout: bool = False, f_dropout_rate: float = 0.4, f_width_size: int = 10,
                               f_depth: int = 2, f_activation="relu", f_final_activation="identity",
                               g_use_bias: bool = True, g_time_invariant: bool = True, g_use_dropout: bool = False,
                               g_dropout_rate: float = 0.4, g_width_size: int = 10, g_depth: int = 0,
                               g_activation="relu", g_final_activation="identity", g_feedthrough_u: bool = False,
                               device="cuda") -> NeuralOdeController:
    """same keyword surface as cc/examples/neural_ode_controller.py:16-40"""
    if f_use_dropout or g_use_dropout:
        raise NotImplementedError
    return _build(NeuralOdeController, input_specs, output_specs, control_timestep, state_dim, key, device,
                  f_width=f_width_size, f_depth=f_depth, g_width=g_width_size, g_depth=g_depth,
                  f_act=classify_activation(f_activation), f_act_final=classify_activation(f_final_activation),
                  g_act=classify_activation(g_activation), g_act_final=classify_activation(g_final_activation),
                  f_use_bias=f_use_bias, g_use_bias=g_use_bias, integrator=A.INTEGRATORS[f_integrate_method],
                  f_time_variant=not f_time_invariant, g_time_variant=not g_time_invariant,
                  g_feedthrough_u=g_feedthrough_u)


def make_neural_ode_controller_compact(input_specs, output_specs, control_timestep: float, state_dim: int, key=1,
                                       f_integrate_method: str = "RK4", f_width_size: int = 10, f_depth: int = 2,
                                       f_activation="relu", f_final_activation="identity", g_width_size: int = 10,
                                       g_depth: int = 0, g_activation="relu", g_final_activation="identity",
                                       superposition: bool = False, superposition_trainable: bool = False,
                                       device="cuda") -> NeuralOdeController:
    """cc/examples/neural_ode_controller_compact_example.py:33-52.  With the default zero secondary controller
    (DummyController, :171-181) the output is sigmoid(alpha=0) * g(x) = 0.5 * g(x), readout before the update."""
    if superposition:
        # secondary PID controller (:150-163): host-level composition into ONE linear node, exact for a linear neural part
        if f_depth != 0 or classify_activation(f_final_activation) != A.ACT_IDENTITY or g_depth != 0 or \
                classify_activation(g_final_activation) != A.ACT_IDENTITY or f_integrate_method != "RK4":
            raise NotImplementedError("PID superposition is accelerated for the linear compact controller only (f_depth = g_depth = 0, "
                                      "identity final activations, RK4): see chain_control_b200/examples/superposition.py")
        from .superposition import make_superposed_compact_controller

        I, _ = struct_of_specs(input_specs)
        O, _ = struct_of_specs(output_specs)
        if I != 2 or O != 1:
            raise NotImplementedError("PID superposition: scalar reference / observation / control only")
        return make_superposed_compact_controller(1, 1, control_timestep, state_dim, key, superposition_trainable, device)
    if superposition_trainable:
        raise NotImplementedError("superposition_trainable without superposition has no secondary controller to weigh")
    return _build(NeuralOdeController, input_specs, output_specs, control_timestep, state_dim, key, device,
                  f_width=f_width_size, f_depth=f_depth, g_width=g_width_size, g_depth=g_depth,
                  f_act=classify_activation(f_activation), f_act_final=classify_activation(f_final_activation),
                  g_act=classify_activation(g_activation), g_act_final=classify_activation(g_final_activation),
                  integrator=A.INTEGRATORS[f_integrate_method], readout_pre_step=True, out_scale=0.5)
