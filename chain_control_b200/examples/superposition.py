"""Compact neural-ODE controller in SUPERPOSITION with a PID secondary
(cc/examples/neural_ode_controller_compact_example.py:88-168):

    u = sigmoid(alpha) * g(x_c)  +  (1 - sigmoid(alpha)) * PID([ref ; obs]),     x_c+ = RK4 step of f([x_c ; ref ; obs])

The two parts integrate differently (RK4 vs. the PID's discrete update), so they are not one neural-ODE node -- unless the
neural part is LINEAR (f_depth = 0 with identity activations: the configuration of the reference's notebooks).  Then one RK4
step of  x' = A x + B v + c  is exactly the affine map  x+ = Phi x + Gam (B v + c),  Phi = sum_{n<=4} (hA)^n / n!,
Gam = h sum_{n<=3} (hA)^n / (n+1)!  (nn_lib/integrate.py:6-13 applied to a linear right-hand side), and the superposed
controller IS a single discrete-time linear node on the state [x_c ; last_error ; sum_of_errors] with the readout of the
pre-step state and feed-through -- which the closed-loop kernels already run (forward and BPTT).  `superposed_theta` builds
that node's parameters as a differentiable fp64 torch expression of the reference's trainable leaves, so the kernels'
gradient reaches f, g, alpha and the PID gains through autograd.  A non-linear f raises NotImplementedError.

Flat parameter order (flatten_module, module_utils.py:20-24; field order f, g, alpha, secondary of :88-95; the PID's params
dict flattens with sorted keys): [f.W, f.b, g.W, g.b, (alpha), d_gain, i_gain, p_gain].
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Optional

import torch

from .. import arch as A
from ..core.abstract import AbstractModel
from .neural_ode import NeuralOdeController, _ClosedLoopUnroll, init_flat_params, _key_to_seed


def superposed_theta(flat: torch.Tensor, alpha_fixed, S: int, R: int, O: int, dt: float, alpha_trainable: bool) -> torch.Tensor:
    """reference-order flat parameters -> kernel-order parameters of the combined node (state S + 2, input R + O = 2, one
    output, no-integrate, pre-step readout with feed-through, with bias).  Computed in fp64."""
    assert R == 1 and O == 1, "the PID secondary is defined for a scalar reference / observation (pid_controller.py:9-14)"
    p = flat.to(torch.float64)
    I = R + O
    o = 0
    Wf = p[o:o + S * (S + I)].view(S, S + I); o += S * (S + I)
    bf = p[o:o + S]; o += S
    Wg = p[o:o + S].view(1, S); o += S
    bg = p[o:o + 1]; o += 1
    if alpha_trainable:
        alpha = p[o]; o += 1
    else:
        alpha = torch.as_tensor(alpha_fixed, dtype=torch.float64, device=p.device)
    kd_raw, ki, kp = p[o], p[o + 1], p[o + 2]
    kd = 0.1 * kp + 0.0 * kd_raw  # transform_params of the secondary (:156-160): d = 0.1 p (the d_gain leaf gets a zero gradient)
    h = float(dt)
    Am, Bm = Wf[:, :S], Wf[:, S:]
    hA = h * Am
    eye = torch.eye(S, dtype=torch.float64, device=p.device)
    hA2 = hA @ hA
    hA3 = hA2 @ hA
    Phi = eye + hA + hA2 / 2 + hA3 / 6 + (hA3 @ hA) / 24
    Gam = h * (eye + hA / 2 + hA2 / 6 + hA3 / 24)
    Gv, Gc = Gam @ Bm, Gam @ bf
    a = torch.sigmoid(alpha)
    dc = kp + ki * h + kd / h
    z1 = torch.zeros(1, dtype=torch.float64, device=p.device)
    SS = S + 2
    # f rows: [x_c (S) | last_error | sum_of_errors | ref | obs] -> next state, + bias
    rows, bias = [], []
    for n in range(S):
        rows.append(torch.cat([Phi[n], z1, z1, Gv[n]]))
        bias.append(Gc[n:n + 1])
    zS = torch.zeros(S, dtype=torch.float64, device=p.device)
    one = torch.ones(1, dtype=torch.float64, device=p.device)
    rows.append(torch.cat([zS, z1, z1, one, -one]))          # last_error' = ref - obs
    rows.append(torch.cat([zS, z1, one, h * one, -h * one]))  # sum' = sum + dt (ref - obs)
    bias += [z1, z1]
    Wfk = torch.stack(rows)
    assert Wfk.shape == (SS, SS + I)
    # g row: u = a (Wg x_c + bg) + (1 - a) (-kd/h last + ki sum + dc ref - dc obs)
    grow = torch.cat([a * Wg[0], ((1 - a) * (-kd / h)).reshape(1), ((1 - a) * ki).reshape(1), ((1 - a) * dc).reshape(1),
                      (-(1 - a) * dc).reshape(1)])
    gb = (a * bg).reshape(1)
    return torch.cat([Wfk.reshape(-1), torch.cat(bias), grow, gb])


class SuperposedCompactController(NeuralOdeController):
    def __init__(self, dims, flat: torch.Tensor, alpha_fixed: float, alpha_trainable: bool, dt: float, state=None):
        S, R, O = dims
        self.dims, self.flat, self.alpha_fixed, self.alpha_trainable, self.dt = dims, flat, float(alpha_fixed), bool(alpha_trainable), float(dt)
        self.spec = A.make_node_spec(S + 2, R + O, 1, dt, f_depth=0, g_depth=0, integrator=A.INT_NONE, readout_pre_step=True,
                                     g_feedthrough_u=True)
        self.init_state = (torch.zeros(S + 2, device=flat.device), 0.0)
        self.state = state if state is not None else self.init_state
        self._x0_nonzero = False
        self._out_struct = None
        self._post = lambda y: y

    @property
    def theta(self) -> torch.Tensor:
        S, R, O = self.dims
        return superposed_theta(self.flat, self.alpha_fixed, S, R, O, self.dt, self.alpha_trainable).to(torch.float32)

    def flat_params(self) -> torch.Tensor:
        return self.flat

    def with_params(self, flat: torch.Tensor):
        return SuperposedCompactController(self.dims, flat, self.alpha_fixed, self.alpha_trainable, self.dt, self.state)

    def grad_filter_spec(self):
        return OrderedDict(f=True, g=True, alpha=self.alpha_trainable,
                           secondary=OrderedDict(state=False, params=dict(d_gain=True, i_gain=True, p_gain=True)),
                           state=False, init_state=False)

    def reset(self):
        return SuperposedCompactController(self.dims, self.flat, self.alpha_fixed, self.alpha_trainable, self.dt, None)

    def step(self, x):
        from ..utils import batch_concat

        state, y = self._step_flat(batch_concat(x, 0))
        return SuperposedCompactController(self.dims, self.flat, self.alpha_fixed, self.alpha_trainable, self.dt, state), y

    def unroll(self, model: AbstractModel, merge_x_y: Callable, noise_scale: Optional[float] = None):
        return _ClosedLoopUnroll(self, model, merge_x_y, noise_scale)


def make_superposed_compact_controller(R: int, O: int, control_timestep: float, state_dim: int, key, superposition_trainable: bool,
                                       device) -> SuperposedCompactController:
    nn = A.make_node_spec(state_dim, R + O, 1, control_timestep, f_depth=0, g_depth=0, readout_pre_step=True)
    g = torch.Generator().manual_seed(_key_to_seed(key) + 17)
    gains = torch.rand(3, generator=g)  # jax.random.uniform(key, (3,), maxval=1.0) (:151); distribution-equivalent draw
    parts = [init_flat_params(nn, key)]
    if superposition_trainable:
        parts.append(torch.zeros(1))    # alpha0 = 0 (:64)
    parts.append(gains)
    flat = torch.cat(parts).to(device=device, dtype=torch.float32)
    return SuperposedCompactController((state_dim, R, O), flat, 0.0, superposition_trainable, control_timestep)
