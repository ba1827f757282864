"""PIDController of cc/examples/pid_controller.py:17-79 on the CUDA closed-loop kernels.

    err = ref - obs;  sum' = sum + dt*err;  control = Kp*err + Ki*sum' + Kd*(err - last)/dt;  last' = err

To the kernels this is a discrete-time linear node on the state z = [last_error, sum_of_errors] with input v = [ref, obs]:
    z+ = F [z ; v]   (no-integrate; F = [[0 0 1 -1], [0 1 dt -dt]], constants)
    u  = G [z ; v]   (readout of the PRE-step state with feed-through; G = [-Kd/dt, Ki, dc, -dc], dc = Kp + Ki dt + Kd/dt)
The trainable parameters are the three gains in the reference's flat order -- jax flattens the params dict with SORTED
keys: (d_gain, i_gain, p_gain) -- and `theta` (the kernels' 12 numbers) is a differentiable torch expression of them, so
`transform_params` may be any torch-traceable callable (the compact controller's secondary uses d = 0.1 p,
neural_ode_controller_compact_example.py:156-168) and the BPTT kernels' gradient reaches the gains through autograd.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Optional

import torch

from .. import arch as A
from ..core.abstract import AbstractModel
from .neural_ode import NeuralOdeController, _ClosedLoopUnroll

_KEYS = ("d_gain", "i_gain", "p_gain")  # sorted: jax's dict flattening order


class PIDController(NeuralOdeController):
    def __init__(self, gains: torch.Tensor, trainable, dt, transform_params, state=None):
        self.gains = gains                    # [3] in _KEYS order
        self.trainable = tuple(bool(t) for t in trainable)
        self.dt = float(dt)
        self.transform_params = transform_params
        self.spec = A.make_node_spec(2, 2, 1, dt, f_depth=0, g_depth=0, f_use_bias=False, g_use_bias=False,
                                     integrator=A.INT_NONE, readout_pre_step=True, g_feedthrough_u=True)
        dev = gains.device
        self.init_state = (torch.zeros(2, device=dev), 0.0)
        self.state = state if state is not None else self.init_state
        self._x0_nonzero = False
        self._out_struct = None
        self._post = lambda flat: flat

    @property
    def params(self):
        return {k: self.gains[i:i + 1] for i, k in enumerate(_KEYS)}

    @property
    def theta(self) -> torch.Tensor:
        p = self.transform_params(self.params)
        kp, ki, kd = p["p_gain"].reshape(()), p["i_gain"].reshape(()), p["d_gain"].reshape(())
        dt = self.dt
        z, one = torch.zeros((), device=self.gains.device), torch.ones((), device=self.gains.device)
        dc = kp + ki * dt + kd / dt
        F = torch.stack([z, z, one, -one, z, one, dt * one, -dt * one])
        G = torch.stack([-kd / dt, ki, dc, -dc])
        return torch.cat([F, G]).to(torch.float32)

    def flat_params(self) -> torch.Tensor:
        idx = [i for i, t in enumerate(self.trainable) if t]
        return self.gains[idx]

    def with_params(self, flat: torch.Tensor):
        idx = [i for i, t in enumerate(self.trainable) if t]
        parts = []
        for i in range(3):
            parts.append(flat[idx.index(i):idx.index(i) + 1] if i in idx else self.gains[i:i + 1].detach())
        return PIDController(torch.cat(parts), self.trainable, self.dt, self.transform_params, self.state)

    def grad_filter_spec(self):
        return OrderedDict(state=False, params=dict(zip(_KEYS, self.trainable)))

    def reset(self):
        return PIDController(self.gains, self.trainable, self.dt, self.transform_params, None)

    def step(self, x):
        from ..utils import batch_concat

        state, y = self._step_flat(batch_concat(x, 0))
        return PIDController(self.gains, self.trainable, self.dt, self.transform_params, state), y

    def unroll(self, model: AbstractModel, merge_x_y: Callable, noise_scale: Optional[float] = None):
        return _ClosedLoopUnroll(self, model, merge_x_y, noise_scale)


def make_pid_controller(p_gain: float, i_gain: float, d_gain: float, control_timestep: float, p_gain_trainable: bool = True,
                        i_gain_trainable: bool = True, d_gain_trainable: bool = True, transform_params=lambda params: params,
                        device="cuda") -> PIDController:
    """same keyword surface as cc/examples/pid_controller.py:17-26 (scalar reference / observation)"""
    gains = torch.tensor([d_gain, i_gain, p_gain], dtype=torch.float32, device=device)
    return PIDController(gains, (d_gain_trainable, i_gain_trainable, p_gain_trainable), control_timestep, transform_params)
