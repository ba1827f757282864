"""LinearModel of cc/examples/linear_model.py:13-84 on the CUDA rollout kernels.

    x+ = x + (A x + B u) * dt   (continuous_time, explicit Euler)      |   x+ = A x + B u   (discrete time)
    y  = C x + D u              (readout of the PRE-step state with feed-through; y0 = C x0)

To the kernels this is a depth-0 node: f = Linear([x ; u] -> S) without bias, integrator EE / no-integrate, g = Linear([x ; u] -> O)
without bias, `readout_pre_step`, `g_feedthrough_u` (dropped when D is identically zero and not trained: the model then
qualifies for the blocked-linear tcgen05 family).  The trainable parameters live in ONE flat vector in the reference's
flatten_module order [A, B, C, (D), ] (module_utils.py:20-24, field order of linear_model.py:47-53); `theta` is the
same data in the kernels' order ([A|B] rows then [C|D] rows), built with differentiable torch indexing so that the BPTT
kernels' gradient flows back to the flat vector through autograd.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .. import arch as A
from ..utils import struct_of_specs
from .neural_ode import NeuralOdeModel, _key_to_seed
from ..utils import make_postprocess_fn


class LinearModel(NeuralOdeModel):
    def __init__(self, dims, params: torch.Tensor, D_fixed, x0: torch.Tensor, state, dt, continuous_time, include_D, out_struct):
        S, I, O = dims
        self.dims = dims
        self.params = params          # flat [A (S*S), B (S*I), C (O*S), (D (O*I) if include_D)]
        self.D_fixed = D_fixed        # the non-trained D (include_D = False), else None
        self.dt, self.continuous_time, self.include_D = float(dt), bool(continuous_time), bool(include_D)
        D = self.D
        self._ft = self.include_D or bool(torch.any(D != 0))  # (host sync once, at construction)
        self.spec = A.make_node_spec(S, I, O, dt, f_depth=0, g_depth=0, f_use_bias=False, g_use_bias=False,
                                     integrator=A.INT_EE if continuous_time else A.INT_NONE, readout_pre_step=True,
                                     g_feedthrough_u=self._ft)
        self.init_state = (x0, 0.0)
        self.state = state if state is not None else self.init_state
        self._x0_nonzero = bool(torch.any(x0 != 0))
        self._out_struct = out_struct
        self._post = make_postprocess_fn(out_struct)

    # ---- the reference's fields as views of the flat vector --------------------------------------
    def _split(self):
        S, I, O = self.dims
        p = self.params
        a, b, c = S * S, S * S + S * I, S * S + S * I + O * S
        return p[:a].view(S, S), p[a:b].view(S, I), p[b:c].view(O, S), (p[c:c + O * I].view(O, I) if self.include_D else self.D_fixed)

    A = property(lambda self: self._split()[0])
    B = property(lambda self: self._split()[1])
    C = property(lambda self: self._split()[2])
    D = property(lambda self: self._split()[3])
    x0 = property(lambda self: self.init_state[0])

    @property
    def theta(self) -> torch.Tensor:
        Am, Bm, Cm, Dm = self._split()
        f = torch.cat([Am, Bm], 1).reshape(-1)
        g = torch.cat([Cm, Dm.to(Cm.device)], 1).reshape(-1) if self._ft else Cm.reshape(-1)
        return torch.cat([f, g])

    def flat_params(self) -> torch.Tensor:
        return self.params

    def with_params(self, params: torch.Tensor):
        return LinearModel(self.dims, params, self.D_fixed, self.init_state[0], self.state, self.dt, self.continuous_time,
                           self.include_D, self._out_struct)

    def replace(self, A=None, B=None, C=None, D=None):
        """eqx.tree_at(lambda m: (m.A, m.B, m.C, m.D), m, (A, B, C, D)) of the reference's CLI (train_controller.py:40-58)"""
        Am, Bm, Cm, Dm = self._split()
        dev = self.params.device
        cv = lambda new, old: old if new is None else torch.as_tensor(np.asarray(new), dtype=torch.float32, device=dev).reshape(old.shape)
        Am, Bm, Cm, Dm = cv(A, Am), cv(B, Bm), cv(C, Cm), cv(D, Dm)
        parts = [Am.reshape(-1), Bm.reshape(-1), Cm.reshape(-1)] + ([Dm.reshape(-1)] if self.include_D else [])
        return LinearModel(self.dims, torch.cat(parts), None if self.include_D else Dm, self.init_state[0], None, self.dt,
                           self.continuous_time, self.include_D, self._out_struct)

    def frozen(self):
        """the same dynamics with nothing trained (the model inside a closed loop, step_fn.py:249-252): D leaves the
        parameter vector, and an all-zero D leaves the kernels' readout"""
        if not self.include_D:
            return self
        Am, Bm, Cm, Dm = (t.detach() for t in self._split())
        return LinearModel(self.dims, torch.cat([Am.reshape(-1), Bm.reshape(-1), Cm.reshape(-1)]), Dm, self.init_state[0],
                           self.state, self.dt, self.continuous_time, False, self._out_struct)

    def grad_filter_spec(self):
        # linear_model.py:69-76: x is never optimised, x0 / D on request
        return OrderedDict(A=True, B=True, C=True, D=self.include_D, x=False, x0=False)

    def reset(self):
        return LinearModel(self.dims, self.params, self.D_fixed, self.init_state[0], None, self.dt, self.continuous_time,
                           self.include_D, self._out_struct)

    def step(self, u):
        from ..utils import batch_concat

        state, y = self._step_flat(batch_concat(u, 0))
        m = LinearModel(self.dims, self.params, self.D_fixed, self.init_state[0], state, self.dt, self.continuous_time,
                        self.include_D, self._out_struct)
        return m, self._post(y)


def make_linear_model(input_specs, output_specs, control_timestep: float, state_dim: int, continuous_time: bool = True,
                      include_D: bool = True, seed: int = 1, randomize_x0: bool = False, include_x0: bool = False,
                      device="cuda") -> LinearModel:
    """same keyword surface as cc/examples/linear_model.py:13-23; A0, B0, C0, D0 (and x0) ~ N(0, 1) (the draw is
    distribution-equivalent: jax's threefry stream cannot be reproduced without jax)"""
    if include_x0:
        raise NotImplementedError("a trainable initial state needs the x0 cotangent, which the BPTT kernels do not return")
    I, _ = struct_of_specs(input_specs)
    O, out_struct = struct_of_specs(output_specs)
    S = state_dim
    g = torch.Generator().manual_seed(_key_to_seed(seed))
    n = S * S + S * I + O * S + O * I
    flat = torch.randn(n, generator=g)
    x0 = torch.randn(S, generator=g) if randomize_x0 else torch.zeros(S)
    D = flat[n - O * I:].view(O, I)
    params = (flat if include_D else flat[:n - O * I]).to(device=device, dtype=torch.float32)
    return LinearModel((S, I, O), params, None if include_D else D.to(device), x0.to(device), None, control_timestep,
                       continuous_time, include_D, out_struct)
