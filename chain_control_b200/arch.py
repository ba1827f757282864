"""Architecture descriptors of the neural-ODE modules and their C-ABI mirror (include/cc_b200.h).

The descriptors carry exactly the information the reference's factories close over:
``make_neural_ode_model`` (cc/examples/neural_ode_model.py:16-95), ``make_neural_ode_controller``
(cc/examples/neural_ode_controller.py:16-97) and the compact variants.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List

CC_MAX_LAYERS = 4

ACT_IDENTITY, ACT_RELU, ACT_TANH = 0, 1, 2
UT_IDENTITY, UT_ARCTAN, UT_KTANH = 0, 1, 2
INT_RK4, INT_EE, INT_NONE = 0, 1, 2

INTEGRATORS = {"RK4": INT_RK4, "EE": INT_EE, "no-integrate": INT_NONE}  # nn_lib/integrate.py:16-28


class CcMlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("sizes", C.c_int32 * (CC_MAX_LAYERS + 1)), ("act", C.c_int32),
                ("act_final", C.c_int32), ("use_bias", C.c_int32)]


class CcModule(C.Structure):
    _fields_ = [("state_dim", C.c_int32), ("input_dim", C.c_int32), ("output_dim", C.c_int32),
                ("f", CcMlp), ("g", CcMlp), ("integrator", C.c_int32), ("f_time_variant", C.c_int32),
                ("g_time_variant", C.c_int32), ("g_feedthrough_u", C.c_int32),
                ("readout_pre_step", C.c_int32), ("u_transform", C.c_int32), ("u_scale", C.c_float),
                ("out_scale", C.c_float), ("dt", C.c_float), ("t0", C.c_float)]


@dataclass
class MlpSpec:
    """sizes = [in] + depth*[width] + [out]   (nn_lib/mlp_network.py:18)."""

    sizes: List[int]
    act: int = ACT_RELU
    act_final: int = ACT_IDENTITY
    use_bias: bool = True

    @property
    def n_layers(self) -> int:
        return len(self.sizes) - 1

    def n_params(self) -> int:
        return sum(o * i + (o if self.use_bias else 0) for i, o in zip(self.sizes[:-1], self.sizes[1:]))

    def to_c(self) -> CcMlp:
        if self.n_layers > CC_MAX_LAYERS:
            raise NotImplementedError(f"MLP depth {self.n_layers - 1} > {CC_MAX_LAYERS - 1} is not supported")
        s = CcMlp()
        s.n_layers = self.n_layers
        for i, v in enumerate(self.sizes):
            s.sizes[i] = int(v)
        s.act, s.act_final, s.use_bias = int(self.act), int(self.act_final), int(bool(self.use_bias))
        return s


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
    readout_pre_step: bool = False
    u_transform: int = UT_IDENTITY
    u_scale: float = 1.0
    out_scale: float = 1.0
    t0: float = 0.0  # time component of init_state

    def n_params(self) -> int:
        return self.f.n_params() + self.g.n_params()

    @property
    def has_time(self) -> bool:
        return self.f_time_variant or self.g_time_variant

    def to_c(self) -> CcModule:
        c = CcModule()
        c.state_dim, c.input_dim, c.output_dim = int(self.state_dim), int(self.input_dim), int(self.output_dim)
        c.f, c.g = self.f.to_c(), self.g.to_c()
        c.integrator = int(self.integrator)
        c.f_time_variant, c.g_time_variant = int(self.f_time_variant), int(self.g_time_variant)
        c.g_feedthrough_u, c.readout_pre_step = int(self.g_feedthrough_u), int(self.readout_pre_step)
        c.u_transform, c.u_scale = int(self.u_transform), float(self.u_scale)
        c.out_scale, c.dt = float(self.out_scale), float(self.dt)
        c.t0 = float(self.t0)
        return c


def make_node_spec(state_dim, input_dim, output_dim, dt, *, f_width=10, f_depth=2, g_width=10, g_depth=0,
                   f_act=ACT_RELU, f_act_final=ACT_IDENTITY, g_act=ACT_RELU, g_act_final=ACT_IDENTITY,
                   f_use_bias=True, g_use_bias=True, integrator=INT_RK4, f_time_variant=False,
                   g_time_variant=False, g_feedthrough_u=False, readout_pre_step=False,
                   u_transform=UT_IDENTITY, u_scale=1.0, out_scale=1.0) -> NodeSpec:
    """size bookkeeping of neural_ode_model.py:45-59 / neural_ode_controller.py:45-60."""
    f_in = state_dim + input_dim + (1 if f_time_variant else 0)
    g_in = state_dim + (1 if g_time_variant else 0) + (input_dim if g_feedthrough_u else 0)
    f = MlpSpec([f_in] + f_depth * [f_width] + [state_dim], f_act, f_act_final, f_use_bias)
    g = MlpSpec([g_in] + g_depth * [g_width] + [output_dim], g_act, g_act_final, g_use_bias)
    return NodeSpec(state_dim, input_dim, output_dim, f, g, dt, integrator, f_time_variant, g_time_variant,
                    g_feedthrough_u, readout_pre_step, u_transform, u_scale, out_scale)
