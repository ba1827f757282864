"""Host-side mirror of cc/core/abstract.py: AbstractModel (:39-74) and AbstractController (:82-127).

Same method contract (reset / step / unroll / y0 / grad_filter_spec; functional updates: `step` returns
a NEW module).  What differs is underneath: `unroll` and its batched form run the fused CUDA rollout
kernels through the C ABI instead of jax.lax.scan, and the batched forms are differentiable through
torch.autograd (custom VJP = the BPTT kernels)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable, Optional


class AbstractModel(ABC):
    @abstractmethod
    def step(self, x):
        """-> (new model, y)"""

    @abstractmethod
    def grad_filter_spec(self):
        """which leaves are trained (f and g; never state / init_state)"""

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def y0(self):
        pass

    @abstractmethod
    def unroll(self, time_series_of_x, include_y0: bool = True):
        """abstract.py:54-70: reset, scan step over time, optionally prepend y0 -> length T+1"""

    @abstractmethod
    def unroll_batched(self, inputs, include_y0: bool = True):
        """the vmapped form the trainers call: eqx.filter_vmap(model.unroll)(inputs), step_fn.py:124,136"""


class AbstractController(ABC):
    @abstractmethod
    def step(self, x):
        pass

    @abstractmethod
    def grad_filter_spec(self):
        pass

    @abstractmethod
    def reset(self):
        pass

    @abstractmethod
    def unroll(self, model: AbstractModel, merge_x_y: Callable, noise_scale: Optional[float] = None) -> Callable:
        """abstract.py:97-127: returns unroll_closed_loop(time_series_of_refs, key) -> time series of y"""


def filter_vmap(fn):
    """eqx.filter_vmap for the two call shapes of the hot path (step_fn.py:124,136,234-236,257-259):
    `filter_vmap(model.unroll)(inputs)` and `filter_vmap(controller.unroll(model, merge_x_y, noise))(refss, keys)`.
    Anything else has no batched kernel and raises."""
    owner = getattr(fn, "__self__", None)
    if owner is not None and isinstance(owner, AbstractModel) and getattr(fn, "__name__", "") == "unroll":
        return owner.unroll_batched
    if hasattr(fn, "batched"):
        return fn.batched
    raise NotImplementedError("filter_vmap is only defined for model.unroll and controller.unroll(...): "
                              "chain_control_b200 accelerates the batched unroll, nothing else")
