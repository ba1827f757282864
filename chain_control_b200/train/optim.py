"""The optax pieces the reference's recipes use, on a flat parameter vector:
adam (bias-corrected, b1=0.9, b2=0.999, eps=1e-8, update -lr*m_hat/(sqrt(v_hat)+eps)),
clip_by_global_norm, clip, chain — e.g. chain(clip_by_global_norm(1.0), adam(1e-3)) (cc/train/test_trainer.py:58,107)."""
from __future__ import annotations

from typing import Callable, NamedTuple

import torch


class GradientTransformation(NamedTuple):
    init: Callable
    update: Callable  # (updates, state, params=None) -> (updates, state)


def adam(learning_rate: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> GradientTransformation:
    def init(params):
        return dict(count=0, mu=torch.zeros_like(params), nu=torch.zeros_like(params))

    def update(g, state, params=None):
        count = state["count"] + 1
        mu = b1 * state["mu"] + (1 - b1) * g
        nu = b2 * state["nu"] + (1 - b2) * g * g
        mu_hat = mu / (1 - b1 ** count)
        nu_hat = nu / (1 - b2 ** count)
        return -learning_rate * mu_hat / (torch.sqrt(nu_hat) + eps), dict(count=count, mu=mu, nu=nu)

    return GradientTransformation(init, update)


def clip_by_global_norm(max_norm: float) -> GradientTransformation:
    def update(g, state, params=None):
        norm = torch.linalg.vector_norm(g)
        scale = torch.where(norm > max_norm, max_norm / norm.clamp_min(1e-38), torch.ones_like(norm))
        return g * scale, state

    return GradientTransformation(lambda params: (), update)


def clip(max_delta: float) -> GradientTransformation:
    return GradientTransformation(lambda params: (), lambda g, state, params=None: (g.clamp(-max_delta, max_delta), state))


def sgd(learning_rate: float) -> GradientTransformation:
    return GradientTransformation(lambda params: (), lambda g, state, params=None: (-learning_rate * g, state))


def chain(*transforms: GradientTransformation) -> GradientTransformation:
    def init(params):
        return tuple(t.init(params) for t in transforms)

    def update(g, state, params=None):
        new = []
        for t, s in zip(transforms, state):
            g, s = t.update(g, s, params)
            new.append(s)
        return g, tuple(new)

    return GradientTransformation(init, update)


def apply_updates(params, updates):
    return params + updates
