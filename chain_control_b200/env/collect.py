"""Input / reference generators and data collection of cc/env/collect (source.py, circus.py, collect.py), batched.

The reference records one episode at a time through acme's EnvironmentLoop (collect.py:78-98); here all episodes of a call
advance together: feed-forward collection is ONE launch of the two-segment kernel over [n_seeds, T], closed-loop
collection steps the batched controller kernel and the batched environment kernel alternately (two launches per control
step for ALL references, instead of a Python episode per reference)."""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Optional, Sequence, Union

import numpy as np
import torch

from .. import arch as A
from .. import ops
from .two_segments import BatchedTwoSegmentsEnv


@dataclass
class ReplaySample:  # cc/env/buffer/replay_element_sample.py:30-51
    obs: OrderedDict
    action: np.ndarray
    rew: Optional[np.ndarray] = None
    done: Optional[np.ndarray] = None
    extras: dict = field(default_factory=dict)

    @property
    def bs(self):
        assert self.action.ndim == 3
        return self.action.shape[0]

    @property
    def n_timesteps(self):
        assert self.action.ndim == 3
        return self.action.shape[1]


class ObservationReferenceSource:  # cc/env/collect/source.py:18-42
    def __init__(self, yss: OrderedDict, ts: Optional[np.ndarray] = None, uss=None, i_actor: int = 0):
        self._i_actor = i_actor
        self._ts = None if ts is None else np.asarray(ts)
        self._yss = OrderedDict((k, _np(v)) for k, v in yss.items())
        self._uss = None if uss is None else _np(uss)

    def get_references(self) -> OrderedDict:
        return OrderedDict((k, np.atleast_3d(v)) for k, v in self._yss.items())

    def get_references_for_optimisation(self) -> OrderedDict:
        return OrderedDict((k, torch.as_tensor(v)) for k, v in self.get_references().items())

    def get_reference_actor(self) -> OrderedDict:
        return OrderedDict((k, v[self._i_actor]) for k, v in self.get_references_for_optimisation().items())

    def change_reference_of_actor(self, i_actor: int) -> None:
        self._i_actor = i_actor


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def draw_u_from_gaussian_process(ts, kernel=None, seed=1, scale_u_by: float = 0.25):
    """cc/env/collect/source.py:44-54: one draw of a GP with kernel 0.15 * RBF(0.75), standardised, times scale_u_by"""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF

    kernel = 0.15 * RBF(0.75) if kernel is None else kernel
    ts = np.asarray(ts)[:, np.newaxis]
    us = GaussianProcessRegressor(kernel=kernel).sample_y(ts, random_state=seed)
    us = (us - np.mean(us)) / np.std(us)
    us *= scale_u_by
    return us.astype(np.float32)


def draw_u_from_cosines(ts, seed):
    """cc/env/collect/source.py:57-61 (float64 out, like the reference)"""
    freq = seed
    ts = np.asarray(ts) / np.asarray(ts)[-1]
    omega = 2 * np.pi * freq
    return np.cos(omega * ts)[:, None] * np.sqrt(freq)


def _sibling(env: BatchedTwoSegmentsEnv, n: int) -> BatchedTwoSegmentsEnv:
    if env._params.shape[0] != 1:
        raise ValueError("collection replicates ONE parameter set over the episodes of the call")
    return BatchedTwoSegmentsEnv(env._params.cpu().numpy(), n, env.time_limit(), env.control_timestep(), env.device)


def sample_feedforward_collect_and_make_source(env: BatchedTwoSegmentsEnv, draw_fn=draw_u_from_gaussian_process,
                                               seeds: Sequence[Union[int, float]] = (0,), global_scale_u: float = 1.0):
    """cc/env/collect/collect.py:101-125: one feed-forward episode per seed -> (source, sample, {})"""
    assert len(seeds) > 0
    ts = env.timestep_array()
    # (the reference converts every draw to a float32 jax array before it reaches the environment: to_jax, x64 disabled)
    us = np.stack([np.asarray(draw_fn(ts, seed=seed), np.float32) * np.float32(global_scale_u) for seed in seeds])
    obs = _sibling(env, len(seeds)).rollout(torch.as_tensor(us, device=env.device))
    sample = ReplaySample(obs=OrderedDict((k, v.cpu().numpy()) for k, v in obs.items()), action=us)
    return ObservationReferenceSource(sample.obs, ts, sample.action), sample, {}


def sample_feedforward_and_collect(env: BatchedTwoSegmentsEnv, seeds_gp: Sequence[int] = (), seeds_cos: Sequence[Union[int, float]] = (),
                                   global_scale_u: float = 1.0) -> ReplaySample:
    """cc/env/collect/collect.py:33-52"""
    samples = []
    if len(seeds_gp) > 0:
        samples.append(sample_feedforward_collect_and_make_source(env, seeds=seeds_gp, global_scale_u=global_scale_u)[1])
    if len(seeds_cos) > 0:
        samples.append(sample_feedforward_collect_and_make_source(env, draw_u_from_cosines, seeds=seeds_cos, global_scale_u=global_scale_u)[1])
    assert len(samples) > 0, "Either `seeds_gp` or `seeds_cos` must be given"
    return ReplaySample(obs=OrderedDict((k, np.concatenate([s.obs[k] for s in samples])) for k in samples[0].obs),
                        action=np.concatenate([s.action for s in samples]))


def random_steps_source(env: BatchedTwoSegmentsEnv, seeds: Sequence[int], min_abs_amplitude: float = 0.0, max_abs_amplitude: float = 3.0):
    """cc/env/collect/circus.py:14-36: one constant reference per seed, sign * U(min, max), numpy's legacy global stream"""
    ts = env.timestep_array()
    yss = []
    for seed in seeds:
        np.random.seed(seed)
        sign = np.random.choice([-1.0, 1.0])
        value = sign * np.random.uniform(min_abs_amplitude, max_abs_amplitude, size=(1, 1))
        yss.append(np.repeat(value, len(ts) + 1, axis=0))
    return ObservationReferenceSource(OrderedDict([(env.obs_key, np.stack(yss))]), ts=ts)


def collect_exhaust_source(env: BatchedTwoSegmentsEnv, source: ObservationReferenceSource, controller) -> ReplaySample:
    """cc/env/collect/collect.py:55-75: the controller's closed loop with the ENVIRONMENT for every reference of the source.
    Per control step (cc/env/wrappers/add_reference_and_reward.py:46-71, collect/actor.py:69-78): the controller sees
    merge_x_y(ref_k, obs_k) and its action drives the environment to obs_{k+1}.  All references advance together."""
    refs = torch.as_tensor(source.get_references()[env.obs_key], dtype=torch.float32, device=env.device)  # [N, T + 1, 1]
    N, T = refs.shape[0], env._step_limit
    benv = _sibling(env, N)
    obs = [benv.reset()[env.obs_key]]
    c = controller.reset()
    spec, theta = c.spec, c.theta.detach()
    x = c.init_state[0].to(env.device).expand(N, -1).contiguous()
    t = np.float32(c.init_state[1])
    actions = []
    for k in range(T):
        s = spec
        if spec.has_time and float(t) != spec.t0:
            s = A.NodeSpec(**{**spec.__dict__})
            s.t0 = float(t)
        inp = torch.cat([refs[:, k], obs[-1]], dim=1)[:, None, :]  # OrderedDict(ref, obs) order (step_fn.py:187-192)
        y, x, _ = ops.rollout_fwd(s, theta, inp.contiguous(), x, want_x_final=True)
        u = y[:, 1, :1]
        actions.append(u)
        o, _ = benv.step(u)
        obs.append(o[env.obs_key])
        t = np.float32(t + np.float32(spec.dt))
    obs_all = torch.stack(obs, 1)
    # default_reward_fn of the reference's wrapper (add_reference_and_reward.py:17-20, 46-61): -mean((ref - obs)^2) per step,
    # none for the first timestep (dm_env convention) -> [N, T]
    rew = -((refs[:, 1:T + 1] - obs_all[:, 1:]) ** 2).mean(dim=2)
    return ReplaySample(obs=OrderedDict([(env.obs_key, obs_all.cpu().numpy())]), action=torch.stack(actions, 1).cpu().numpy(),
                        rew=rew.cpu().numpy(), extras={"ref": refs.cpu().numpy()})
