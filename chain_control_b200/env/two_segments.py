"""`make_env("two_segments_v1" | "_v2" | "_v3" | "two_segments")` of cc/env/make_env.py:50-78 + cc/env/sample_envs.py:13-60,
as a batch of independent environments stepped by one CUDA kernel (csrc/cc_env.cuh) instead of dm_control / MuJoCo.

Same observable behaviour as the reference's `SinglePrecisionWrapper(control.Environment(physics, SegmentTask, time_limit,
control_timestep))` for these ids: `reset()` / `step(action)` return the float32 observation `xpos_of_segment_end`
(cc/env/envs/two_segments.py:119-136,260-263), an episode has `time_limit / control_timestep` steps, actions pass through
arctan (:134-136).  What differs: the leading axis is a batch of environments (n_envs), the tensors live on the GPU, and
`rollout(actions[B, T])` advances a whole episode in one launch (what `collect` needs).  There is no CPU path: without the
CUDA library `make_env` raises."""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .. import _lib

PHYSICS_TIMESTEP = 1e-3  # two_segments.xml:6


@dataclass
class JointParams:  # cc/env/envs/two_segments.py:37-41
    damping: float = 0.0
    springref: float = 0.0
    stiffness: float = 0.0


# cc/env/sample_envs.py:13-47: (slider, hinge) joint parameters per id
ENV_PARAMS = {
    "two_segments_v1": (JointParams(damping=1e-3), JointParams(damping=1e-1, springref=0, stiffness=10)),
    "two_segments_v2": (JointParams(damping=1e-3), JointParams(damping=3e-2, springref=0, stiffness=3)),
    "two_segments_v3": (JointParams(damping=1, stiffness=1), JointParams(damping=3e-2, springref=0, stiffness=2)),
    "two_segments": (JointParams(damping=1e-3), JointParams(damping=3e-3, stiffness=3)),
}


def _param_row(slider: JointParams, hinge: JointParams):
    if slider.springref != 0 or hinge.springref != 0:
        raise NotImplementedError("springref != 0 is not built into the two-segment kernel (every shipped id uses 0)")
    return [slider.damping, slider.stiffness, hinge.damping, hinge.stiffness]


class BatchedTwoSegmentsEnv:
    """n_envs copies of one two-segment chain (or one parameter set per environment)."""

    obs_key = "xpos_of_segment_end"

    def __init__(self, params, n_envs: int = 1, time_limit: float = 10.0, control_timestep: float = 0.01, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CcError("the two-segment data source runs on the GPU only: chain_control_b200 has no CPU path")
        _lib.lib()  # fails loudly when the extension is missing
        p = np.asarray(params, np.float64).reshape(-1, 4)
        if p.shape[0] not in (1, n_envs):
            raise ValueError(f"{p.shape[0]} parameter sets for {n_envs} environments")
        self._params = torch.as_tensor(p, device=self.device)
        self.n_envs = int(n_envs)
        self._control_timestep = float(control_timestep)
        n_sub = control_timestep / PHYSICS_TIMESTEP
        if abs(n_sub - round(n_sub)) > 1e-6 or round(n_sub) < 1:  # dm_control's compute_n_steps raises for non-multiples
            raise ValueError(f"control_timestep {control_timestep} is not a multiple of the physics timestep {PHYSICS_TIMESTEP}")
        self._n_sub = int(round(n_sub))
        self._step_limit = int(round(time_limit / control_timestep))  # control.Environment: time_limit / control_timestep
        self._state = torch.zeros(self.n_envs, 6, dtype=torch.float64, device=self.device)
        self._step_count = 0
        self._reset_next = True

    # ---- the pieces of the dm_env interface the reference's code reads ------------------------------------------------
    def control_timestep(self) -> float:
        return self._control_timestep

    def time_limit(self) -> float:  # cc/utils/utils.py:13-14
        return self._control_timestep * self._step_limit

    def timestep_array(self) -> np.ndarray:  # cc/utils/utils.py:17-18
        return np.arange(self.time_limit(), step=self._control_timestep)

    def action_spec(self):
        return OrderedDict(shape=(1,), dtype=np.float32)

    def observation_spec(self):
        return OrderedDict([(self.obs_key, OrderedDict(shape=(1,), dtype=np.float32))])

    def _call(self, actions: Optional[torch.Tensor], T: int, state: Optional[torch.Tensor]):
        obs = torch.empty(self.n_envs, T + 1, dtype=torch.float32, device=self.device)
        a32 = a64 = None
        if T:
            if actions.dtype == torch.float64:
                a64 = actions.contiguous()
            else:
                a32 = actions.to(torch.float32).contiguous()
        L = _lib.lib()
        with torch.cuda.device(self.device):
            rc = L._dll.cc_env_two_segments_rollout(
                self._params.data_ptr(), self._params.shape[0], None if a32 is None else a32.data_ptr(),
                None if a64 is None else a64.data_ptr(), obs.data_ptr(), None if state is None else state.data_ptr(),
                self.n_envs, T, self._n_sub, C.c_double(PHYSICS_TIMESTEP), torch.cuda.current_stream().cuda_stream)
        L.check(rc, "cc_env_two_segments_rollout")
        return obs

    def _actions(self, a):
        """float64 actions stay float64 (arctan in fp64, like Python-float actions in the reference), everything else is float32"""
        if isinstance(a, (float, int)):  # numpy semantics of the reference: np.arctan(<Python float>) is a float64 operation
            a = torch.tensor(float(a), dtype=torch.float64)
        a = torch.as_tensor(a, device=self.device)
        return a if a.dtype in (torch.float32, torch.float64) else a.to(torch.float32)

    def reset(self):
        """-> observation [n_envs, 1] of the initial state (qpos = qvel = 0)"""
        self._state.zero_()
        self._step_count = 0
        self._reset_next = False
        return OrderedDict([(self.obs_key, self._call(None, 0, self._state)[:, :1])])

    def step(self, action):
        """action [n_envs] or [n_envs, 1] (or a scalar for n_envs = 1) -> (observation [n_envs, 1], last: bool).
        Like dm_env, the first call after construction / after an episode's last step resets and ignores the action."""
        if self._reset_next:
            return self.reset(), False
        a = self._actions(action).reshape(self.n_envs, 1)
        obs = self._call(a, 1, self._state)[:, 1:]
        self._step_count += 1
        last = self._step_count >= self._step_limit
        self._reset_next = last
        return OrderedDict([(self.obs_key, obs)]), last

    def rollout(self, actions):
        """a whole episode from reset in ONE launch: actions [n_envs, T(, 1)] -> observations [n_envs, T + 1, 1]
        (the reset observation first), i.e. what EnvironmentLoop.run_episode records (cc/env/collect/collect.py:78-98)."""
        a = self._actions(actions)
        if a.ndim == 3:
            a = a[..., 0]
        if a.shape[0] != self.n_envs:
            raise ValueError(f"actions for {a.shape[0]} environments, have {self.n_envs}")
        if a.shape[1] > self._step_limit:
            raise ValueError(f"{a.shape[1]} actions exceed the episode length {self._step_limit}")
        return OrderedDict([(self.obs_key, self._call(a, a.shape[1], None)[..., None])])


def make_env(id: str, time_limit: float = 10.0, control_timestep: float = 0.01, n_envs: int = 1, device="cuda",
             physics_kwargs: Optional[dict] = None, **kwargs) -> BatchedTwoSegmentsEnv:
    """cc/env/make_env.py:50-78.  `physics_kwargs` = {damping, stiffness} of the hinges for id "two_segments"
    (cc/env/sample_envs.py:36-45).  Other ids (muscle, rover, ackermann: MuJoCo models outside this path) raise."""
    if id not in ENV_PARAMS:
        raise Exception(f"Unknown environment id {id}, available ids are: {ENV_PARAMS.keys()}")
    slider, hinge = ENV_PARAMS[id]
    if physics_kwargs:
        if id != "two_segments":
            raise TypeError(f"{id} takes no physics_kwargs")
        hinge = JointParams(damping=physics_kwargs.get("damping", 3e-3), stiffness=physics_kwargs.get("stiffness", 3))
    return BatchedTwoSegmentsEnv(_param_row(slider, hinge), n_envs, time_limit, control_timestep, device)
