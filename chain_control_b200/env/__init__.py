"""Host-side mirror of the part of cc/env a train step's data comes from: the `two_segments*`
environments as ONE batched CUDA rollout (cc_env_two_segments_rollout) instead of dm_control / MuJoCo episodes, plus the
input / reference generators of cc/env/collect."""
from .collect import (ObservationReferenceSource, ReplaySample, collect_exhaust_source, draw_u_from_cosines,  # noqa: F401
                      draw_u_from_gaussian_process, random_steps_source, sample_feedforward_and_collect,
                      sample_feedforward_collect_and_make_source)
from .two_segments import ENV_PARAMS, BatchedTwoSegmentsEnv, JointParams, make_env  # noqa: F401
