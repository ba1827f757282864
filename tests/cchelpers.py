"""shared helpers of the parity tests."""
import numpy as np

from chain_control_b200 import arch as A


def conv(s):
    """oracle NodeSpec -> product NodeSpec (same field names)."""
    return A.NodeSpec(s.state_dim, s.input_dim, s.output_dim,
                      A.MlpSpec(list(s.f.sizes), s.f.act, s.f.act_final, s.f.use_bias),
                      A.MlpSpec(list(s.g.sizes), s.g.act, s.g.act_final, s.g.use_bias),
                      s.dt, s.integrator, s.f_time_variant, s.g_time_variant, s.g_feedthrough_u,
                      s.readout_pre_step, s.u_transform, s.u_scale, s.out_scale)


def normwise(a, ref):
    """per-trajectory max-abs error / max-abs value (the states/outputs tolerance of BASELINE.json)."""
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    ax = tuple(range(1, ref.ndim))
    return float((np.abs(a - ref).max(axis=ax) / np.maximum(np.abs(ref).max(axis=ax), 1e-30)).max())


def l2rel(a, ref):
    """L2-normwise error (the gradient tolerance of BASELINE.json)."""
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30))


TOL_STATE = 1e-5  # BASELINE.json north_star: states within rel 1e-5 (normwise, fp32)
TOL_GRAD = 1e-4   # gradients within rel 1e-4 (L2-normwise, fp32)
