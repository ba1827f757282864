"""NumPy fp64 restatement of the reference's two-segment-chain data source (TEST INFRASTRUCTURE, like np_oracle.py).

What it follows (all relative to /root/reference):
  * the MJCF model          cc/env/envs/two_segments.xml:6-14 (timestep 1 ms, integrator RK4, contact and gravity disabled,
                            pole = capsule fromto 0 0 0 0 0 1, radius 0.045, mass 0.1)
  * the generated bodies    cc/env/envs/two_segments.py:61-95 (cart: slide joint along x, box of mass 1; pole 1 hangs at
                            (0, 0, -0.1) turned by euler 0 180 0, hinge about y; pole 2 at (0, 0, 1.1) in pole 1, hinge about y;
                            segment_end at (0, 0, 1.0) in pole 2), motor with gear 5 on the slider :98-104
  * the task                cc/env/envs/two_segments.py:119-136, 242-263 (ctrl = arctan(action); observation = world x of
                            segment_end), cc/env/make_env.py:17-46 (control_timestep 0.01 = 10 physics steps; float32 observations)
  * the parameter sets      cc/env/sample_envs.py:13-47 (two_segments_v1/v2/v3 and the `two_segments` default)

The arithmetic lives in a third-party dependency that is absent from /root/reference and from this image: MuJoCo (through
dm_control; unpinned in the reference's setup.py).  Restated here from its published algorithm for exactly this model:
  * joint-space dynamics  M(q) qacc = qfrc_passive + qfrc_actuator - qfrc_bias,  qfrc_passive = -stiffness (q - springref)
    - damping qvel,  qfrc_actuator = gear * ctrl, no gravity / contacts / armature / active limits;
  * body inertias from the geoms: the box only translates (mass 1); a capsule of given mass splits it by volume into a
    cylinder (pi r^2 H) and two half spheres (4/3 pi r^3), inertia about the centre perpendicular to the axis
    m_c (3 r^2 + H^2) / 12 + m_s (2 r^2 / 5 + H (3 r + 2 H) / 8);
  * integrator RK4: the classical tableau on (qpos, qvel) with ctrl held, h = 1 ms.
PARITY PIN: the reference's own golden vectors cc/env/test_make_env.py:146-335 (40 observations of v1, v2, v3 under
DYNAMIC_INPUTS) -- tests/test_two_segments.py requires this file to reproduce them (float32 observations).
"""
from __future__ import annotations

import numpy as np

# (slider damping, slider stiffness, hinge damping, hinge stiffness); springref = 0 everywhere (sample_envs.py:13-47)
ENV_PARAMS = {
    "two_segments_v1": (1e-3, 0.0, 1e-1, 10.0),
    "two_segments_v2": (1e-3, 0.0, 3e-2, 3.0),
    "two_segments_v3": (1.0, 1.0, 3e-2, 2.0),
    "two_segments": (1e-3, 0.0, 3e-3, 3.0),
}

M_CART, M_POLE, R_POLE, H_POLE = 1.0, 0.1, 0.045, 1.0
L1, L2_END, LC = 1.1, 1.0, 0.5  # pole-2 offset along pole 1, segment_end offset along pole 2, centre of mass of a pole
GEAR = 5.0
PHYSICS_DT = 1e-3


def capsule_inertia_perp(mass=M_POLE, r=R_POLE, H=H_POLE):
    v_c, v_s = np.pi * r * r * H, 4.0 / 3.0 * np.pi * r ** 3
    m_s = mass * v_s / (v_c + v_s)
    m_c = mass - m_s
    return m_c * (3 * r * r + H * H) / 12 + m_s * (2 * r * r / 5 + H * (3 * r + 2 * H) / 8)


def qacc(q, v, ctrl, params):
    """q, v: [..., 3] = (slider, hinge 1, hinge 2); ctrl [...]; returns the joint accelerations"""
    ds, ks, dh, kh = params
    I = capsule_inertia_perp()
    p1 = np.pi + q[..., 1]  # absolute pole angles about y (pole 1 hangs: euler 0 180 0)
    p2 = p1 + q[..., 2]
    w1, w2 = v[..., 1], v[..., 1] + v[..., 2]
    a = M_POLE * LC + M_POLE * L1
    b = M_POLE * LC
    c = M_POLE * L1 * LC
    # mass matrix in (x, p1, p2)
    Ma = np.zeros(q.shape[:-1] + (3, 3))
    Ma[..., 0, 0] = M_CART + 2 * M_POLE
    Ma[..., 0, 1] = Ma[..., 1, 0] = a * np.cos(p1)
    Ma[..., 0, 2] = Ma[..., 2, 0] = b * np.cos(p2)
    Ma[..., 1, 1] = M_POLE * LC * LC + I + M_POLE * L1 * L1
    Ma[..., 1, 2] = Ma[..., 2, 1] = c * np.cos(p1 - p2)
    Ma[..., 2, 2] = M_POLE * LC * LC + I
    s12 = np.sin(p1 - p2)
    ca = np.stack([-a * np.sin(p1) * w1 * w1 - b * np.sin(p2) * w2 * w2, c * s12 * w2 * w2, -c * s12 * w1 * w1], -1)
    J = np.array([[1.0, 0, 0], [0, 1, 0], [0, 1, 1]])  # (x, p1, p2)' = J (x, th1, th2)'
    M = J.T @ Ma @ J
    bias = ca @ J
    tau = np.stack([GEAR * ctrl - ds * v[..., 0] - ks * q[..., 0], -dh * v[..., 1] - kh * q[..., 1], -dh * v[..., 2] - kh * q[..., 2]], -1)
    return np.linalg.solve(M, (tau - bias)[..., None])[..., 0]


def rk4_step(q, v, ctrl, params, h=PHYSICS_DT):
    a1 = qacc(q, v, ctrl, params)
    q2, v2 = q + 0.5 * h * v, v + 0.5 * h * a1
    a2 = qacc(q2, v2, ctrl, params)
    q3, v3 = q + 0.5 * h * v2, v + 0.5 * h * a2
    a3 = qacc(q3, v3, ctrl, params)
    q4, v4 = q + h * v3, v + h * a3
    a4 = qacc(q4, v4, ctrl, params)
    return q + h / 6 * (v + 2 * v2 + 2 * v3 + v4), v + h / 6 * (a1 + 2 * a2 + 2 * a3 + a4)


def xpos_of_segment_end(q):
    p1 = np.pi + q[..., 1]
    return q[..., 0] + L1 * np.sin(p1) + L2_END * np.sin(p1 + q[..., 2])


def rollout(env_id, actions, control_timestep=0.01, params=None, state=None, return_state=False):
    """actions [B, T] (raw, before arctan) -> observations [B, T + 1] in fp64: the observation after reset, then one per
    control step (control.Environment.step: before_step sets ctrl, n_sub_steps physics steps).  The arctan is evaluated in
    the dtype of `actions` (np.arctan of the float32 actions the reference's actors produce is a float32 operation; Python
    floats / float64 arrays give a float64 one).  params: 4 numbers or [B, 4]; state [B, 6] = (qpos, qvel) to start from."""
    params = ENV_PARAMS[env_id] if params is None else params
    params = tuple(np.asarray(params, np.float64).T)
    actions = np.atleast_2d(np.asarray(actions))
    if actions.dtype != np.float32:
        actions = actions.astype(np.float64)
    B, T = actions.shape
    n_sub = int(round(control_timestep / PHYSICS_DT))
    q = np.zeros((B, 3)) if state is None else np.array(state[:, :3], np.float64)
    v = np.zeros((B, 3)) if state is None else np.array(state[:, 3:], np.float64)
    out = np.empty((B, T + 1))
    out[:, 0] = xpos_of_segment_end(q)
    for k in range(T):
        ctrl = np.arctan(actions[:, k]).astype(np.float64)
        for _ in range(n_sub):
            q, v = rk4_step(q, v, ctrl, params)
        out[:, k + 1] = xpos_of_segment_end(q)
    return (out, np.concatenate([q, v], 1)) if return_state else out
