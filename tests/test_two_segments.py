"""Two-segment data source: the reference's golden vectors (cc/env/test_make_env.py:146-335, committed as
tests/golden/two_segments_golden.json by make_two_segments_golden.py) pin the NumPy oracle; the CUDA kernel is compared with
the golden vectors and with the oracle (emulated on the CPU, real on the GPU) through the C ABI."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import two_segments_np as TS

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "two_segments_golden.json")))
ENVS = ["two_segments_v1", "two_segments_v2", "two_segments_v3"]


def golden(env_id):
    # observation 0 = reset (dm_env resets on the first step and ignores that action), observation i follows DYNAMIC_INPUTS[i]
    return np.array(GOLD["dynamic_inputs"][1:], np.float64)[None], np.array(GOLD["expected_positions"][env_id], np.float32)


def abi_rollout(L, params, actions, state=None, n_sub=10, T_=None):
    """numpy (emulated library, host pointers) or torch (product library, device pointers; T_ = to-device) call"""
    params = np.ascontiguousarray(np.asarray(params, np.float64).reshape(-1, 4))
    B, T = actions.shape
    if T_ is None:
        a = np.ascontiguousarray(actions)
        obs = np.full((B, T + 1), np.nan, np.float32)
        st = None if state is None else np.ascontiguousarray(state, np.float64)
        ptr = lambda x: None if x is None else x.ctypes.data
        stream = None
        tensors = (params, a, obs, st)
    else:
        import torch

        a = T_(np.ascontiguousarray(actions))
        obs = torch.full((B, T + 1), float("nan"), dtype=torch.float32, device=a.device)
        st = None if state is None else T_(np.ascontiguousarray(state, np.float64))
        params = T_(params)
        ptr = lambda x: None if x is None else x.data_ptr()
        stream = torch.cuda.current_stream().cuda_stream
        tensors = (params, a, obs, st)
    f64 = actions.dtype == np.float64
    rc = L._dll.cc_env_two_segments_rollout(ptr(params), params.shape[0], None if f64 or T == 0 else ptr(a), ptr(a) if f64 and T else None,
                                            ptr(obs), ptr(st), B, T, n_sub, C.c_double(1e-3), stream)
    L.check(rc, "cc_env_two_segments_rollout")
    del tensors
    if T_ is None:
        return obs, st
    return obs.cpu().numpy(), None if st is None else st.cpu().numpy()


# ------------------------------------------------------------------ oracle vs the reference's golden vectors (CPU)
@pytest.mark.parametrize("env_id", ENVS)
def test_oracle_reproduces_reference_golden_vectors(env_id):
    acts, want = golden(env_id)
    got = TS.rollout(env_id, acts)[0]
    assert (got.astype(np.float32) == want).all()  # the reference asserts equality of the float32 observations
    assert np.abs(got - want).max() < 3e-9


def test_oracle_capsule_inertia_and_reset_observation():
    # uniform-density capsule: between the thin rod m L^2 / 12 of the cylinder part and that of the full length L + 2 r
    I = TS.capsule_inertia_perp()
    assert 0.1 * 1.0 / 12 < I < 0.1 * 1.09 ** 2 / 12 + 0.1 * 0.045 ** 2 / 4
    assert TS.rollout("two_segments_v1", np.zeros((1, 0)))[0, 0] == np.float64(2.1) * np.sin(np.pi)  # 2.57e-16: golden[0]


# ------------------------------------------------------------------ emulated kernel (CPU fibers) vs oracle / golden
@pytest.mark.parametrize("env_id", ENVS)
def test_emu_kernel_reproduces_golden_vectors(env_id):
    import emu_lib as E

    acts, want = golden(env_id)
    p = TS.ENV_PARAMS[env_id]
    obs, _ = abi_rollout(E.emu(), p, acts)
    assert (obs[0] == want).all()


def _cases():
    rng = np.random.default_rng(0)
    return rng.uniform(-3, 3, size=(70, 45)).astype(np.float32), rng.uniform(0.5, 2.0, size=(70, 4)) * np.array([1e-2, 1.0, 5e-2, 4.0])


def test_emu_kernel_vs_oracle_ragged_batch_per_env_params_and_state_carry():
    import emu_lib as E

    acts, params = _cases()  # 70 environments = 3 warps with a ragged tail; 45 steps = one full chunk of 32 + 13
    ref, ref_state = TS.rollout(None, acts, params=params, return_state=True)
    state = np.zeros((70, 6))
    obs, state = abi_rollout(E.emu(), params, acts, state)
    assert np.abs(obs - ref).max() < 1e-6 * np.abs(ref).max()  # (float32 arctan of numpy vs libm's atanf: 1 ulp of the control)
    assert np.abs(state - ref_state).max() < 1e-6 * np.abs(ref_state).max()
    # fp64 actions: same arctan on both sides -> agreement at fp64 rounding (observations are rounded to float32)
    a64 = acts.astype(np.float64)
    ref64, rs64 = TS.rollout(None, a64, params=params, return_state=True)
    st = np.zeros((70, 6))
    o64, st = abi_rollout(E.emu(), params, a64, st)
    assert (o64 == ref64.astype(np.float32)).mean() > 0.999 and np.abs(o64 - ref64).max() < 1e-6
    assert np.abs(st - rs64).max() < 1e-11
    # an episode in two calls (state carried) = one call; shared parameter row; control step of 3 physics steps
    s2 = np.zeros((70, 6))
    oa, s2 = abi_rollout(E.emu(), params[0], a64[:, :20], s2, n_sub=3)
    ob, s2 = abi_rollout(E.emu(), params[0], a64[:, 20:], s2, n_sub=3)
    oc, _ = abi_rollout(E.emu(), params[0], a64, None, n_sub=3)
    assert (np.concatenate([oa, ob[:, 1:]], 1) == oc).all() and (ob[:, 0] == oa[:, -1]).all()


def test_abi_rejects_bad_arguments():
    import emu_lib as E

    L = E.emu()
    p = np.zeros((2, 4))
    a = np.zeros((3, 4), np.float32)
    o = np.zeros((3, 5), np.float32)
    call = L._dll.cc_env_two_segments_rollout
    assert call(p.ctypes.data, 2, a.ctypes.data, None, o.ctypes.data, None, 3, 4, 10, C.c_double(1e-3), None) == -1  # n_params not in {1, B}
    assert call(p.ctypes.data, 1, None, None, o.ctypes.data, None, 3, 4, 10, C.c_double(1e-3), None) == -1  # no actions
    assert call(p.ctypes.data, 1, a.ctypes.data, a.ctypes.data, o.ctypes.data, None, 3, 4, 10, C.c_double(1e-3), None) == -1  # both
    assert call(p.ctypes.data, 1, a.ctypes.data, None, o.ctypes.data, None, 3, 4, 0, C.c_double(1e-3), None) == -1
    assert "two_segments" in L.last_error()


# ------------------------------------------------------------------ host-side generators (CPU)
def test_input_and_reference_generators():
    from chain_control_b200.env import collect as CL

    ts = np.arange(10.0, step=0.01)
    u = CL.draw_u_from_gaussian_process(ts, seed=3)
    assert u.shape == (1000, 1) and u.dtype == np.float32
    assert abs(float(u.std()) - 0.25) < 1e-6 and abs(float(u.mean())) < 1e-7  # standardised, times scale_u_by
    assert (u == CL.draw_u_from_gaussian_process(ts, seed=3)).all() and not (u == CL.draw_u_from_gaussian_process(ts, seed=4)).all()
    c = CL.draw_u_from_cosines(ts, 4)
    assert c.shape == (1000, 1) and c[0, 0] == 2.0 and np.allclose(c[:, 0], 2 * np.cos(8 * np.pi * ts / ts[-1]))

    class _Env:  # the generators only read the time grid and the observation key
        obs_key = "xpos_of_segment_end"

        def timestep_array(self):
            return ts

    src = CL.random_steps_source(_Env(), seeds=[0, 1, 2])
    yss = src.get_references()["xpos_of_segment_end"]
    assert yss.shape == (3, 1001, 1) and (yss == yss[:, :1]).all() and (np.abs(yss) <= 3).all()
    np.random.seed(1)
    sign = np.random.choice([-1.0, 1.0])
    assert yss[1, 0, 0] == sign * np.random.uniform(0.0, 3.0, size=(1, 1))[0, 0]  # circus.py:22-33 draw order
    src.change_reference_of_actor(2)
    assert float(src.get_reference_actor()["xpos_of_segment_end"][5, 0]) == yss[2, 5, 0]


def test_make_env_without_gpu_fails_loudly():
    import torch

    from chain_control_b200 import _lib
    from chain_control_b200.env import make_env

    with pytest.raises(Exception, match="Unknown environment id"):
        make_env("rover")
    if not torch.cuda.is_available():
        with pytest.raises(_lib.CcError, match="no CPU path"):
            make_env("two_segments_v1", device="cpu")


# ------------------------------------------------------------------ GPU: the product kernel through the C ABI
@pytest.fixture
def T_():
    import torch

    return lambda a: torch.as_tensor(a, device="cuda")


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ENVS)
def test_gpu_kernel_reproduces_golden_vectors(env_id, T_):
    from chain_control_b200 import _lib

    acts, want = golden(env_id)
    obs, _ = abi_rollout(_lib.lib(), TS.ENV_PARAMS[env_id], acts, T_=T_)
    # CUDA's fp64 sincos / atan are not libm's: agreement to the last float32 bit is expected but not guaranteed (<= 1 ulp)
    assert np.abs(obs[0].astype(np.float64) - want).max() <= 1.2e-7 * np.abs(want).max()
    assert (obs[0] == want).mean() >= 0.9


@pytest.mark.gpu
def test_gpu_kernel_vs_oracle_full_episode(T_):
    from chain_control_b200 import _lib

    rng = np.random.default_rng(1)
    B, T = 1000 + 37, 1000
    # smooth inputs (sums of cosines) like the reference's data collection, one parameter set per environment
    t = np.arange(T) * 0.01
    acts = (rng.uniform(0.2, 1.5, (B, 1)) * np.cos(rng.uniform(0.5, 8.0, (B, 1)) * t + rng.uniform(0, 6.28, (B, 1)))).astype(np.float64)
    params = rng.uniform(0.5, 2.0, size=(B, 4)) * np.array([1e-2, 1.0, 5e-2, 4.0])
    ref, rstate = TS.rollout(None, acts, params=params, return_state=True)
    st = np.zeros((B, 6))
    obs, st = abi_rollout(_lib.lib(), params, acts, st, T_=T_)
    err = np.abs(obs - ref).max(axis=1) / np.abs(ref).max(axis=1)
    assert err.max() < 1e-6  # float32 observations of an fp64 simulation
    assert np.abs(st - rstate).max() < 1e-9 * np.abs(rstate).max()
    o2, _ = abi_rollout(_lib.lib(), params, acts, np.zeros((B, 6)), T_=T_)
    assert (o2 == obs).all()
    # float32 actions: arctan in fp32 on both sides
    a32 = acts.astype(np.float32)
    r32 = TS.rollout(None, a32[:64], params=params[:64])
    o32, _ = abi_rollout(_lib.lib(), params[:64], a32[:64], T_=T_)
    assert (np.abs(o32 - r32).max(axis=1) / np.abs(r32).max(axis=1)).max() < 1e-5


@pytest.mark.gpu
def test_gpu_make_env_step_rollout_and_collect(T_):
    import torch

    from chain_control_b200.env import collect as CL
    from chain_control_b200.env import make_env

    env = make_env("two_segments_v1", time_limit=1.0)
    assert env.time_limit() == 1.0 and len(env.timestep_array()) == 100
    # dm_env protocol: the first step resets; an episode = 100 steps, then the next step resets again
    acts, want = golden("two_segments_v1")
    o, last = env.step(123.0)
    assert not last and abs(float(o[env.obs_key][0, 0]) - want[0]) < 1e-20
    seq = [float(o[env.obs_key][0, 0])]
    for k in range(39):
        o, last = env.step(float(acts[0, k]))  # a Python float, like the reference's golden test
        seq.append(float(o[env.obs_key][0, 0]))
    assert np.abs(np.array(seq) - want).max() <= 1.2e-7 * np.abs(want).max()
    for k in range(61):
        o, last = env.step(0.0)
    assert last
    o, last = env.step(5.0)
    assert not last and abs(float(o[env.obs_key][0, 0]) - want[0]) < 1e-20
    # feed-forward collection: 3 GP draws + 2 cosines in one launch each; shapes of docs/2_collecting_data notebook
    env = make_env("two_segments_v1", time_limit=10.0)
    sample = CL.sample_feedforward_and_collect(env, seeds_gp=[0, 1, 2], seeds_cos=[2, 4])
    assert sample.action.shape == (5, 1000, 1) and sample.obs[env.obs_key].shape == (5, 1001, 1) and sample.bs == 5
    ref = TS.rollout("two_segments_v1", sample.action[..., 0].astype(np.float32))
    got = sample.obs[env.obs_key][..., 0]
    assert (np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)).max() < 1e-5
    assert np.abs(got).max() > 0.05  # the chain actually moves


@pytest.mark.gpu
def test_gpu_closed_loop_collection_with_a_controller(T_):
    """collect_exhaust_source: controller kernel and environment kernel alternate; consistency with (i) the oracle environment
    driven by the recorded actions and (ii) the controller unrolled open loop over the recorded [ref ; obs] series"""
    import torch

    from chain_control_b200 import ops
    from chain_control_b200.env import collect as CL
    from chain_control_b200.env import make_env
    from chain_control_b200.examples.neural_ode import make_neural_ode_controller

    env = make_env("two_segments_v2", time_limit=2.0)
    src = CL.random_steps_source(env, seeds=[0, 1, 2, 3, 4])
    specs = dict(input_specs=dict(ref=(1,), obs=(1,)), output_specs=(1,))
    ctrl = make_neural_ode_controller(control_timestep=0.01, state_dim=6, key=3, f_width_size=10, f_depth=2, **specs)
    sample = CL.collect_exhaust_source(env, src, ctrl)
    obs, act = sample.obs[env.obs_key], sample.action
    assert obs.shape == (5, 201, 1) and act.shape == (5, 200, 1)
    assert sample.rew.shape == (5, 200) and np.allclose(sample.rew, -((sample.extras["ref"][:, 1:201, 0] - obs[:, 1:, 0]) ** 2), atol=1e-6)
    ref = TS.rollout("two_segments_v2", act[..., 0].astype(np.float32))
    assert np.abs(obs[..., 0] - ref).max() < 1e-5 * max(np.abs(ref).max(), 1e-3)
    inp = torch.as_tensor(np.concatenate([sample.extras["ref"][:, :200], obs[:, :200]], axis=2), device="cuda")
    y = ops.batched_unroll(ctrl.spec, ctrl.theta, inp.contiguous(), None)
    assert np.abs(y[:, 1:, :].detach().cpu().numpy() - act).max() < 1e-5 * max(np.abs(act).max(), 1e-3)


@pytest.mark.gpu
def test_gpu_end_to_end_workflow_collect_train_model_train_controller_evaluate():
    """the reference's documented workflow (docs/2_collecting_data..., 3_training_a_model, 4_training_a_controller notebooks) on this
    framework only: record trajectories from the two-segment source, fit the compact neural-ODE model (ModelTrainer step =
    the BPTT kernels), train a compact controller against the fitted model (ControllerTrainer step = closed-loop BPTT kernels),
    then run that controller in closed loop with the ENVIRONMENT kernel"""
    import torch
    from collections import OrderedDict

    from chain_control_b200.env import collect as CL
    from chain_control_b200.env import make_env
    from chain_control_b200.examples.neural_ode_controller_compact_example import make_neural_ode_controller
    from chain_control_b200.examples.neural_ode_model_compact_example import make_neural_ode_model
    from chain_control_b200.train import (DictLogger, EvaluationMetrices, ModelControllerTrainer, SupervisedDataset, Tracker,
                                          TrainingOptionsController, TrainingOptionsModel, UnsupervisedDataset, make_dataloader, optim)
    from chain_control_b200.utils import rmse

    env = make_env("two_segments_v1", time_limit=5.0)
    key = env.obs_key
    train = CL.sample_feedforward_and_collect(env, seeds_gp=[0, 1, 2, 3], seeds_cos=[4, 8, 12, 2])  # defaults.py:21-23
    val = CL.sample_feedforward_and_collect(env, seeds_gp=[15, 16], seeds_cos=[2.5, 5.0])
    assert train.action.shape == (8, 500, 1) and train.obs[key].shape == (8, 501, 1)
    to = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32), device="cuda")
    u_tr, y_tr = to(train.action), OrderedDict([(key, to(train.obs[key]))])
    u_va, y_va = to(val.action), OrderedDict([(key, to(val.obs[key]))])

    model = make_neural_ode_model(1, OrderedDict([(key, 1)]), 0.01, state_dim=20, f_depth=0, u_transform=np.arctan)
    opts = TrainingOptionsModel(make_dataloader(SupervisedDataset(u_tr, y_tr), key=1, n_minibatches=1),
                                optim.chain(optim.clip_by_global_norm(1.0), optim.adam(3e-3)),
                                metrices=(EvaluationMetrices(data=(u_va, y_va), metrices=(lambda y, yhat: {"val_rmse": rmse(y, yhat)},)),))
    trainer = ModelControllerTrainer(model, model_train_options=opts, loggers=[DictLogger()], trackers=[Tracker("val_rmse")])
    trainer.run(150)
    logs = trainer.get_logs()[0]
    assert np.all(np.isfinite(logs["train_mse"]))
    assert logs["train_mse"][-10:].mean() < 0.5 * logs["train_mse"][:3].mean(), (logs["train_mse"][:3], logs["train_mse"][-3:])
    fitted = trainer.trackers[0].best_model_or_controller()

    src = CL.random_steps_source(env, seeds=list(range(6)), max_abs_amplitude=1.0)
    refs = OrderedDict([(key, to(src.get_references()[key]))])
    controller = make_neural_ode_controller(2, 1, 0.01, 5, f_depth=0)
    copts = TrainingOptionsController(make_dataloader(UnsupervisedDataset(refs), key=1, n_minibatches=1),
                                      optim.chain(optim.clip_by_global_norm(1.0), optim.adam(1e-2)))
    ct = ModelControllerTrainer(fitted, controller, controller_train_options=copts, loggers=[DictLogger()], trackers=[Tracker("loss")])
    ct.run(40)
    clogs = ct.get_logs()[0]
    assert np.all(np.isfinite(clogs["loss"])) and clogs["loss"][-5:].mean() < clogs["loss"][:2].mean()
    best = ct.trackers[0].best_model_or_controller()
    sample = CL.collect_exhaust_source(env, src, best)
    assert sample.obs[key].shape == (6, 501, 1) and sample.action.shape == (6, 500, 1)
    assert np.all(np.isfinite(sample.obs[key])) and np.all(np.isfinite(sample.action))
