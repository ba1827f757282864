"""GPU tests of the host-side mirrors, modelled on the reference's own tests for this path:
cc/train/test_trainer.py:25-147 (one step of each trainer, log keys, trackers, multi-model dict) and
cc/core/test_module.py:84-103 (the scan contract: unroll == successive steps, include_y0)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from cchelpers import TOL_GRAD, TOL_STATE, l2rel, normwise
from oracle import c_oracle as CO
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def _data(B, T, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(T)[None, :, None] * 0.01
    u = (np.cos(2 * np.pi * rng.uniform(1, 6, size=(B, 1, 1)) * t / 10) * 0.5).astype(np.float32)
    return u


def test_trainer_steps_and_log_keys():
    from chain_control_b200.examples.neural_ode_controller_compact_example import make_neural_ode_controller
    from chain_control_b200.examples.neural_ode_model import make_neural_ode_model
    from chain_control_b200.train import (DictLogger, EvaluationMetrices, ModelControllerTrainer, Regularisation,
                                          SupervisedDataset, Tracker, TrainingOptionsController, TrainingOptionsModel,
                                          UnsupervisedDataset, make_dataloader, optim)
    from chain_control_b200.utils import l2_norm, rmse

    T = 200
    act_spec, obs_spec = 1, OrderedDict(xpos_of_segment_end=1)
    model = make_neural_ode_model(act_spec, obs_spec, 0.01, state_dim=1, f_depth=0, u_transform=np.arctan, f_time_invariant=False)
    u_tr, u_va = _data(2, T, 0), _data(2, T, 1)
    teacher = make_neural_ode_model(act_spec, obs_spec, 0.01, state_dim=3, f_depth=0, key=7, u_transform=np.arctan)
    with torch.no_grad():
        y_tr, y_va = teacher.unroll_batched(u_tr), teacher.unroll_batched(u_va)
    assert tuple(y_tr["xpos_of_segment_end"].shape) == (2, T + 1, 1)

    opts = TrainingOptionsModel(
        make_dataloader(SupervisedDataset(u_tr, y_tr), key=2, n_minibatches=1),
        optim.chain(optim.clip_by_global_norm(1.0), optim.adam(3e-3)),
        regularisers=(Regularisation(0.5, lambda v: {"l2_norm": l2_norm(v)}),),
        metrices=(EvaluationMetrices(data=(u_va, y_va), metrices=(lambda y, yhat: {"val_rmse": rmse(y, yhat)},)),))
    trainer = ModelControllerTrainer(model, model_train_options=opts, loggers=[DictLogger()], trackers=[Tracker("val_rmse")])
    trainer.run(2)
    logs = trainer.get_logs()[0]
    for k in ("train_loss", "train_mse", "l2_norm", "val_rmse"):
        assert k in logs and logs[k].shape[0] == 2 and np.all(np.isfinite(logs[k]))
    fitted = trainer.trackers[0].best_model_or_controller()
    assert fitted is not None

    refs = OrderedDict(xpos_of_segment_end=(np.sign(np.random.default_rng(0).normal(size=(4, 1, 1))) * np.ones((4, T + 1, 1))).astype(np.float32))
    controller = make_neural_ode_controller(2, act_spec, 0.01, 1, f_depth=0)
    copts = TrainingOptionsController(make_dataloader(UnsupervisedDataset(refs), key=1, n_minibatches=1),
                                      optim.chain(optim.clip_by_global_norm(1.0), optim.adam(1e-3)))
    ct = ModelControllerTrainer(fitted, controller, controller_train_options=copts, trackers=[Tracker("model0", "train_mse")])
    logs = ct.step()
    for k in ("train_loss", "loss", "loss_without_regu", "model0_train_mse"):
        assert k in logs
    model2 = make_neural_ode_model(act_spec, obs_spec, 0.01, state_dim=2, f_depth=0, u_transform=np.arctan)
    ct = ModelControllerTrainer(dict(m1=fitted, m2=model2), controller, controller_train_options=copts,
                                trackers=[Tracker("m2", "train_mse"), Tracker("loss_without_regu")])
    logs = ct.step()
    assert "m1_train_mse" in logs and "m2_train_mse" in logs
    assert abs(logs["loss_without_regu"] - 0.5 * (logs["m1_train_mse"] + logs["m2_train_mse"])) < 1e-6


def test_model_train_step_gradient_matches_oracle():
    """value_and_grad(mse(vmap(unroll))) through the mirror == the oracle's fp64 gradient."""
    from chain_control_b200.examples.neural_ode_model_compact_example import make_neural_ode_model
    from chain_control_b200.utils import mse

    model = make_neural_ode_model(1, 1, 0.01, state_dim=50, f_depth=0, u_transform=np.arctan, key=3)
    u = _data(8, 300, 4)
    targets = np.random.default_rng(2).normal(size=(8, 301, 1)).astype(np.float32)
    theta = model.flat_params().detach().requires_grad_(True)
    loss = mse(torch.tensor(targets, device="cuda"), model.with_params(theta).unroll_batched(u))
    (g,) = torch.autograd.grad(loss, theta)
    ref = CO.rollout(model.spec, theta.detach().cpu().numpy(), u, targets=targets, dtype=np.float64)
    assert abs(float(loss.detach()) - ref["loss"]) < 1e-5 * max(1.0, ref["loss"])
    assert l2rel(g.cpu().numpy(), ref["theta_bar"]) < TOL_GRAD


def test_unroll_equals_successive_steps():
    """scan contract of cc/core/test_module.py:84-103 on a neural ODE: unroll == step, step, ...; include_y0."""
    from chain_control_b200.examples.neural_ode_model import make_neural_ode_model

    for kw in (dict(f_depth=1, f_width_size=7), dict(f_depth=0, f_time_invariant=False, g_time_invariant=False)):
        model = make_neural_ode_model(1, 2, 0.01, state_dim=4, key=5, **kw)
        u = _data(1, 6, 9)[0]
        y = model.unroll(u)
        assert tuple(y.shape) == (7, 2)
        np.testing.assert_allclose(model.unroll(u, include_y0=False).cpu().numpy(), y[1:].cpu().numpy(), rtol=0, atol=0)
        np.testing.assert_allclose(y[0].cpu().numpy(), model.y0().cpu().numpy(), rtol=1e-6, atol=1e-7)
        m = model.reset()
        for k in range(6):
            m, yk = m.step(u[k])
            np.testing.assert_allclose(yk.cpu().numpy(), y[k + 1].cpu().numpy(), rtol=2e-6, atol=1e-7)


def test_closed_loop_mirror_matches_oracle():
    from chain_control_b200.core import filter_vmap
    from chain_control_b200.examples.neural_ode_controller import make_neural_ode_controller
    from chain_control_b200.examples.neural_ode_model_compact_example import make_neural_ode_model
    from chain_control_b200.train import merge_x_y

    model = make_neural_ode_model(1, 1, 0.01, state_dim=10, f_depth=0, u_transform=np.arctan, key=3)
    controller = make_neural_ode_controller(2, 1, 0.01, state_dim=25, f_width_size=10, f_depth=2, key=4)  # CLI defaults
    refs = (np.random.default_rng(1).uniform(-2, 2, size=(16, 1, 1)) * np.ones((16, 151, 1))).astype(np.float32)
    yhat = filter_vmap(controller.unroll(model, merge_x_y))(refs, 0)
    ref = CO.closedloop(controller.spec, model.spec, controller.theta.cpu().numpy(), model.theta.cpu().numpy(), refs, dtype=np.float64)
    assert normwise(yhat.detach().cpu().numpy(), ref["yhat"]) < TOL_STATE
    single = controller.unroll(model, merge_x_y)(refs[3], 0)
    np.testing.assert_allclose(single.detach().cpu().numpy(), yhat[3].detach().cpu().numpy(), rtol=1e-6, atol=1e-7)


def _gold(name):
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))


@pytest.mark.parametrize("name,dims", [("open_linear_ct_s3", (3, 1, 1, True)), ("open_linear_dt_s6_i2_o2_x0", (6, 2, 2, False))])
def test_linear_model_mirror_vs_reference_golden(name, dims):
    """LinearModel (cc/examples/linear_model.py:47-84) through the public mirror: outputs and the gradient of the reference's
    loss_fn_model in the REFERENCE's flat order [A, B, C, D] against the fixture produced by the reference's own code"""
    from chain_control_b200.examples.linear_model import make_linear_model
    from chain_control_b200.utils import mse

    S, I, O, ct = dims
    g = _gold(name)
    osp = O if O == 1 else OrderedDict(o0=1, o1=1)
    m = make_linear_model(I, osp, float(np.float32(0.01)), S, continuous_time=ct, randomize_x0=bool(np.any(g["x0"])))
    m.init_state = (torch.tensor(g["x0"], dtype=torch.float32, device="cuda"), 0.0)
    m._x0_nonzero = bool(np.any(g["x0"]))
    p = torch.tensor(g["theta_ref_order"], dtype=torch.float32, device="cuda", requires_grad=True)
    m = m.with_params(p)
    y = m.unroll_batched(torch.tensor(g["u"], dtype=torch.float32, device="cuda"))
    yf = y if torch.is_tensor(y) else torch.cat(list(y.values()), -1)
    assert normwise(yf.detach().cpu().numpy(), g["y"]) < TOL_STATE
    loss = ((yf - torch.tensor(g["targets"], dtype=torch.float32, device="cuda")) ** 2).mean()
    (gr,) = torch.autograd.grad(loss, p)
    assert abs(float(loss.detach()) - float(g["train_mse"])) < 1e-5 * max(1.0, float(g["train_mse"]))
    assert l2rel(gr.cpu().numpy(), g["theta_bar_ref_order"]) < TOL_GRAD


def test_pid_controller_mirror_vs_reference_golden():
    """PIDController (pid_controller.py:34-79) in closed loop with a LinearModel: outputs and d loss / d (d, i, p gains)"""
    from chain_control_b200.examples.linear_model import make_linear_model
    from chain_control_b200.examples.pid_controller import make_pid_controller
    from chain_control_b200.train.step_fn import merge_x_y

    g = _gold("closed_pid__linear")
    dt = float(g["pid_dt"])
    d, i, p = g["pid_gains_dip"]
    gains = torch.tensor(g["theta_c_ref_order"], dtype=torch.float32, device="cuda", requires_grad=True)
    c = make_pid_controller(p, i, d, dt).with_params(gains)
    thm = g["theta_m__sys0"]  # kernel order of a zero-D LinearModel: [A|B] rows then C
    A_, B_, C_ = thm[:12].reshape(3, 4)[:, :3], thm[:12].reshape(3, 4)[:, 3:], thm[12:].reshape(1, 3)
    m = make_linear_model(1, 1, dt, 3).replace(A=A_, B=B_, C=C_, D=np.zeros((1, 1)))
    refs = torch.tensor(g["refs"], dtype=torch.float32, device="cuda")
    yhat = c.unroll(m, merge_x_y).batched(refs)
    assert normwise(yhat.detach().cpu().numpy(), g["yhat__sys0"]) < TOL_STATE
    loss = ((yhat - refs) ** 2).mean()
    (gr,) = torch.autograd.grad(loss, gains)
    assert l2rel(gr.cpu().numpy(), g["theta_c_bar"]) < TOL_GRAD


def test_cli_workflow_linear_systems():
    """the reference's CLI workflow (docs/examples/cli/train_controller.py:120-190): a deep-MLP controller trained against
    several LTI systems given as (A, B, C) with D = 0; the multi-model mean loss falls"""
    from chain_control_b200.examples.linear_model import make_linear_model
    from chain_control_b200.examples.neural_ode_controller import make_neural_ode_controller
    from chain_control_b200.train import (DictLogger, ModelControllerTrainer, Tracker, TrainingOptionsController,
                                          UnsupervisedDataset, make_dataloader, optim)

    rng = np.random.default_rng(0)
    systems = {}
    for k in range(3):
        S = 3
        A_ = -np.eye(S) * rng.uniform(0.5, 2.0) + 0.3 * rng.normal(size=(S, S))
        systems[f"system{k}"] = make_linear_model(1, OrderedDict(output=1), 0.01, 3).replace(
            A=A_, B=rng.normal(size=(S, 1)), C=rng.normal(size=(1, S)), D=np.zeros((1, 1)))
    refs = OrderedDict(output=(rng.uniform(-1, 1, size=(30, 1, 1)) * np.ones((30, 301, 1))).astype(np.float32))
    controller = make_neural_ode_controller(2, 1, 0.01, state_dim=5, f_depth=2, f_width_size=10)
    copts = TrainingOptionsController(make_dataloader(UnsupervisedDataset(refs), key=1, n_minibatches=1),
                                      optim.chain(optim.clip_by_global_norm(1.0), optim.adam(3e-3)))
    tr = ModelControllerTrainer(systems, controller, controller_train_options=copts, trackers=[Tracker("loss")], loggers=[DictLogger()])
    tr.run(6)
    logs = tr.get_logs()[0]
    assert all(f"system{k}_train_mse" in logs for k in range(3))
    assert np.all(np.isfinite(logs["loss"])) and logs["loss"][-1] < logs["loss"][0]


def test_superposed_compact_controller_vs_reference_golden():
    """compact controller with the PID secondary (neural_ode_controller_compact_example.py:88-168, superposition=True) through
    the public factory: outputs (with the recorded observation noise) and the gradient w.r.t. [f, g, d, i, p gains]"""
    import json

    from cchelpers import conv
    from chain_control_b200 import ops
    from chain_control_b200.examples.neural_ode_controller_compact_example import make_neural_ode_controller

    g = _gold("closed_compact_superposition__compact_s50")
    dt = float(g["superposed_dt"])
    c = make_neural_ode_controller(2, 1, dt, 5, f_depth=0, superposition=True)
    flat = torch.tensor(g["theta_c_ref_order"], dtype=torch.float32, device="cuda", requires_grad=True)
    c = c.with_params(flat)
    assert c.flat_params().numel() == g["theta_c_ref_order"].size == 5 * 7 + 5 + 5 + 1 + 3
    mspec = conv(O.make_node(**json.loads(str(g["mspecs"]))["m0"]))
    t = lambda a: torch.tensor(np.asarray(a, np.float32), device="cuda")
    refs = t(g["refs"])
    yhat = ops.batched_closed_loop(c.spec, mspec, c.theta, t(g["theta_m__m0"]), refs, t(g["noise__m0"]))
    assert normwise(yhat.detach().cpu().numpy(), g["yhat__m0"]) < TOL_STATE
    loss = ((yhat - refs) ** 2).mean()
    (gr,) = torch.autograd.grad(loss, flat)
    assert l2rel(gr.cpu().numpy(), g["theta_c_bar"]) < TOL_GRAD
    with pytest.raises(NotImplementedError):
        make_neural_ode_controller(2, 1, dt, 5, f_depth=2, superposition=True)


def test_multi_model_streams_match_serial(monkeypatch):
    """the multi-model mean loss (step_fn.py:256-271) with one CUDA stream per model gives the same step as the serial loop"""
    import time

    from chain_control_b200.examples.linear_model import make_linear_model
    from chain_control_b200.examples.neural_ode_controller import make_neural_ode_controller
    from chain_control_b200.train import (ModelControllerTrainer, TrainingOptionsController, UnsupervisedDataset, make_dataloader,
                                          optim)

    def run(streams):
        monkeypatch.setenv("CC_B200_MODEL_STREAMS", streams)
        rng = np.random.default_rng(0)
        systems = {}
        for k in range(8):  # the CLI's 8 LTI systems x 30 references x 1501 steps
            A_ = -np.eye(3) * rng.uniform(0.5, 2.0) + 0.3 * rng.normal(size=(3, 3))
            systems[f"system{k}"] = make_linear_model(1, OrderedDict(output=1), 0.01, 3).replace(
                A=A_, B=rng.normal(size=(3, 1)), C=rng.normal(size=(1, 3)), D=np.zeros((1, 1)))
        refs = OrderedDict(output=(rng.uniform(-1, 1, size=(30, 1, 1)) * np.ones((30, 1501, 1))).astype(np.float32))
        controller = make_neural_ode_controller(2, 1, 0.01, state_dim=5, f_depth=2, f_width_size=10, key=3)
        copts = TrainingOptionsController(make_dataloader(UnsupervisedDataset(refs), key=1, n_minibatches=1),
                                          optim.chain(optim.clip_by_global_norm(1.0), optim.adam(3e-3)))
        tr = ModelControllerTrainer(systems, controller, controller_train_options=copts)
        tr.step()
        torch.cuda.synchronize()
        t0 = time.time()
        logs = [tr.step() for _ in range(3)]
        torch.cuda.synchronize()
        return logs[-1], (time.time() - t0) / 3

    l1, t1 = run("1")
    l0, t0 = run("0")
    print(f"[multi-model] 8 systems x 30 x 1501: {1e3 * t1:.1f} ms/step on 8 streams, {1e3 * t0:.1f} ms/step serial")
    for k in l0:
        assert abs(float(l0[k]) - float(l1[k])) <= 1e-6 * max(1.0, abs(float(l0[k]))), k


def test_autograd_entry_points_checkpoint_automatically():
    """ops.set_checkpointing: with a small tape budget the differentiable entry points switch to ckpt_every = ceil(sqrt(T)) on
    their own (forward writes checkpoints only, the reverse pass replays segments); gradients agree with the full-tape ones"""
    from chain_control_b200 import ops
    from chain_control_b200.examples.neural_ode_controller_compact_example import make_neural_ode_controller
    from chain_control_b200.examples.neural_ode_model import make_neural_ode_model
    from chain_control_b200.examples.neural_ode_model_compact_example import make_neural_ode_model as make_compact

    obs_spec = OrderedDict(xpos_of_segment_end=1)
    B, T = 1100, 90
    u = torch.rand(B, T, 1, device="cuda") * 2 - 1
    grads = {}
    try:
        for tag, kw in (("full", dict(max_tape_bytes=None)), ("auto", dict(max_tape_bytes=1 << 16)), ("k7", dict(every=7))):
            ops.set_checkpointing(**kw)
            for name, model in (("deep64", make_neural_ode_model(1, obs_spec, 0.01, state_dim=40, f_width_size=64, f_depth=2, key=3)),
                                ("default", make_neural_ode_model(1, obs_spec, 0.01, state_dim=5, key=4)),
                                ("compact50", make_compact(1, obs_spec, 0.01, state_dim=50, f_depth=0, key=5, u_transform=np.arctan))):
                theta = model.theta.clone().requires_grad_(True)
                y = model.with_params(theta).unroll_batched(u)["xpos_of_segment_end"]
                (y.square().mean()).backward()
                grads[(tag, name)] = theta.grad.clone()
            model = make_compact(1, obs_spec, 0.01, state_dim=50, f_depth=0, key=5, u_transform=np.arctan)
            ctrl = make_neural_ode_controller(2, 1, 0.01, 5, f_depth=0, key=6)
            theta = ctrl.theta.clone().requires_grad_(True)
            refs = OrderedDict(xpos_of_segment_end=torch.ones(B, T + 1, 1, device="cuda") * torch.linspace(-2, 2, B, device="cuda")[:, None, None])
            from chain_control_b200.train.step_fn import merge_x_y
            yhat = ctrl.with_params(theta).unroll(model, merge_x_y).batched(refs)["xpos_of_segment_end"]
            ((yhat - refs["xpos_of_segment_end"]).square().mean()).backward()
            grads[(tag, "closed")] = theta.grad.clone()
    finally:
        ops.set_checkpointing()
    for name in ("deep64", "default", "compact50", "closed"):
        ref = grads[("full", name)]
        for tag in ("auto", "k7"):
            err = float((grads[(tag, name)] - ref).norm() / ref.norm())
            assert err < 2e-5, (name, tag, err)
