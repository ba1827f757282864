"""CPU tests of the host-side mirrors that need no GPU: optimiser (vs the optax formulas), dataloader state
machine, tracker / logger bookkeeping, callable classification."""
import numpy as np
import pytest
import torch

from chain_control_b200 import arch as A
from chain_control_b200.examples.neural_ode import classify_activation, classify_u_transform, init_flat_params
from chain_control_b200.train import (DictLogger, SupervisedDataset, Tracker, UnsupervisedDataset, make_dataloader, optim)
from chain_control_b200.utils import batch_concat, l1_norm, l2_norm, mse, rmse, weighted_mse


def test_adam_matches_optax_formula():
    opt = optim.chain(optim.clip_by_global_norm(1.0), optim.adam(3e-3))
    p = torch.tensor([0.5, -1.0, 2.0])
    st = opt.init(p)
    m = np.zeros(3); v = np.zeros(3); pn = p.numpy().astype(np.float64)
    for t in range(1, 6):
        g = torch.tensor([3.0 * t, -0.1, 0.4 * t])
        upd, st = opt.update(g, st, p)
        p = optim.apply_updates(p, upd)
        gn = g.numpy().astype(np.float64)
        gn = gn * min(1.0, 1.0 / np.linalg.norm(gn))
        m = 0.9 * m + 0.1 * gn; v = 0.999 * v + 0.001 * gn * gn
        pn = pn - 3e-3 * (m / (1 - 0.9 ** t)) / (np.sqrt(v / (1 - 0.999 ** t)) + 1e-8)
        np.testing.assert_allclose(p.numpy(), pn, rtol=2e-6)
    np.testing.assert_allclose(optim.clip(0.5).update(torch.tensor([2.0, -3.0, 0.1]), (), None)[0].numpy(), [0.5, -0.5, 0.1], rtol=1e-6)


def test_dataloader_epochs_cover_every_sample_once():
    x = np.arange(24, dtype=np.float32).reshape(24, 1, 1)
    dl = make_dataloader(SupervisedDataset(x, 2 * x), key=2, n_minibatches=4, device="cpu")
    st = dl.minibatch_state
    assert (st.bs, st.n_minibatches, st.minibatch_size) == (24, 4, 6)
    for epoch in range(2):
        seen = []
        for _ in range(4):
            st, mb = dl.update_fn(st)
            assert tuple(mb.inputs.shape) == (6, 1, 1) and tuple(mb.weights.shape) == (6,)
            np.testing.assert_array_equal(mb.targets.numpy(), 2 * mb.inputs.numpy())
            seen += mb.inputs.flatten().tolist()
        assert sorted(seen) == list(range(24))
    with pytest.raises(ValueError):
        make_dataloader(np.zeros(3))
    dl = make_dataloader(UnsupervisedDataset({"xpos": x}), n_minibatches=1, device="cpu")
    st, mb = dl.update_fn(dl.minibatch_state)
    assert tuple(mb.refss["xpos"].shape) == (24, 1, 1)


def test_tracker_and_logger():
    tr = Tracker("m2", "train_mse")
    assert tr.metric_key == "m2_train_mse"
    for i, v in enumerate([3.0, 1.0, 2.0]):
        tr.report(f"model{i}", {"m2_train_mse": v})
    assert tr.best_model_or_controller() == "model1" and tr.best_metric() == 1.0 and tr.best_metric_i() == 2
    lg = DictLogger()
    lg.log({"train_loss": np.float32(1.0)}); lg.log({"train_loss": np.float32(0.5)})
    np.testing.assert_array_equal(lg.get_logs()["train_loss"], [1.0, 0.5])


def test_metrics_and_tree_helpers():
    y = {"a": torch.ones(2, 3, 1), "b": torch.zeros(2, 3, 2)}
    yh = {"a": torch.zeros(2, 3, 1), "b": torch.zeros(2, 3, 2)}
    assert tuple(batch_concat(y, 2).shape) == (2, 3, 3)
    assert abs(float(mse(y, yh)) - 1 / 3) < 1e-7 and abs(float(rmse(y, yh)) - np.sqrt(1 / 3)) < 1e-7
    assert abs(float(weighted_mse(y, yh, torch.ones(2))) - 2 / 3) < 1e-6  # B x the plain mse when w = 1
    v = torch.tensor([3.0, -4.0])
    assert float(l2_norm(v)) == 5.0 and float(l1_norm(v)) == 7.0


def test_callable_classification_and_init():
    assert classify_activation(np.tanh) == A.ACT_TANH
    assert classify_activation(lambda x: x) == A.ACT_IDENTITY
    assert classify_activation(torch.relu) == A.ACT_RELU
    assert classify_u_transform(np.arctan) == (A.UT_ARCTAN, 1.0)
    kind, k = classify_u_transform(lambda u: 6 * np.tanh(u / 6))
    assert kind == A.UT_KTANH and abs(k - 6) < 1e-6
    with pytest.raises(NotImplementedError):
        classify_activation(np.sin)
    spec = A.make_node_spec(50, 1, 1, 0.01, f_depth=0)
    th = init_flat_params(spec, 1)
    assert th.numel() == 2651 and float(th[:2550].abs().max()) <= 1 / np.sqrt(51) + 1e-7


def test_linear_model_and_pid_parameter_layout_match_reference_fixtures():
    """LinearModel / PIDController host mirrors: the kernel-order `theta` they derive from the reference-order flat vector
    equals what the reference's own modules gave (tests/golden/*.npz: theta = theta_ref_order[perm]); CPU tensors, no kernel call"""
    import os

    import numpy as np
    import torch

    from chain_control_b200.examples.linear_model import make_linear_model
    from chain_control_b200.examples.pid_controller import make_pid_controller

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name, (S, I, O, ct) in {"open_linear_ct_s3": (3, 1, 1, True), "open_linear_dt_s6_i2_o2_x0": (6, 2, 2, False)}.items():
        g = np.load(os.path.join(gold, name + ".npz"))
        m = make_linear_model(I, O, 0.01, S, continuous_time=ct, device="cpu")
        p = torch.tensor(g["theta_ref_order"], dtype=torch.float32, requires_grad=True)
        m = m.with_params(p)
        assert m.spec.n_params() == g["theta"].size and m.spec.g_feedthrough_u and m.spec.readout_pre_step
        np.testing.assert_array_equal(m.theta.detach().numpy(), g["theta"].astype(np.float32))
        w = torch.arange(1.0, 1.0 + g["theta"].size)
        (gr,) = torch.autograd.grad((m.theta * w).sum(), p)  # the scatter back to the reference order
        np.testing.assert_array_equal(gr.numpy()[g["perm"]], w.numpy())
        z = m.replace(D=np.zeros((O, I))).frozen()  # the CLI's LTI systems: D = 0 -> no feed-through -> fast kernel families
        assert not z.spec.g_feedthrough_u and z.flat_params().numel() == S * S + S * I + O * S
    g = np.load(os.path.join(gold, "closed_pid__linear.npz"))
    d, i, p = g["pid_gains_dip"]
    c = make_pid_controller(p, i, d, float(g["pid_dt"]), device="cpu")
    np.testing.assert_allclose(c.theta.numpy(), g["theta_c"], rtol=1e-6)
    np.testing.assert_allclose(c.flat_params().numpy(), g["theta_c_ref_order"], rtol=1e-7)
    c2 = make_pid_controller(p, i, d, 0.01, d_gain_trainable=False, device="cpu",
                             transform_params=lambda q: {"p_gain": q["p_gain"], "i_gain": q["i_gain"], "d_gain": q["p_gain"] * 0.1})
    assert c2.flat_params().numel() == 2 and abs(float(c2.theta[8]) + 0.1 * p / 0.01) < 1e-3


def test_numa_binding_is_a_safe_no_op_without_topology(monkeypatch):
    """parallel.bind_host_to_gpu_numa never raises and never empties the affinity mask: without a GPU / sysfs NUMA files it
    returns None and leaves the process where it is; CC_B200_NUMA_BIND=0 disables it"""
    import os

    from chain_control_b200 import parallel

    before = os.sched_getaffinity(0)
    assert parallel.bind_host_to_gpu_numa(0) is None or os.sched_getaffinity(0) <= before
    assert len(os.sched_getaffinity(0)) >= 1
    monkeypatch.setenv("CC_B200_NUMA_BIND", "0")
    assert parallel.bind_host_to_gpu_numa(0) is None
    os.sched_setaffinity(0, before)
