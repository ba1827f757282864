"""The reference's documented workflow (docs/2_collecting_data_from_an_environment, 3_training_a_model, 4_training_a_controller)
on chain_control_b200 -- every stage runs on the GPU through the C ABI:

  1. record trajectories from the two-segment environment      cc_env_two_segments_rollout (one launch for all episodes)
  2. fit a neural-ODE model to them (ModelTrainer)               cc_rollout_fwd / cc_rollout_bwd + cc_mse + cc_opt_step
  3. train a feedback controller against the fitted model        cc_closedloop_fwd / cc_closedloop_bwd
  4. run that controller in closed loop with the ENVIRONMENT     controller kernel and environment kernel alternate

    python docs/workflow_two_segments.py          (needs a B200; ~10 s)

The same sequence is a test: tests/test_two_segments.py::test_gpu_end_to_end_workflow_collect_train_model_train_controller_evaluate."""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from chain_control_b200.env import collect, make_env  # noqa: E402
from chain_control_b200.examples.neural_ode_controller_compact_example import make_neural_ode_controller  # noqa: E402
from chain_control_b200.examples.neural_ode_model_compact_example import make_neural_ode_model  # noqa: E402
from chain_control_b200.train import (DictLogger, EvaluationMetrices, ModelControllerTrainer, SupervisedDataset, Tracker,  # noqa: E402
                                      TrainingOptionsController, TrainingOptionsModel, UnsupervisedDataset, make_dataloader, optim)
from chain_control_b200.utils import rmse  # noqa: E402


def main(model_steps=300, controller_steps=100):
    env = make_env("two_segments_v1", time_limit=10.0, control_timestep=0.01)
    key = env.obs_key
    to = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32), device="cuda")

    # 1. data: 4 Gaussian-process draws + 4 cosines for training, 2 + 2 for validation (cc/utils/high_level/defaults.py:21-26)
    train = collect.sample_feedforward_and_collect(env, seeds_gp=[0, 1, 2, 3], seeds_cos=[4, 8, 12, 2])
    val = collect.sample_feedforward_and_collect(env, seeds_gp=[15, 16], seeds_cos=[2.5, 5.0])
    print(f"collected {train.bs} + {val.bs} episodes of {train.n_timesteps} steps")

    # 2. model: compact neural ODE, 50 states, linear f and g, arctan on the input (notebook 3)
    model = make_neural_ode_model(1, OrderedDict([(key, 1)]), 0.01, state_dim=50, f_depth=0, u_transform=np.arctan)
    data = SupervisedDataset(to(train.action), OrderedDict([(key, to(train.obs[key]))]))
    vdat = (to(val.action), OrderedDict([(key, to(val.obs[key]))]))
    opts = TrainingOptionsModel(make_dataloader(data, key=1, n_minibatches=1), optim.chain(optim.clip_by_global_norm(1.0), optim.adam(3e-3)),
                                metrices=(EvaluationMetrices(data=vdat, metrices=(lambda y, yhat: {"val_rmse": rmse(y, yhat)},)),))
    trainer = ModelControllerTrainer(model, model_train_options=opts, loggers=[DictLogger()], trackers=[Tracker("val_rmse")])
    trainer.run(model_steps)
    logs = trainer.get_logs()[0]
    print(f"model: train_mse {logs['train_mse'][0]:.3e} -> {logs['train_mse'][-1]:.3e}, best val_rmse {np.min(logs['val_rmse']):.3e}")
    fitted = trainer.trackers[0].best_model_or_controller()

    # 3. controller: compact neural ODE with 5 states against the fitted model, step references (notebook 4)
    source = collect.random_steps_source(env, seeds=list(range(16)), max_abs_amplitude=1.0)
    refs = OrderedDict([(key, to(source.get_references()[key]))])
    controller = make_neural_ode_controller(2, 1, 0.01, 5, f_depth=0)
    copts = TrainingOptionsController(make_dataloader(UnsupervisedDataset(refs), key=1, n_minibatches=1),
                                      optim.chain(optim.clip_by_global_norm(1.0), optim.adam(1e-2)))
    ct = ModelControllerTrainer(fitted, controller, controller_train_options=copts, loggers=[DictLogger()], trackers=[Tracker("loss")])
    ct.run(controller_steps)
    clogs = ct.get_logs()[0]
    print(f"controller: loss {clogs['loss'][0]:.3e} -> {clogs['loss'][-1]:.3e}")

    # 4. the trained controller in closed loop with the environment itself, all references at once
    sample = collect.collect_exhaust_source(env, source, ct.trackers[0].best_model_or_controller())
    print(f"closed loop with the environment: mean reward {sample.rew.mean():.3e} over {sample.bs} references x {sample.n_timesteps} steps")


if __name__ == "__main__":
    main()
