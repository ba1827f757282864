"""mirror of cc/train/step_fn.py: the train-step builders around the accelerated hot path.

  make_step_fn_model       step_fn.py:132-184   loss(vmap(model.unroll)(inputs)) + regularisers, grad, optimiser update
  make_step_fn_controller  step_fn.py:244-327   closed-loop loss per frozen model, mean over models, grad w.r.t. controller
The loss / regulariser / metric functions stay user-pluggable Python callables on torch tensors (the reference's
are jnp callables); gradients flow through the custom VJP of chain_control_b200.ops (the BPTT kernels)."""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Callable, Optional

import torch

from ..core.module_utils import flatten_module
from ..utils import batch_concat, l1_norm, l2_norm, mse, tree_map
from .minibatch import Dataloader, MiniBatchState, SupervisedDatasetWithWeights, UnsupervisedDataset
from .. import parallel
from .optim import GradientTransformation, apply_updates


def default_loss_fn_reduce_along_models(log_of_loss_values):
    flat_logs = batch_concat(log_of_loss_values, 0)
    return {"loss_without_regu": torch.mean(flat_logs)}


@dataclass
class EvaluationMetrices:
    data: tuple
    metrices: tuple


@dataclass
class Regularisation:
    prefactor: float
    reduce_weights: Callable


def l1_l2_regularisers(lambda_l1, lambda_l2):
    return (Regularisation(lambda_l1, lambda v: {"l1_norm": l1_norm(v)}),
            Regularisation(lambda_l2, lambda v: {"l2_norm": l2_norm(v)}))


@dataclass
class TrainingOptionsModel:
    training_data: Dataloader
    optimizer: GradientTransformation
    metrices: tuple = tuple()
    regularisers: tuple = tuple()
    loss_fn: Callable = lambda y, yhat, weights: dict(train_mse=mse(y, yhat))


def compute_regularisers(model_or_controller, regularisers):
    if len(regularisers) == 0:
        return 0.0, {}
    flat_module = flatten_module(model_or_controller)
    regu_value, log_of_regus = 0.0, {}
    for regu in regularisers:
        this_value = regu.reduce_weights(flat_module)
        log_of_regus.update(this_value)
        regu_value = regu_value + regu.prefactor * batch_concat(this_value, 0)
    return regu_value, log_of_regus


def eval_metrices(model, metrices):
    log_of_metrices = {}
    with torch.no_grad():
        for metric in metrices:
            inputs, targets = metric.data
            preds = model.unroll_batched(inputs)  # eqx.filter_vmap(model.unroll)(inputs), step_fn.py:124
            for metric_fn in metric.metrices:
                log_of_metrices.update(metric_fn(preds, targets))
    return log_of_metrices


def _detach_logs(logs):
    return {k: (v.detach() if isinstance(v, torch.Tensor) else torch.as_tensor(v)) for k, v in logs.items()}


def _mean_logs(minibatched_logs):
    keys = minibatched_logs[0].keys()
    return {k: torch.stack([torch.as_tensor(l[k], dtype=torch.float32).reshape(()) for l in minibatched_logs]).mean() for k in keys}


def make_step_fn_model(model, options: TrainingOptionsModel):
    optimizer = options.optimizer

    def loss_fn_model(model, inputs, targets, weights):
        logs = {}
        preds = model.unroll_batched(inputs)  # step_fn.py:136 -- the hot path
        loss_value = options.loss_fn(targets, preds, weights)
        logs.update(loss_value)
        regu_value, log_of_regus = compute_regularisers(model, options.regularisers)
        logs.update(log_of_regus)
        loss_value = batch_concat(loss_value, 0) + regu_value
        assert loss_value.ndim in (0, 1)
        return loss_value.squeeze(), logs

    def step_fn_model(model, opt_state, minibatch_state: MiniBatchState):
        minibatched_logs = []
        for _ in range(minibatch_state.n_minibatches):
            minibatch_state, mb = options.training_data.update_fn(minibatch_state)
            assert isinstance(mb, SupervisedDatasetWithWeights)
            theta = model.flat_params().detach().requires_grad_(True)
            m = model.with_params(theta)
            # one process per GPU (torch.distributed initialised): every rank takes its slice of the minibatch
            (inputs, targets, weights), frac = parallel.dp_shard((mb.inputs, mb.targets, mb.weights))
            loss, logs = loss_fn_model(m, inputs, targets, weights)
            (grads,) = torch.autograd.grad(loss, theta)
            logs = _detach_logs(logs)
            logs["train_loss"] = loss.detach()
            grads, logs = parallel.dp_combine(grads, logs, frac)
            minibatched_logs.append(logs)
            updates, opt_state = optimizer.update(grads, opt_state, theta.detach())
            model = model.with_params(apply_updates(theta.detach(), updates))
        logs = _mean_logs(minibatched_logs)
        logs.update(_detach_logs(eval_metrices(model, options.metrices)))
        return model, opt_state, minibatch_state, logs

    opt_state = optimizer.init(model.flat_params())
    return step_fn_model, opt_state, options.training_data.minibatch_state


def merge_x_y(x, y):
    """ref, obs"""
    d = OrderedDict()
    d["ref"] = x
    d["obs"] = y
    return d


@dataclass
class EvaluationMetricesController:
    refs: object
    model: object
    metrices: tuple


@dataclass
class TrainingOptionsController:
    refss: Dataloader
    optimizer: GradientTransformation
    regularisers: tuple = tuple()
    merge_x_y: Callable = merge_x_y
    loss_fn: Callable = lambda y, yhat: dict(train_mse=mse(y, yhat))
    loss_fn_reduce_along_models: Callable = default_loss_fn_reduce_along_models
    noise_scale: Optional[float] = None
    metrices: tuple = tuple()


def eval_metrices_controller(controller, metrices, noise_scale, key):
    log_of_metrices = {}
    with torch.no_grad():
        for metric in metrices:
            refss = metric.refs
            assert batch_concat(refss, 2).ndim == 3, "Must be batched so (B, T, F)"
            refsshat = controller.unroll(metric.model, merge_x_y, noise_scale).batched(refss, key)
            for metric_fn in metric.metrices:
                log_of_metrices.update(metric_fn(refsshat, refss))
    return log_of_metrices


_STREAMS = {}


def _model_streams(n, device):
    """a cached pool of side streams, one per model of the multi-model loss"""
    import os

    if os.environ.get("CC_B200_MODEL_STREAMS", "1") == "0":
        return None
    pool = _STREAMS.setdefault(str(device), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device))
    return pool[:n]


class _null_ctx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def make_step_fn_controller(controller, models: dict, options: TrainingOptionsController):
    optimizer = options.optimizer

    def loss_fn_controller(controller, refss, key):
        logs, log_of_loss_values = {}, OrderedDict()
        # step_fn.py:256-261 loops over the models inside one jitted function (XLA is free to overlap them).  Here every
        # model's closed loop (rollout, loss and -- because autograd replays on the forward stream -- its BPTT) runs on its own
        # CUDA stream: at the reference's sizes (tens of trajectories) each launch is one latency-bound wave, so K models
        # overlap almost perfectly.  CC_B200_MODEL_STREAMS=0 serialises them (tests compare both).
        theta_c = controller.theta  # (composed controllers build it with torch ops: once, on the current stream)
        streams = _model_streams(len(models), theta_c.device) if (len(models) > 1 and theta_c.is_cuda) else None
        cur = torch.cuda.current_stream(theta_c.device) if streams else None
        if streams:
            start = torch.cuda.Event()
            start.record(cur)
        for k, (model_name, model) in enumerate(models.items()):
            if streams:
                streams[k].wait_event(start)
            with (torch.cuda.stream(streams[k]) if streams else _null_ctx()):
                refsshat = controller.unroll(model, options.merge_x_y, options.noise_scale).batched(refss, key)  # step_fn.py:257-259
                log_of_loss_values[model_name] = options.loss_fn(refss, refsshat)
            if streams:
                for v in log_of_loss_values[model_name].values():
                    if torch.is_tensor(v):
                        v.record_stream(cur)  # consumed by the reduction below on the main stream
                done = torch.cuda.Event()
                done.record(streams[k])
                cur.wait_event(done)
        regu_value, log_of_regus = compute_regularisers(controller, options.regularisers)
        logs.update(log_of_regus)
        loss_value = options.loss_fn_reduce_along_models(log_of_loss_values)
        logs.update(loss_value)
        for model_name, name_and_value in log_of_loss_values.items():
            for loss_name, v in name_and_value.items():
                logs[model_name + "_" + loss_name] = v
        loss_value = batch_concat(loss_value, 0) + regu_value
        logs["loss"] = loss_value
        assert loss_value.ndim in (0, 1)
        return loss_value.squeeze(), logs

    def step_fn_controller(controller, opt_state, minibatch_state: MiniBatchState):
        minibatched_logs = []
        key = minibatch_state.key
        for _ in range(minibatch_state.n_minibatches):
            minibatch_state, batch = options.refss.update_fn(minibatch_state)
            assert isinstance(batch, UnsupervisedDataset)
            key = key + 1  # "steal key for simulation of noise in closed loop" (step_fn.py:295-297)
            theta = controller.flat_params().detach().requires_grad_(True)
            c = controller.with_params(theta)
            refss, frac = parallel.dp_shard(batch.refss)
            loss, logs = loss_fn_controller(c, refss, key + 1000003 * parallel.world()[0])  # (independent noise per shard)
            (grads,) = torch.autograd.grad(loss, theta)
            logs = _detach_logs(logs)
            logs["train_loss"] = loss.detach()
            grads, logs = parallel.dp_combine(grads, logs, frac)
            minibatched_logs.append(logs)
            updates, opt_state = optimizer.update(grads, opt_state)
            controller = controller.with_params(apply_updates(theta.detach(), updates))
        logs = _mean_logs(minibatched_logs)
        logs.update(_detach_logs(eval_metrices_controller(controller, options.metrices, options.noise_scale, key + 1)))
        minibatch_state = minibatch_state._replace(key=key + 1)
        return controller, opt_state, minibatch_state, logs

    opt_state = optimizer.init(controller.flat_params())
    return step_fn_controller, opt_state, options.refss.minibatch_state
