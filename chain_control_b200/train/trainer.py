"""mirror of cc/train/trainer.py: Tracker (:19-64), Logger / DictLogger (:67-87), Callback (:90-101),
ModelControllerTrainer (:104-215).  `jit` is accepted for signature compatibility; there is no tracing compiler
here — the step already runs as a handful of fused CUDA kernels."""
from __future__ import annotations

from collections import deque
from typing import Optional, Union

import numpy as np

from ..utils import to_numpy
from .step_fn import TrainingOptionsController, TrainingOptionsModel, make_step_fn_controller, make_step_fn_model


class Tracker:
    def __init__(self, *tracking_metric: str, moving_average_samples: int = 1, mode="min"):
        self.metric = deque(maxlen=moving_average_samples)
        self.metric_key = "_".join(tracking_metric)
        self._best_model = None
        self._associated_metric = None
        assert mode == "min"
        self._i = 0
        self._i_best = 0

    def report(self, model, logs):
        self._i += 1
        metric = logs[self.metric_key]
        self.metric.append(metric)
        if self._best_model is None:
            self._best_model, self._associated_metric, self._i_best = model, metric, self._i
            return
        current_metric = np.mean(list(self.metric))
        if current_metric < self._associated_metric:
            self._associated_metric, self._best_model, self._i_best = current_metric, model, self._i

    def best_model_or_controller(self):
        return self._best_model

    def best_metric(self):
        return self._associated_metric

    def log_entry(self):
        return {self.metric_key: self.best_metric()}

    def best_metric_i(self):
        return self._i_best


class Logger:
    def log(self, metrics):
        pass

    def get_logs(self):
        pass


class DictLogger(Logger):
    def __init__(self):
        self._logs = None

    def log(self, metrics):
        metrics = {k: np.asarray(v)[None] for k, v in metrics.items()}
        if self._logs is None:
            self._logs = metrics
        else:
            self._logs = {k: np.concatenate([self._logs[k], metrics[k]], 0) for k in metrics}

    def get_logs(self):
        return self._logs


class Callback:
    def __call__(self, model, controller, opt_state, minibatch_state, logs, loggers, trackers) -> None:
        pass


class ModelControllerTrainer:
    """Trains a module, evaluates and logs it."""

    def __init__(self, model, controller=None, model_train_options: Optional[TrainingOptionsModel] = None,
                 controller_train_options: Optional[TrainingOptionsController] = None, loggers=(), trackers=(),
                 callbacks=(), jit: bool = True):
        assert not (model_train_options is not None and controller_train_options is not None), \
            "Specify only `TrainingOptionsModel` or `TrainingOptionsController` depending on what you want to optimise for.."
        self._model = self._controller = None
        if model_train_options is not None:
            assert not isinstance(model, dict), "Multiple models are only suppored for controller training."
            self._model = model
            self._step_fn, self._opt_state, self._minibatch_state = make_step_fn_model(model, model_train_options)
        elif controller_train_options is not None:
            assert controller is not None, "Controller is missing."
            if not isinstance(model, dict):
                model = dict(model0=model)  # registered with model name `model0` (trainer.py:137-143)
            self._controller = controller
            self._step_fn, self._opt_state, self._minibatch_state = make_step_fn_controller(controller, model, controller_train_options)
        self.loggers, self.trackers, self.callbacks = list(loggers), list(trackers), list(callbacks)

    def step(self, i_step: int = 0):
        if self._model is not None:
            self._model, self._opt_state, self._minibatch_state, logs = self._step_fn(self._model, self._opt_state, self._minibatch_state)
        else:
            self._controller, self._opt_state, self._minibatch_state, logs = self._step_fn(self._controller, self._opt_state, self._minibatch_state)
        logs = to_numpy(logs)
        for callback in self.callbacks:
            callback(self._model, self._controller, self._opt_state, self._minibatch_state, logs, self.loggers, self.trackers)
        for logger in self.loggers:
            logger.log(logs)
        for tracker in self.trackers:
            tracker.report(self._model if self._model is not None else self._controller, logs)
        return logs

    def run(self, steps: int = 1):
        for i in range(steps):
            self.step(i)

    def get_tracker_logs(self):
        return [tracker.log_entry() for tracker in self.trackers]

    def get_logs(self):
        return [logger.get_logs() for logger in self.loggers]
