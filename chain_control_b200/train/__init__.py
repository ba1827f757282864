"""mirror of cc/train/__init__.py:1-20"""
from . import optim  # noqa: F401
from .minibatch import SupervisedDataset, SupervisedDatasetWithWeights, UnsupervisedDataset, make_dataloader  # noqa: F401
from .step_fn import (EvaluationMetrices, EvaluationMetricesController, Regularisation, TrainingOptionsController,  # noqa: F401
                      TrainingOptionsModel, l1_l2_regularisers, make_step_fn_controller, make_step_fn_model, merge_x_y)
from .trainer import Callback, DictLogger, ModelControllerTrainer, Tracker  # noqa: F401
