"""mirror of cc/core/module_utils.py:20-28.  The scan (filter_scan_module, :8-17) is what the CUDA kernels replace."""


def flatten_module(model_or_controller):
    """flat trainable-parameter vector in the reference's flatten_module order
    (f.layers[i].weight, .bias, ..., g....); here the modules store exactly this vector."""
    return model_or_controller.flat_params()


def number_of_params(model_or_controller) -> int:
    return int(flatten_module(model_or_controller).numel())
