"""mirror of cc/core/__init__.py (the parts the hot path touches)."""
from .abstract import AbstractController, AbstractModel, filter_vmap  # noqa: F401
from .module_utils import flatten_module, number_of_params  # noqa: F401
