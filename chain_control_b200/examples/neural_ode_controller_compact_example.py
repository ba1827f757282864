"""cc/examples/neural_ode_controller_compact_example.py"""
from .neural_ode import NeuralOdeController  # noqa: F401
from .neural_ode import make_neural_ode_controller_compact as make_neural_ode_controller  # noqa: F401
