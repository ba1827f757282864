"""cc/examples/neural_ode_controller.py"""
from .neural_ode import NeuralOdeController, make_neural_ode_controller  # noqa: F401
