"""cc/examples/neural_ode_model.py"""
from .neural_ode import NeuralOdeModel, make_neural_ode_model  # noqa: F401
