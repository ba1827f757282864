"""cc/examples/neural_ode_model_compact_example.py"""
from .neural_ode import NeuralOdeModel  # noqa: F401
from .neural_ode import make_neural_ode_model_compact as make_neural_ode_model  # noqa: F401
