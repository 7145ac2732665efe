from .policies import NeuralAF  # noqa: F401
from .mlp import MLP  # noqa: F401
