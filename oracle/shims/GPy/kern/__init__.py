import numpy as np


class _Stationary:
    kind = None

    def __init__(self, input_dim, variance=1.0, lengthscale=None, ARD=False):
        self.input_dim = input_dim
        self.variance = variance
        self.lengthscale = lengthscale
        self.ARD = ARD


class RBF(_Stationary):
    kind = "RBF"


class Matern32(_Stationary):
    kind = "Matern32"


class Matern52(_Stationary):
    kind = "Matern52"
