import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()

    def contains(self, x):
        x = int(np.asarray(x))
        return 0 <= x < self.n
