import numpy as np
from oracle.gp_oracle import GPyRoute


class _Noise:
    def __init__(self, v):
        self.variance = v


class GPRegression:
    """set_XY -> full refit; predict_noiseless -> (mu[M,1], var[M,1]) via the GPyRoute restatement."""

    def __init__(self, X, Y, kernel=None, Y_metadata=None, normalizer=None, noise_var=1.0, mean_function=None):
        self.kern = kernel
        self.Gaussian_noise = _Noise(noise_var)
        self.mean_function = mean_function
        self.X, self.Y = X, Y
        self._route = None

    # GPy exposes the kernel as an attribute named after it (policies.py:372-373 `gp.rbf.lengthscale = ...`)
    @property
    def rbf(self):
        assert self.kern.kind == "RBF"
        return self.kern

    @property
    def Mat32(self):
        assert self.kern.kind == "Matern32"
        return self.kern

    @property
    def Mat52(self):
        assert self.kern.kind == "Matern52"
        return self.kern

    def set_XY(self, X, Y):
        self.X, self.Y = X, Y
        k = self.kern
        r = GPyRoute(k.kind, float(np.asarray(k.variance)), np.asarray(k.lengthscale, dtype=np.float64),
                     float(np.asarray(self.Gaussian_noise.variance)), X.shape[1])
        m = float(self.mean_function.f(X)) if self.mean_function is not None else 0.0
        r.set_XY(X, Y, m)
        self._route = r

    def predict_noiseless(self, Xnew):
        if self._route is None:
            self.set_XY(self.X, self.Y)
        # GP._raw_predict adds mean_function.f(Xnew), evaluated lazily on the env's *current* Y
        r = self._route
        if self.mean_function is not None:
            m_now = float(self.mean_function.f(Xnew))
            mu, var = r.predict_noiseless(Xnew)
            return mu - r.m + m_now, var
        return r.predict_noiseless(Xnew)
