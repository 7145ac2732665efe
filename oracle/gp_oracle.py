"""ORACLE (test infrastructure, never shipped/imported by the product path).

CPU restatement of the GP arithmetic the reference obtains from GPy 1.9.8
(``environment.yml:6``; GPy's source is not vendored under /root/reference and
is not installable here).  Call sites restated:

  * ``metabo/environment/metabo_gym.py:549-593``  reset_gp  (kernel ctor, noise, prior mean)
  * ``metabo/environment/metabo_gym.py:610-613``  update_gp (gp.set_XY -> full refit)
  * ``metabo/environment/metabo_gym.py:732-745``  eval_gp   (predict_noiseless, sqrt)

Two implementations live here:

``GPyRoute``   mirrors GPy's *arithmetic order* (published algorithm of
               ``Stationary._scaled_dist``, ``ExactGaussianInference.inference``,
               ``Posterior._raw_predict``): Gram-trick distances, +1e-8 jitter on
               the noise variance, Cholesky, alpha = cho_solve, explicit inverse
               (dpotri), var = Kdiag - sum((Ky^-1 Kx) * Kx), clip at 1e-15.
``truth_gp``   an independent extended-precision (x87 long double, 64-bit
               mantissa) direct-distance / triangular-solve GP.  Needed because the
               GPyRoute itself is only ~5e-4 accurate in sigma on the
               ill-conditioned states the deterministic policy produces
               (duplicate observations, cond ~ 4e9; SURVEY.md section 7 hard part 1).

PARITY UNPINNED at the GPy boundary: the reference ships no golden vectors for
GP mean/std and GPy cannot be executed here, so ``GPyRoute`` is pinned only
against (a) ``truth_gp`` on well-conditioned states (1e-9) and (b) an mpmath
50-digit evaluation on tiny cases (tests/test_oracle_gp.py).
"""
import numpy as np
import scipy.linalg

KERNEL_IDS = {"RBF": 0, "Matern32": 1, "Matern52": 2}
GPY_NOISE_JITTER = 1e-8      # GPy ExactGaussianInference: Ky = K + (sigma_n^2 + 1e-8) I
GPY_VAR_CLIP = 1e-15         # GPy Posterior._raw_predict: var = clip(var, 1e-15, inf)


def _k_of_r(kind, variance, r):
    if kind == "RBF":
        return variance * np.exp(-0.5 * r ** 2)
    if kind == "Matern32":
        return variance * (1.0 + np.sqrt(3.0) * r) * np.exp(-np.sqrt(3.0) * r)
    if kind == "Matern52":
        return variance * (1.0 + np.sqrt(5.0) * r + 5.0 / 3.0 * r ** 2) * np.exp(-np.sqrt(5.0) * r)
    raise ValueError("Unknown kernel function for GP model!")


def gpy_scaled_dist(X, X2, lengthscale):
    """GPy Stationary._scaled_dist with ARD: Gram trick on X/l, clip at 0, sqrt."""
    l = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (X.shape[1],))
    Xs = X / l
    same = X2 is None
    X2s = Xs if same else X2 / l
    X1sq = np.sum(np.square(Xs), 1)
    X2sq = X1sq if same else np.sum(np.square(X2s), 1)
    r2 = -2.0 * np.dot(Xs, X2s.T) + (X1sq[:, None] + X2sq[None, :])
    if same:
        r2[np.diag_indices(X.shape[0])] = 0.0
    r2 = np.clip(r2, 0, np.inf)
    return np.sqrt(r2)


class GPyRoute:
    """fp64 restatement of GPRegression(fixed hyper-parameters).set_XY / predict_noiseless."""

    def __init__(self, kind, variance, lengthscale, noise_variance, D):
        self.kind = kind
        self.variance = float(variance)
        self.lengthscale = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (D,)).copy()
        self.noise_variance = float(noise_variance)
        self.D = D
        self.X = None

    def K(self, X, X2=None):
        return _k_of_r(self.kind, self.variance, gpy_scaled_dist(X, X2, self.lengthscale))

    def set_XY(self, X, Y, prior_mean=0.0):
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
        self.X, self.Y, self.m = X, Y, float(prior_mean)
        Ky = self.K(X)
        Ky[np.diag_indices(X.shape[0])] += self.noise_variance + GPY_NOISE_JITTER
        self.L = scipy.linalg.cholesky(Ky, lower=True)
        self.alpha = scipy.linalg.cho_solve((self.L, True), Y - self.m)
        Ki, info = scipy.linalg.lapack.dpotri(self.L, lower=1)
        assert info == 0
        Ki = np.tril(Ki) + np.tril(Ki, -1).T          # GPy symmetrify
        self.woodbury_inv = Ki

    def predict_noiseless(self, Xnew):
        Xnew = np.asarray(Xnew, dtype=np.float64)
        Kx = self.K(self.X, Xnew)
        mu = np.dot(Kx.T, self.alpha)
        Kxx = np.full(Xnew.shape[0], self.variance)
        var = (Kxx - np.sum(np.dot(self.woodbury_inv.T, Kx) * Kx, 0))[:, None]
        var = np.clip(var, GPY_VAR_CLIP, np.inf)
        mu = mu + self.m
        return mu, var


def gpy_route_posterior(kind, variance, lengthscale, noise_variance, X, Y, Xstar, prior_mean=0.0):
    """eval_gp (metabo_gym.py:732-745) for a non-empty GP: returns (mean[M], std[M]) fp64."""
    gp = GPyRoute(kind, variance, lengthscale, noise_variance, X.shape[1])
    gp.set_XY(X, Y, prior_mean)
    mu, var = gp.predict_noiseless(Xstar)
    return mu[:, 0], np.sqrt(var[:, 0])


def empty_gp_posterior(variance, M):
    """metabo_gym.py:736-738: mean = 0 (not the prior mean), var = kernel_variance."""
    return np.zeros(M), np.sqrt(variance * np.ones(M))


# ----------------------------------------------------------------------------------------------
# extended-precision truth
# ----------------------------------------------------------------------------------------------
def _ld(a):
    return np.asarray(a, dtype=np.longdouble)


def truth_kernel(kind, variance, lengthscale, X, X2):
    """direct (no Gram trick) distances in long double."""
    X, X2 = _ld(X), _ld(X2)
    l = np.broadcast_to(_ld(lengthscale), (X.shape[1],))
    d = (X[:, None, :] - X2[None, :, :]) / l
    r2 = np.sum(d * d, axis=2)
    r = np.sqrt(r2)
    v = _ld(variance)
    if kind == "RBF":
        return v * np.exp(-r2 / 2)
    if kind == "Matern32":
        s3 = np.sqrt(_ld(3))
        return v * (1 + s3 * r) * np.exp(-s3 * r)
    if kind == "Matern52":
        s5 = np.sqrt(_ld(5))
        return v * (1 + s5 * r + _ld(5) / 3 * r2) * np.exp(-s5 * r)
    raise ValueError(kind)


def truth_cholesky(A):
    A = _ld(A).copy()
    n = A.shape[0]
    L = np.zeros_like(A)
    for j in range(n):
        s = A[j, j] - np.dot(L[j, :j], L[j, :j])
        if s <= 0:
            raise np.linalg.LinAlgError("not positive definite")
        L[j, j] = np.sqrt(s)
        if j + 1 < n:
            L[j + 1:, j] = (A[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    return L


def truth_solve_lower(L, B):
    """forward substitution L V = B, B [n, M], long double."""
    n = L.shape[0]
    V = np.zeros_like(B)
    for i in range(n):
        V[i] = (B[i] - L[i, :i] @ V[:i]) / L[i, i]
    return V


def truth_gp(kind, variance, lengthscale, noise_variance, X, Y, Xstar, prior_mean=0.0):
    """Extended-precision GP posterior with the same model as GPy's
    (noise + 1e-8 jitter, variance clip 1e-15).  Returns fp64 (mean[M], std[M])."""
    n = X.shape[0]
    Ky = truth_kernel(kind, variance, lengthscale, X, X)
    Ky[np.diag_indices(n)] = _ld(variance) + _ld(noise_variance) + _ld(GPY_NOISE_JITTER)
    L = truth_cholesky(Ky)
    Kx = truth_kernel(kind, variance, lengthscale, X, Xstar)
    V = truth_solve_lower(L, Kx)
    beta = truth_solve_lower(L, (_ld(Y).reshape(-1, 1) - _ld(prior_mean)))
    mu = (V * beta).sum(axis=0) + _ld(prior_mean)
    var = _ld(variance) - (V * V).sum(axis=0)
    var = np.maximum(var, _ld(GPY_VAR_CLIP))
    return np.asarray(mu, dtype=np.float64), np.asarray(np.sqrt(var), dtype=np.float64)


def condition_number(kind, variance, lengthscale, noise_variance, X):
    K = np.asarray(truth_kernel(kind, variance, lengthscale, X, X), dtype=np.float64)
    K[np.diag_indices(X.shape[0])] += noise_variance + GPY_NOISE_JITTER
    w = np.linalg.eigvalsh(K)
    return float(w[-1] / max(w[0], 1e-300))
