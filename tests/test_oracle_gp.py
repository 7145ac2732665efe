"""Pins oracle/gp_oracle.py: GPyRoute (restated GPy arithmetic) vs the long-double truth GP
vs an mpmath 50-digit evaluation.  Reference: metabo_gym.py:549-613, 732-745."""
import mpmath
import numpy as np
import pytest

from oracle import gp_oracle as G


def _mp_gp(kind, var, ls, noise, X, Y, Xs, m=0.0):
    mpmath.mp.dps = 50
    n = len(X)

    def k(a, b):
        r2 = sum(((mpmath.mpf(float(a[d])) - mpmath.mpf(float(b[d]))) / mpmath.mpf(float(ls[d]))) ** 2
                 for d in range(len(a)))
        r = mpmath.sqrt(r2)
        if kind == "RBF":
            return var * mpmath.exp(-r2 / 2)
        if kind == "Matern32":
            return var * (1 + mpmath.sqrt(3) * r) * mpmath.exp(-mpmath.sqrt(3) * r)
        return var * (1 + mpmath.sqrt(5) * r + mpmath.mpf(5) / 3 * r2) * mpmath.exp(-mpmath.sqrt(5) * r)

    K = mpmath.matrix(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = k(X[i], X[j])
        K[i, i] = mpmath.mpf(var) + mpmath.mpf(noise) + mpmath.mpf(1e-8)
    Ki = K ** -1
    mus, sds = [], []
    yv = mpmath.matrix([mpmath.mpf(float(y)) - m for y in Y])
    for xs in Xs:
        kv = mpmath.matrix([k(X[i], xs) for i in range(n)])
        mu = (kv.T * Ki * yv)[0] + m
        v = mpmath.mpf(var) - (kv.T * Ki * kv)[0]
        v = max(v, mpmath.mpf(1e-15))
        mus.append(float(mu))
        sds.append(float(mpmath.sqrt(v)))
    return np.array(mus), np.array(sds)


@pytest.mark.parametrize("kind", ["RBF", "Matern32", "Matern52"])
def test_routes_agree_well_conditioned(kind):
    rng = np.random.RandomState(0)
    X = rng.rand(12, 3)
    Y = rng.randn(12)
    Xs = rng.rand(40, 3)
    ls = np.array([0.3, 0.5, 0.2])
    a = G.gpy_route_posterior(kind, 1.3, ls, 1e-4, X, Y, Xs, 0.2)
    b = G.truth_gp(kind, 1.3, ls, 1e-4, X, Y, Xs, 0.2)
    c = _mp_gp(kind, 1.3, ls, 1e-4, X, Y, Xs[:6], 0.2)
    np.testing.assert_allclose(a[0], b[0], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(a[1], b[1], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(b[0][:6], c[0], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(b[1][:6], c[1], rtol=1e-10, atol=1e-12)


def test_truth_is_accurate_when_ill_conditioned():
    # duplicate observations + tiny noise: the regime the deterministic policy produces
    rng = np.random.RandomState(1)
    X = rng.rand(8, 2)
    X = np.concatenate([X, X[:3]], axis=0)
    Y = np.sin(X.sum(1))
    Xs = np.concatenate([rng.rand(5, 2), X[:2] + 1e-3], axis=0)
    ls = np.array([0.235, 0.578])
    t = G.truth_gp("RBF", 2.0, ls, 8.9e-16, X, Y, Xs)
    m = _mp_gp("RBF", 2.0, ls, 8.9e-16, X, Y, Xs)
    np.testing.assert_allclose(t[0], m[0], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(t[1], m[1], rtol=1e-6, atol=1e-9)
    # and the GPy route is measurably worse there (documented, SURVEY section 7 hard part 1)
    g = G.gpy_route_posterior("RBF", 2.0, ls, 8.9e-16, X, Y, Xs)
    assert np.max(np.abs(g[1] - m[1])) < 5e-3


def test_empty_gp_and_floors():
    mu, sd = G.empty_gp_posterior(0.83, 7)
    assert np.all(mu == 0) and np.allclose(sd, np.sqrt(0.83))
    X = np.array([[0.5, 0.5]])
    mu, sd = G.gpy_route_posterior("RBF", 2.0, [0.2, 0.2], 1e-20, X, np.array([1.0]), X)
    assert sd[0] >= np.sqrt(1e-15)
    assert abs(sd[0] - 1e-4) < 2e-5       # the 1e-8 jitter is the noise floor: sigma ~ 1e-4 at a datum
