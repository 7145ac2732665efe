"""Pins oracle/sobol.py (restated sobol_seq.i4_sobol_generate; reference call sites
metabo_gym.py:61,85,91) where an independent implementation exists."""
import numpy as np
from scipy.stats import qmc

from oracle.sobol import i4_sobol_generate


def test_first_point_and_shape():
    p = i4_sobol_generate(3, 5)
    assert p.shape == (5, 3)
    assert np.all(p[0] == 0.5)          # skip=1 drops the all-zero point
    assert np.all((p >= 0) & (p < 1))


def test_first_two_dims_match_scipy_gray_order():
    n = 1024
    ours = i4_sobol_generate(2, n)
    ref = qmc.Sobol(d=2, scramble=False).random(n + 1)[1:]   # scipy includes the zero point
    assert np.array_equal(ours, ref)


def test_stratification_all_dims():
    # a (t,m,s)-net property every correct direction-number table satisfies per dimension:
    # the first 2^k points (incl. the dropped zero point) hit each of the 2^k bins exactly once
    for d in range(1, 9):
        pts = np.concatenate([np.zeros((1, d)), i4_sobol_generate(d, 255)], axis=0)
        for j in range(d):
            assert len(np.unique(np.floor(pts[:, j] * 256))) == 256
