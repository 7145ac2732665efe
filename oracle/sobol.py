"""ORACLE (test infrastructure, never shipped/imported by the product path).

Restatement of ``sobol_seq.i4_sobol_generate`` (sobol-seq 0.1.2, pinned in the
reference's ``environment.yml:16``; the package source is NOT under
/root/reference and cannot be installed here, so this follows the published
Bratley & Fox (ACM TOMS Alg. 659) construction that ``i4_sobol`` ports).

Reference call sites: ``metabo/environment/metabo_gym.py:61,85,91,278,324,371,461``.

Parity status: the first two dimensions are pinned against scipy's
``qmc.Sobol(scramble=False)`` (identical direction numbers for d<=2) in
tests/test_oracle_sobol.py.  Dimensions >=3 use the Bratley-Fox initial
direction numbers restated from the published table ("parity unpinned" for
those; they only determine which candidate points are generated, and every
candidate array is an *input* to both the CUDA path and the oracle).
"""
import numpy as np

# primitive polynomials (binary encoding) and initial direction numbers m_i,
# Bratley & Fox table, dimensions 1..8 (index 0 is the van-der-Corput dimension)
_POLY = [1, 3, 7, 11, 13, 19, 25, 37]
_MINIT = [
    [1],
    [1],
    [1, 1],
    [1, 3, 7],
    [1, 1, 5],
    [1, 3, 1, 1],
    [1, 1, 3, 7],
    [1, 3, 3, 9, 9],
]
_MAXCOL = 30
DIM_MAX = len(_POLY)


def _direction_numbers(dim_num):
    v = np.zeros((dim_num, _MAXCOL), dtype=np.int64)
    v[0, :] = 1
    for i in range(1, dim_num):
        poly = _POLY[i]
        m = poly.bit_length() - 1
        init = _MINIT[i]
        v[i, :m] = init[:m]
        # includ[k] = coefficient of x^(m-1-k) of the polynomial (leading 1 dropped)
        includ = [(poly >> (m - 1 - k)) & 1 for k in range(m)]
        for j in range(m, _MAXCOL):
            newv = int(v[i, j - m])
            l = 1
            for k in range(m):
                l *= 2
                if includ[k]:
                    newv ^= l * int(v[i, j - k - 1])
            v[i, j] = newv
    # scale column j by 2^(maxcol-1-j)
    l = 1
    for j in range(_MAXCOL - 2, -1, -1):
        l *= 2
        v[:, j] *= l
    recipd = 1.0 / (2 * l)
    return v, recipd


def i4_sobol_generate(dim_num, n, skip=1):
    """Same signature/semantics as sobol_seq.i4_sobol_generate: returns [n, dim_num]
    float64, point j is the Gray-code Sobol point with index j+skip (so the all-zero
    point is dropped for the default skip=1 and the first row is 0.5,...,0.5)."""
    dim_num = int(dim_num)
    n = int(n)
    if dim_num < 1 or dim_num > DIM_MAX:
        raise ValueError("oracle sobol supports 1..%d dimensions" % DIM_MAX)
    v, recipd = _direction_numbers(dim_num)
    out = np.zeros((n, dim_num), dtype=np.float64)
    # state before seed `skip`: XOR of v[:, lowzero(s)] for s in 0..skip-1
    lastq = np.zeros(dim_num, dtype=np.int64)
    for s in range(skip):
        lastq ^= v[:, _bit_lo0(s) - 1]
    for j in range(n):
        s = j + skip
        out[j] = lastq * recipd
        lastq = lastq ^ v[:, _bit_lo0(s) - 1]
    return out


def _bit_lo0(n):
    """1-based position of the lowest zero bit of n (i4_bit_lo0)."""
    bit = 1
    while n & 1:
        n >>= 1
        bit += 1
    return bit
