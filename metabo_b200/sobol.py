"""Sobol candidate tables with sobol_seq.i4_sobol_generate semantics (reference call sites
metabo/environment/metabo_gym.py:61,85,91): Gray-code order, Bratley-Fox direction numbers, the
all-zero point skipped, so row 0 is (0.5, ..., 0.5).  Vectorised: the j-th point is the XOR of the
direction vectors selected by the bits of gray(j + skip)."""
import numpy as np

_POLY = (1, 3, 7, 11, 13, 19, 25, 37)
_M_INIT = ((1,), (1,), (1, 1), (1, 3, 7), (1, 1, 5), (1, 3, 1, 1), (1, 1, 3, 7), (1, 3, 3, 9, 9))
_BITS = 30
MAX_DIM = len(_POLY)


def _directions(dim):
    V = np.zeros((dim, _BITS), dtype=np.int64)
    V[0] = 1 << np.arange(_BITS - 1, -1, -1)
    for d in range(1, dim):
        p = _POLY[d]
        deg = p.bit_length() - 1
        m = list(_M_INIT[d][:deg])
        coef = [(p >> (deg - 1 - i)) & 1 for i in range(deg)]        # a_1 .. a_deg (a_deg == 1)
        while len(m) < _BITS:
            j = len(m)
            v = m[j - deg]
            for i in range(deg):
                if coef[i]:
                    v ^= m[j - 1 - i] << (i + 1)
            m.append(v)
        V[d] = [m[j] << (_BITS - 1 - j) for j in range(_BITS)]
    return V


def i4_sobol_generate(dim_num, n, skip=1):
    dim_num, n = int(dim_num), int(n)
    if not 1 <= dim_num <= MAX_DIM:
        raise ValueError("sobol: dimension must be in 1..%d" % MAX_DIM)
    V = _directions(dim_num)
    idx = np.arange(skip, skip + n, dtype=np.int64)
    gray = idx ^ (idx >> 1)
    q = np.zeros((n, dim_num), dtype=np.int64)
    for bit in range(_BITS):
        sel = ((gray >> bit) & 1).astype(bool)
        if sel.any():
            q[sel] ^= V[:, bit]
    return q.astype(np.float64) / float(1 << _BITS)
