"""Device-resident objective functions (fp64 torch ops on the episode axis) so that the BO
episode loop never leaves the GPU.  Same formulas as metabo/environment/objectives.py:
Branin :26-80, Hartmann-3 :165-219 (translated / scaled variants, clipped translations)."""
import math

import torch

_BRA_T_RANGE = ((-0.12, 0.87), (-0.81, 0.18))
_HM3_T_RANGE = ((-0.11, 0.88), (-0.55, 0.44), (-0.85, 0.14))


_CONST = {}


def _const(name, values, like):
    """device-resident constant, created once per (device, dtype): the evaluation must not issue host->device
    copies (it runs inside CUDA-graph captures of the episode loop)"""
    key = (name, like.device, like.dtype)
    if key not in _CONST:
        _CONST[key] = torch.tensor(values, dtype=like.dtype, device=like.device)
    return _CONST[key]


def _clip_t(t, rng):
    lo = _const(("lo", rng), [r[0] for r in rng], t)
    hi = _const(("hi", rng), [r[1] for r in rng], t)
    return torch.minimum(torch.maximum(t, lo), hi)


def bra(x):
    """x [..., 2] in the unit square -> [...] (normalised, negated Branin: maximise)."""
    x1 = x[..., 0] * 15.0 - 5.0
    x2 = x[..., 1] * 15.0
    a, b, c, r, s, t = 1.0, 5.1 / (4 * math.pi ** 2), 5 / math.pi, 6.0, 10.0, 1 / (8 * math.pi)
    f = a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * torch.cos(x1) + s
    return -(1 / 51.44 * (f - 54.44))


def bra_var(x, t, s):
    """x [B,2], t [B,2], s [B] -> [B]"""
    return s * bra(x - _clip_t(t, _BRA_T_RANGE))


def bra_max(t, s):
    pos = torch.tensor([(-math.pi + 5.0) / 15.0, 12.275 / 15.0], dtype=t.dtype, device=t.device)
    return s * bra(pos.expand_as(t)), pos + _clip_t(t, _BRA_T_RANGE)


_HM3_ALPHA = (1.0, 1.2, 3.0, 3.2)
_HM3_A = ((3.0, 10, 30), (0.1, 10, 35), (3.0, 10, 30), (0.1, 10, 35))
_HM3_P = ((3689, 1170, 2673), (4699, 4387, 7470), (1091, 8732, 5547), (381, 5743, 8828))


def hm3(x):
    """x [..., 3] -> [...] (normalised, negated Hartmann-3: maximise)."""
    A = _const("hm3A", _HM3_A, x)
    P = _const("hm3P", [[1e-4 * v for v in row] for row in _HM3_P], x)
    al = _const("hm3alpha", _HM3_ALPHA, x)
    e = (A * (x[..., None, :] - P) ** 2).sum(-1)
    f = -(al * torch.exp(-e)).sum(-1)
    return -(1 / 0.95 * (f - (-0.93)))


def hm3_var(x, t, s):
    return s * hm3(x - _clip_t(t, _HM3_T_RANGE))


def hm3_max(t, s):
    pos = torch.tensor([0.114614, 0.555649, 0.852547], dtype=t.dtype, device=t.device)
    return s * hm3(pos.expand_as(t)), pos + _clip_t(t, _HM3_T_RANGE)
