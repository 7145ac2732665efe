"""Objective-function families of the MetaBO experiments, batched over episode lanes.

The *parameters* of every episode's objective are drawn on the host from the lane's own
`numpy.random.RandomState` in exactly the order the reference consumes it
(metabo/environment/metabo_gym.py:231-304 GP / Branin, :360-406 Hartmann-3), because the env RNG is
only touched at episode start; the *evaluation* f(x) runs on the device
(fp64 torch ops over the lane axis), so the BO loop never leaves the GPU.
"""
import math

import numpy as np
import torch

from .. import _lib as L
from .. import objectives as obj
from ..sobol import i4_sobol_generate


def _scale_to_domain(X, domain):
    return X * np.ptp(domain, axis=1) + domain[:, 0]          # util.py:35-37


class FunctionBatch:
    """y = f_b(x_b) for every lane b; y_max / y_min per lane (metabo_gym.py:546-547).
    The parameter tensors listed in `_tensors` are persistent: a new draw is copied INTO them (`copy_from`), so
    device pointers stay valid across episodes (CUDA-graph replays of the episode loop read them)."""
    _tensors = ()

    def copy_from(self, other):
        for name in self._tensors:
            getattr(self, name).copy_(getattr(other, name))

    def eval(self, x):                                         # x [B, D] fp64 device -> [B]
        raise NotImplementedError


class BraninBatch(FunctionBatch):
    D = 2
    family = L.ENV_BRANIN           # analytic family of mbo_env_step (the fused device-resident step)
    _tensors = ("t", "s", "y_max", "x_max", "y_min")

    def __init__(self, t, s, device):
        self.t = torch.as_tensor(np.asarray(t, dtype=np.float64).reshape(-1, 2), device=device)
        self.s = torch.as_tensor(np.asarray(s, dtype=np.float64).reshape(-1), device=device)
        self.y_max, self.x_max = obj.bra_max(self.t, self.s)
        zero = torch.zeros_like(self.t)
        self.y_min = self.s * obj.bra(zero)                    # objectives.py:66-68 min at (0, 0) (untranslated value)

    def eval(self, x):
        return obj.bra_var(x, self.t, self.s)


class Hartmann3Batch(FunctionBatch):
    D = 3
    family = L.ENV_HARTMANN3
    _tensors = ("t", "s", "y_max", "x_max", "y_min")

    def __init__(self, t, s, device):
        self.t = torch.as_tensor(np.asarray(t, dtype=np.float64).reshape(-1, 3), device=device)
        self.s = torch.as_tensor(np.asarray(s, dtype=np.float64).reshape(-1), device=device)
        self.y_max, self.x_max = obj.hm3_max(self.t, self.s)
        mp = torch.tensor([1.0, 1.0, 0.0], dtype=torch.float64, device=device).expand_as(self.t)
        self.y_min = self.s * obj.hm3(mp)                      # objectives.py:203-206

    def eval(self, x):
        return obj.hm3_var(x, self.t, self.s)


class GPSampleBatch(FunctionBatch):
    """Sparse-spectrum GP samples (objectives.py:239-314): f(x) = sqrt(2 sv / m) * theta . cos(W x + b)."""
    _tensors = ("W", "b", "theta", "scale", "y_max", "y_min")

    def __init__(self, W, b, theta, signal_var, min_regret, device):
        self.W = torch.as_tensor(W, device=device)             # [B, m, D]
        self.b = torch.as_tensor(b, device=device)             # [B, m]
        self.theta = torch.as_tensor(theta, device=device)     # [B, m]
        m = self.W.shape[1]
        self.scale = torch.sqrt(2 * torch.as_tensor(signal_var, dtype=torch.float64, device=device) / m)   # [B]
        self.min_regret = float(min_regret)
        self.y_max = None
        self.y_min = None

    def raw(self, x):                                          # x [B, P, D] -> [B, P]
        z = torch.einsum("bmd,bpd->bpm", self.W, x) + self.b[:, None, :]
        return self.scale[:, None] * torch.einsum("bpm,bm->bp", torch.cos(z), self.theta)

    def set_extrema_on(self, grid, chunk=32):
        """metabo_gym.py:254-265: maximum / minimum of the sample over a grid shared by all lanes."""
        B = self.W.shape[0]
        ymax, ymin = [], []
        g = grid[None]
        for i in range(0, B, chunk):
            sl = slice(i, min(B, i + chunk))
            sub = GPSampleBatch.__new__(GPSampleBatch)
            sub.W, sub.b, sub.theta, sub.scale = self.W[sl], self.b[sl], self.theta[sl], self.scale[sl]
            y = sub.raw(g.expand(sub.W.shape[0], -1, -1))
            ymax.append(y.max(dim=1).values)
            ymin.append(y.min(dim=1).values)
        self.y_max, self.y_min = torch.cat(ymax), torch.cat(ymin)

    def eval(self, x):
        return self.raw(x[:, None, :])[:, 0] - self.min_regret   # metabo_gym.py:267


class FunctionSampler:
    """Draws one episode's objective per lane, consuming each lane's RandomState like the reference."""

    def __init__(self, f_type, f_opts, D, device, discrete_grid=None):
        self.f_type, self.f_opts, self.D, self.device = f_type, dict(f_opts), D, device
        self.discrete_grid = discrete_grid
        if f_type not in ("BRA-var", "HM3-var", "GP"):
            raise ValueError("Unknown f_type!")                # metabo_gym.py:544 (families outside the hot path)
        if f_type == "GP" and discrete_grid is None:
            # continuous domain (local_af_opt=True): approximate extrema on a uniform grid with
            # ceil(1000^(1/D)) points per dimension (metabo_gym.py:254-258)
            from ..engine import uniform_grid
            n_dim = int(np.ceil(1000 ** (1 / D)))
            assert n_dim > 3
            self.discrete_grid = torch.as_tensor(uniform_grid(D, n_dim), device=device)
        self._grid = None
        if f_type in ("BRA-var", "HM3-var") and self.f_opts.get("M") is not None:
            nd = 3 if f_type == "BRA-var" else 4
            bt, bs = self.f_opts["bound_translation"], self.f_opts["bound_scaling"]
            dom = np.array([[-bt, bt]] * (nd - 1) + [[1 - bs, 1 + bs]])
            self._grid = _scale_to_domain(i4_sobol_generate(nd, self.f_opts["M"]), dom)

    def _draw_ts(self, rng, nt):
        o = self.f_opts
        if self._grid is not None:
            idx = rng.randint(0, o["M"])      # == rng.choice(M) of the reference (same draw, same state: the legacy choice IS this call), 3x cheaper
            return self._grid[idx, 0:nt], self._grid[idx, nt]
        if "bound_translation" in o:
            t = rng.uniform(low=-o["bound_translation"], high=o["bound_translation"], size=(1, nt))
            s = rng.uniform(low=1 - o["bound_scaling"], high=1 + o["bound_scaling"])
            return t[0], s
        if "translation_min" in o:
            t = rng.uniform(low=o["translation_min"], high=o["translation_max"], size=(1, nt))
            for i in range(nt):
                if rng.uniform(low=0.0, high=1.0) > 0.5:
                    t[0, i] = -t[0, i]
            s = rng.uniform(low=o["scaling_min"], high=o["scaling_max"])
            return t[0], s
        raise ValueError("Missspecified translation/scaling parameters!")

    def draw(self, rngs):
        """rngs: list of per-lane RandomState.  Returns (FunctionBatch, gp_hyper or None) where
        gp_hyper = (lengthscale [B], variance [B], noise [B]) for the GP family (metabo_gym.py:249-252)."""
        B = len(rngs)
        if self.f_type in ("BRA-var", "HM3-var"):
            nt = 2 if self.f_type == "BRA-var" else 3
            if self._grid is not None:       # one integer draw per lane, the grid rows gathered in one go (1 024 lanes: 6.6 -> 2 ms)
                idx = np.fromiter((r.randint(0, self.f_opts["M"]) for r in rngs), dtype=np.int64, count=B)
                t, s = self._grid[idx, 0:nt], self._grid[idx, nt]
            else:
                ts = [self._draw_ts(r, nt) for r in rngs]
                t = np.stack([np.asarray(a) for a, _ in ts])
                s = np.array([float(b) for _, b in ts])
            cls = BraninBatch if self.f_type == "BRA-var" else Hartmann3Batch
            return cls(t, s, self.device), None
        o = self.f_opts
        m = 500
        kernel = o.get("kernel", "RBF")
        W = np.zeros((B, m, self.D)); bb = np.zeros((B, m)); th = np.zeros((B, m))
        ls = np.zeros(B); sv = np.zeros(B)
        g = np.random.RandomState(0)
        for i, r in enumerate(rngs):
            seed = r.randint(100000)
            ls[i] = r.uniform(low=o["lengthscale_low"], high=o["lengthscale_high"])
            r.uniform(low=o["noise_var_low"], high=o["noise_var_high"])          # noise_var: only scales the prior
            sv[i] = r.uniform(low=o["signal_var_low"], high=o["signal_var_high"])
            g.seed(seed)            # (one generator object re-seeded per lane: constructing a RandomState costs 120 us, 40 % of a draw)
            if kernel == "RBF":
                w = g.randn(m, self.D) / ls[i]
            elif kernel == "Matern32":
                w = g.standard_t(3, (m, self.D)) / ls[i]
            elif kernel == "Matern52":
                w = g.standard_t(5, (m, self.D)) / ls[i]
            else:
                raise ValueError("Unknown kernel function for GP model!")
            W[i] = w
            bb[i] = g.uniform(0, 2 * np.pi, size=m)
            # train() on the empty set: theta ~ N(0, (1 + 1e-10) I)   (objectives.py:269-285)
            th[i] = math.sqrt(1.0 + 1e-10) * g.randn(m, 1)[:, 0]
        fb = GPSampleBatch(W, bb, th, sv, o["min_regret"], self.device)
        fb.set_extrema_on(self.discrete_grid)
        return fb, (ls, sv, np.full(B, 1e-20))
