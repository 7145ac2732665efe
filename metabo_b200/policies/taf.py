"""TAF drop-in (metabo/policies/policies.py:310-468): the Transfer Acquisition Function of Wistuba et al. (2018),
modes "me" (product of experts) and "ranking", with the M source GPs evaluated by ONE batched call of the GP posterior
kernel (`mbo_gp_posterior`, the M source models are the "episodes" of the batch) instead of M sequential
`GPy.predict_noiseless` calls.  Same constructor / `af(state, X_target, model_target)` / `act` contract; it is a foreign AF
for the environment (`pass_X_to_pi=True`: MetaBO.set_af_functions drives it through the host-side gridding).

`datafile`: the reference's pickle layout (policies/taf/generate_taf_data_*.py: D, M, X[i] [N, D], Y[i] [N, 1],
kernel_lengthscale[i], kernel_variance[i], noise_variance[i], use_prior_mean_function[i]) or the dict itself;
`generate_taf_data_branin / _hm3` restate the two shipped generators (fixed GP hyper-parameters, Sobol inputs)."""
import pickle as pkl

import numpy as np
import torch

from .. import _lib as L
from .. import objectives as obj
from .. import ops
from ..sobol import i4_sobol_generate


def _scale(X, dom):
    return X * np.ptp(dom, axis=1) + dom[:, 0]


def generate_taf_data_branin(M, N, datafile=None):
    """policies/taf/generate_taf_data_branin.py:9-52"""
    grid = _scale(i4_sobol_generate(3, M), np.array([[-0.1, 0.1], [-0.1, 0.1], [0.9, 1.1]]))
    X = i4_sobol_generate(2, N)
    data = {"D": 2, "M": M, "X": M * [X], "Y": [], "kernel_lengthscale": M * [[0.235, 0.578]],
            "kernel_variance": M * [2.0], "noise_variance": M * [8.9e-16], "use_prior_mean_function": M * [False]}
    for p in grid:
        y = obj.bra_var(torch.as_tensor(X), torch.as_tensor(p[0:2])[None], torch.as_tensor(p[2])[None])
        data["Y"].append(y.numpy().reshape(-1, 1))
    if datafile is not None:
        with open(datafile, "wb") as f:
            pkl.dump(data, f)
    return data


def generate_taf_data_hm3(M, N, datafile=None):
    """policies/taf/generate_taf_data_hm3.py (same scheme, 3 translations + 1 scaling, Hartmann-3 GP hyper-parameters)"""
    grid = _scale(i4_sobol_generate(4, M), np.array([[-0.1, 0.1]] * 3 + [[0.9, 1.1]]))
    X = i4_sobol_generate(3, N)
    data = {"D": 3, "M": M, "X": M * [X], "Y": [], "kernel_lengthscale": M * [[0.716, 0.298, 0.186]],
            "kernel_variance": M * [0.83], "noise_variance": M * [1.688e-11], "use_prior_mean_function": M * [False]}
    for p in grid:
        y = obj.hm3_var(torch.as_tensor(X), torch.as_tensor(p[0:3])[None], torch.as_tensor(p[3])[None])
        data["Y"].append(y.numpy().reshape(-1, 1))
    if datafile is not None:
        with open(datafile, "wb") as f:
            pkl.dump(data, f)
    return data


class TAF:
    def __init__(self, datafile, mode="me", rho=None):
        self.datafile = datafile
        self.mode = mode
        self.rho = rho
        if self.mode == "me":
            assert self.rho is None
        elif self.mode == "ranking":
            assert self.rho > 0
        else:
            raise ValueError("Unknown TAF-mode!")
        self.generate_source_models()

    def generate_source_models(self):
        """policies.py:327-339 + train_gp :351-375: the M source GPs (RBF-ARD, fixed hyper-parameters) are fitted by one
        `mbo_gp_fit` call; their state blobs stay on the device"""
        if isinstance(self.datafile, dict):
            data = self.datafile
        else:
            with open(self.datafile, "rb") as f:
                data = pkl.load(f)
        self.data = data
        self.D, self.M = data["D"], data["M"]
        if not torch.cuda.is_available():
            raise L.MboError("metabo_b200 TAF needs a CUDA device (no CPU fallback)")
        dev = torch.device("cuda", torch.cuda.current_device())
        ns = [np.asarray(x).shape[0] for x in data["X"]]
        nmax = max(ns)
        X = np.zeros((self.M, nmax, self.D)); Y = np.zeros((self.M, nmax))
        for i in range(self.M):
            X[i, :ns[i]] = np.asarray(data["X"][i], dtype=np.float64)
            Y[i, :ns[i]] = np.asarray(data["Y"][i], dtype=np.float64).reshape(-1)
        pm = set(bool(x) for x in data["use_prior_mean_function"])
        assert len(pm) == 1, "mixed prior-mean settings among the source tasks"
        t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
        ls = np.stack([np.broadcast_to(np.asarray(l, dtype=np.float64), (self.D,)) for l in data["kernel_lengthscale"]])
        self.nmax = nmax
        self.gp_state, status = ops.gp_fit(t(np.array(ns), torch.int32), t(X), t(Y), t(ls),
                                           t(np.asarray(data["kernel_variance"], dtype=np.float64)),
                                           t(np.asarray(data["noise_variance"], dtype=np.float64)), kernel="RBF",
                                           use_prior_mean=pm.pop())
        self.dev = dev
        self.models_source = list(range(self.M))          # the models live in the batched blob

    def _predict_sources(self, xs):
        """means, stds [Mc, M] of every source GP at xs [Mc, D] (policies.py:388-396), one kernel launch"""
        cs = ops.CandidateSet(self.D, tail=torch.as_tensor(np.ascontiguousarray(xs, dtype=np.float64), device=self.dev))
        mean, std = ops.gp_posterior(self.gp_state, self.nmax, cs, self.M, kernel="RBF", precision=L.GP_FP64)
        return mean.t().cpu().numpy(), std.t().cpu().numpy()

    def af(self, state, X_target, model_target):
        from scipy.stats import norm
        state = np.asarray(state)
        mean_idx, std_idx = 0, 1
        means_target, stds_target = state[:, mean_idx], state[:, std_idx]
        incumbents_target = state[:, std_idx + self.D + 1]
        xs = state[:, std_idx + 1:std_idx + 1 + self.D]
        means_source, stds_source = self._predict_sources(xs)
        if self.mode == "me":       # product of experts (policies.py:399-403)
            beta = 1 / (self.M + 1)
            weights = np.concatenate([beta * stds_source ** (-2), (beta * stds_target ** (-2))[:, None]], axis=1)
        else:                       # ranking-based meta-features (policies.py:404-437), vectorised
            t = X_target.shape[0] if X_target is not None else 0
            chi = np.zeros((self.M + 1, t * t))
            if t > 0:
                mu_s, _ = self._predict_sources(X_target)                      # [t, M]
                mu_t, _ = model_target.predict_noiseless(X_target)
                mu = np.concatenate([mu_s, np.asarray(mu_t).reshape(t, 1)], axis=1)      # [t, M + 1]
                with np.errstate(divide="ignore"):
                    val = 1 / (t * (t - 1)) if t > 1 else np.inf
                for k in range(self.M + 1):
                    chi[k] = np.where(mu[:, None, k] > mu[None, :, k], val, 0.0).reshape(-1)   # [i, j] -> j + i * t
            d = np.linalg.norm(chi - chi[self.M][None], axis=1) / self.rho
            w = np.where(d <= 1, 3 / 4 * (1 - d ** 2), 0.0)
            weights = np.tile(w, (xs.shape[0], 1))
        # EI of the target model (policies.py:439-445)
        mask = stds_target != 0.0
        eis_target, zs = np.zeros((means_target.shape[0],)), np.zeros((means_target.shape[0],))
        zs[mask] = (means_target[mask] - incumbents_target[mask]) / stds_target[mask]
        pdf_zs, cdf_zs = norm.pdf(zs), norm.cdf(zs)
        eis_target[mask] = (means_target[mask] - incumbents_target[mask]) * cdf_zs[mask] + stds_target[mask] * pdf_zs[mask]
        # predicted improvements of the source models (policies.py:447-457)
        if X_target is None:
            incumbents_source = np.full(self.M, incumbents_target[0])
        else:
            incumbents_source = np.max(self._predict_sources(X_target)[0], axis=0)
        Is_source = means_source - incumbents_source
        Is_source[Is_source < 0.0] = 0.0
        source_af = np.sum(weights[:, :-1] * Is_source, axis=1)
        target_af = weights[:, -1] * eis_target
        return (source_af + target_af) / np.sum(weights, axis=1)

    def act(self, state, X_target, model_target):
        state = state.numpy() if isinstance(state, torch.Tensor) else state
        tafs = self.af(state, X_target, model_target)
        action = np.random.choice(np.flatnonzero(tafs == tafs.max()))
        return torch.tensor([action], dtype=torch.int64).squeeze(0), torch.tensor([0.0]).squeeze(0)

    def set_requires_grad(self, flag):
        pass

    def reset(self):
        pass
