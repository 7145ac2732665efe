"""MetaBO gym-style environment drop-in (metabo/environment/metabo_gym.py:37-845), single episode.

Same kwargs, attributes and reset/step/seed/set_af_functions/eval_gp/get_state contract as the
reference; it is a B = 1 view of `VecMetaBO` (all numerics on the GPU kernels).  When the AF bound
with `set_af_functions` is the `.af` method of a metabo_b200 NeuralAF, the AF optimisation runs
entirely on the device; any other callable (a foreign numpy AF) is driven through the reference's
host-side hierarchical gridding with the GP posterior still evaluated by the CUDA kernels.
"""
import numpy as np
import torch

from .. import gym_compat as gym
from .. import _lib as L
from .. import ops
from ..engine import uniform_grid
from ..policies.policies import NeuralAF
from .vec_env import VecMetaBO


class MetaBO(gym.Env):
    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self.D = kwargs["D"]
        self.domain = np.stack([np.zeros(self.D), np.ones(self.D)], axis=1)
        if "T" in kwargs:
            self.T_min = self.T_max = kwargs["T"]
        else:
            self.T_min, self.T_max = kwargs["T_min"], kwargs["T_max"]
        assert self.T_min > 0
        assert self.T_min <= self.T_max
        self.T = None
        self.n_init_samples = kwargs["n_init_samples"]
        assert self.n_init_samples <= self.T_max
        self.af = None
        self.do_local_af_opt = kwargs["local_af_opt"]
        self.features = kwargs["features"]
        self.feature_order_eval_envs = ["posterior_mean", "posterior_std", "incumbent", "timestep"]
        self.pass_X_to_pi = kwargs["pass_X_to_pi"]
        self.reward_transformation = kwargs["reward_transformation"]
        self.f_type, self.f_opts = kwargs["f_type"], kwargs["f_opts"]
        self.kernel_variance = kwargs["kernel_variance"]
        self.kernel_lengthscale = kwargs["kernel_lengthscale"]
        self.noise_variance = kwargs["noise_variance"]
        self.use_prior_mean_function = bool(kwargs.get("use_prior_mean_function", False))
        self.rng = None
        self.seeded_with = None
        self._vec = None
        self._pi = None
        self._build(seed=None)
        v = self._vec
        self.discrete_domain = not self.do_local_af_opt
        self.n_features = v.n_features
        self.cardinality_xi_t = v.cardinality_xi_t
        if self.do_local_af_opt:
            self.N_MS, self.k, self.N_LS = v.engine.N_MS, v.engine.k, v.engine.N_LS
            self.multistart_grid = uniform_grid(self.D, v.engine.N_MS_per_dim)
            self.af_max_search_diam = v.engine.diam
        self.observation_space = gym.spaces.Box(low=-100000.0, high=100000.0,
                                                shape=(self.cardinality_xi_t, self.n_features), dtype=np.float32)
        self.action_space = gym.spaces.Discrete(self.cardinality_xi_t)
        self.t = None
        self.X = self.Y = None
        self.xi_t = None
        self.gp_is_empty = True

    def _build(self, seed):
        rng = np.random.RandomState()
        rng.seed(seed)
        self.rng = rng
        mode = "bf16x3" if self._pi is None else self._pi.mlp_mode
        self._vec = VecMetaBO(1, None, lane_rngs=[rng], mlp_mode=mode, **self.kwargs)
        if self._pi is not None:
            self._vec.set_policy(self._pi)

    # -- reference API ---------------------------------------------------------------------------
    def seed(self, seed=None):
        self.seeded_with = seed
        self.rng.seed(seed)

    def set_af_functions(self, af_fun):
        owner = getattr(af_fun, "__self__", None)
        if isinstance(owner, NeuralAF) and not self.pass_X_to_pi:
            self._pi = owner
            self._vec.set_policy(owner)
            self.af = af_fun
        else:
            raise NotImplementedError("only metabo_b200 NeuralAF policies drive the GPU environment; the "
                                      "baseline AFs of the reference (EI, UCB, TAF, ...) are outside the hot path")

    def _sync_views(self):
        v = self._vec
        self.t, self.T = v.t, v.T
        n = v.t
        if n == 0:
            self.X = self.Y = None
        else:
            self.X = v.engine.X[0, :n].cpu().numpy()
            self.Y = v.engine.Y[0, :n].cpu().numpy().reshape(n, 1)
        self.gp_is_empty = n == 0
        idx = torch.arange(v.cardinality_xi_t, dtype=torch.int32, device=v.dev)[None].contiguous()
        self.xi_t = ops.candidates_gather(v.cur["cand"], idx, 1)[0].cpu().numpy()
        self.y_max = float(v.fb.y_max[0])
        self.y_min = float(v.fb.y_min[0])

    def _state(self):
        return self._vec.current_state()[0].cpu().numpy()

    def reset(self):
        if self._pi is None:
            raise RuntimeError("call set_af_functions(pi.af) before reset()")
        self._vec.policy = self._pi
        self._vec.sync_policy()
        self._vec.reset(deterministic=True)
        self._sync_views()
        return self._state()

    def step(self, action):
        v = self._vec
        assert v.t < v.T
        reward, done = v.step(deterministic=True, advance=True, actions=[int(np.asarray(action))])
        if done:   # the reference also evaluates the state after the last step (metabo_gym.py:211-214)
            v.cur = v._round(v.t, True)
        self._sync_views()
        return self._state(), float(reward[0]), bool(done), {}

    def eval_gp(self, X_star):
        """metabo_gym.py:732-745 on the CUDA kernels -> (mean [M], std [M]) float64"""
        assert X_star.shape[1] == self.D
        v = self._vec
        cs = ops.CandidateSet(self.D, tail=torch.as_tensor(np.ascontiguousarray(X_star, dtype=np.float64), device=v.dev))
        v.engine.fit()
        mean, std = ops.gp_posterior(v.engine.gp_state, v.T, cs, 1, kernel=v.kernel, precision=L.GP_FP64)
        return mean[0].cpu().numpy(), std[0].cpu().numpy()

    def get_incumbent(self):
        Y = np.array([self.y_min]) if self.Y is None else self.Y
        return np.max(Y)

    def convert_idx_to_x(self, idx):
        if not isinstance(idx, np.ndarray):
            idx = np.array([idx])
        return self.xi_t[idx, :].reshape(idx.size, self.D)

    def is_terminal(self):
        return self.t == self.T

    def close(self):
        pass
