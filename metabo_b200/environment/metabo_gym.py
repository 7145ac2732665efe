"""MetaBO gym-style environment drop-in (metabo/environment/metabo_gym.py:37-845), single episode.

Same kwargs, attributes and reset/step/seed/set_af_functions/eval_gp/get_state contract as the
reference; it is a B = 1 view of `VecMetaBO` (all numerics on the GPU kernels).

`set_af_functions(af_fun)` (metabo_gym.py:168-174):
  * `af_fun` is the `.af` method of a metabo_b200 NeuralAF or closed-form baseline (EI / UCB / PI): the
    AF optimisation (metabo_gym.py:615-622, 706-721) runs entirely on the device (AFEngine.af_round);
  * any other callable -- a foreign numpy AF `af(state) -> [M]`, or with `pass_X_to_pi` `af(state, X, gp)`
    (TAF, policies.py:310-468) -- is driven through the reference's host-side hierarchical gridding
    (`get_af_maxima`: multistart grid -> top-k -> local Sobol grids -> top-k), exactly as the reference
    does it, with `get_state` / `eval_gp` (the GP refit and posterior) still on the CUDA kernels.
"""
import numpy as np
import torch

from .. import gym_compat as gym
from .. import _lib as L
from .. import ops
from ..engine import uniform_grid
from ..policies.policies import NeuralAF
from ..sobol import i4_sobol_generate
from .vec_env import VecMetaBO


class _GPView:
    """what `af(state, X, gp)` callbacks (pass_X_to_pi, metabo_gym.py:173) get as `gp`: the slice of the GPy model API
    the reference's AFs use -- `predict_noiseless(X) -> (mean [M,1], var [M,1])` (policies.py:436-441), `.X`, `.Y`."""

    def __init__(self, env):
        self._env = env

    @property
    def X(self):
        return self._env.X

    @property
    def Y(self):
        return self._env.Y

    def predict_noiseless(self, X):
        mean, std = self._env.eval_gp(np.asarray(X, dtype=np.float64).reshape(-1, self._env.D))
        return mean.reshape(-1, 1), (std ** 2).reshape(-1, 1)


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
        self._host_af = None            # foreign AF callable: host-side gridding (see module docstring)
        self.gp = _GPView(self)
        self._build(seed=None)
        v = self._vec
        self.discrete_domain = not self.do_local_af_opt
        self.n_features = v.n_features
        self.cardinality_xi_t = v.cardinality_xi_t
        if self.do_local_af_opt:
            self.N_MS, self.k, self.N_LS = v.engine.N_MS, v.engine.k, v.engine.N_LS
            self.multistart_grid = uniform_grid(self.D, v.engine.N_MS_per_dim)
            self.af_max_search_diam = v.engine.diam
            self.local_search_grid = i4_sobol_generate(self.D, self.N_LS)
            self.cardinality_xi_local_t = self.k
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
        on_device = (isinstance(owner, NeuralAF) or hasattr(owner, "engine_spec")) and not self.pass_X_to_pi
        if on_device:
            self._pi, self._host_af = owner, None
            self._vec.set_policy(owner)
            self.af = af_fun
        else:
            # foreign AF: the reference's callback contract (metabo_gym.py:168-174)
            self._pi, self._host_af = None, af_fun
            if not self.pass_X_to_pi:
                self.af = af_fun
            else:
                self.af = lambda state: af_fun(state, self.X, self.gp)

    def _sync_views(self):
        v = self._vec
        self._fit_t = None
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
        if self._host_af is not None:
            return self._host_reset()
        if self._pi is None:
            raise RuntimeError("call set_af_functions(pi.af) before reset()")
        self._vec.policy = self._pi
        self._vec.sync_policy()
        self._vec.reset(deterministic=True)
        self._sync_views()
        return self._state()

    def step(self, action):
        if self._host_af is not None:
            return self._host_step(action)
        v = self._vec
        assert v.t < v.T
        reward, done = v.step(deterministic=True, advance=True, actions=[int(np.asarray(action))])
        if done:   # the reference also evaluates the state after the last step (metabo_gym.py:211-214)
            v.cur = v._round(v.t, True)
        self._sync_views()
        return self._state(), float(reward[0]), bool(done), {}

    # -- foreign-AF path: host-side gridding as the reference, GP on the kernels -------------------
    def _host_views(self):
        v = self._vec
        self.t, self.T = v.t, v.T
        n = v.t
        self.X = v.engine.X[0, :n].cpu().numpy() if n else None
        self.Y = v.engine.Y[0, :n].cpu().numpy().reshape(n, 1) if n else None
        self.gp_is_empty = n == 0
        self.y_max = float(v.fb.y_max[0])
        self.y_min = float(v.fb.y_min[0])
        self._fit_t = None

    def _host_reset(self):
        """reset (metabo_gym.py:176-193): new objective, empty GP, optimize_AF with the callback, get_state(xi_t)"""
        v = self._vec
        v.t = 0
        fb, hyper = v.sampler.draw(v.rngs)
        if v.fb is None:
            v.fb = fb
        else:
            v.fb.copy_from(fb)
        eng = v.engine
        if hyper is not None:
            eng.set_hyperparameters(np.repeat(hyper[0][:, None], self.D, axis=1), hyper[1], hyper[2])
        else:
            eng.set_hyperparameters(v.fixed_hyper[0], v.fixed_hyper[1], v.fixed_hyper[2])
        eng.n.zero_()
        eng.n_active_max = 0
        v.best.fill_(-float("inf"))
        self._host_views()
        self.optimize_AF()
        return self.get_state(self.xi_t)

    def _host_step(self, action):
        """step (metabo_gym.py:195-217) with the action resolved on the host-side xi_t"""
        v = self._vec
        assert v.t < v.T
        if v.t < v.n_init_samples:
            x = v.initial_design[v.t].cpu().numpy().reshape(1, self.D)
        else:
            x = self.convert_idx_to_x(int(np.asarray(action)))
        cs = ops.CandidateSet(self.D, tail=torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=v.dev))
        out = v._step_device(v.t, True, False, {"cand": cs}, [0])
        v.t += 1
        v.engine.n_active_max = v.t
        self._host_views()
        reward = float(out["reward"][0])
        done = v.t == v.T
        self.optimize_AF()
        return self.get_state(self.xi_t), reward, bool(done), {}

    def get_random_sampling_reward(self):
        """metabo_gym.py:754-773: one episode of uniformly random queries (the "Random" baseline of evaluate.py:154-160),
        drawn from the env's own RandomState in the reference's order; objective evaluated on the device."""
        v = self._vec
        v.t = 0
        fb, hyper = v.sampler.draw(v.rngs)
        if v.fb is None:
            v.fb = fb
        else:
            v.fb.copy_from(fb)
        v.engine.n.zero_()
        v.engine.n_active_max = 0
        v.best.fill_(-float("inf"))
        rewards = []
        for t in range(v.T):
            if self.do_local_af_opt:
                x = self.rng.rand(1, self.D)
            else:
                xi = v.engine.xi_table
                x = xi[int(self.rng.choice(np.arange(self.cardinality_xi_t)))].cpu().numpy().reshape(1, -1)
            cs = ops.CandidateSet(self.D, tail=torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=v.dev))
            n_init, v.n_init_samples = v.n_init_samples, 0           # random queries replace the initial design too
            out = v._step_device(t, True, False, {"cand": cs}, [0])
            v.n_init_samples = n_init
            v.t = t + 1
            rewards.append(float(out["reward"][0]))
        self._host_views()
        assert self.is_terminal()
        return rewards

    def optimize_AF(self):
        """metabo_gym.py:615-622"""
        if self.do_local_af_opt:
            self.get_af_maxima()
            self.xi_t = np.concatenate((self.af_maxima_t, self.multistart_grid), axis=0)
            assert self.xi_t.shape[0] == self.cardinality_xi_t
        else:
            self.xi_t = self._vec.engine.xi_table.cpu().numpy()

    def get_af_maxima(self):
        """metabo_gym.py:706-721 with the bound AF callback"""
        state_at_multistarts = self.get_state(self.multistart_grid)
        af_at_multistarts = np.asarray(self.af(state_at_multistarts))
        self.af_opt_startpoints_t = self.multistart_grid[np.argsort(-af_at_multistarts)[:self.k, ...]]
        local_grids = []
        for x in self.af_opt_startpoints_t:
            lo = np.max((x - 0.5 * self.af_max_search_diam, self.domain[:, 0]), axis=0)     # util.py:45-51
            hi = np.min((x + 0.5 * self.af_max_search_diam, self.domain[:, 1]), axis=0)
            local_grids.append(self.local_search_grid * (hi - lo) + lo)                      # util.py:35-37
        local_grids = np.concatenate(local_grids, axis=0)
        af_on_local_grid = np.asarray(self.af(self.get_state(local_grids)))
        self.af_maxima_t = local_grids[np.argsort(-af_on_local_grid)[:self.cardinality_xi_local_t]]
        assert self.af_maxima_t.shape[0] == self.cardinality_xi_local_t

    def get_state(self, X):
        """metabo_gym.py:624-677: state [M, F] float32 (fixed column order) at arbitrary points X [M, D]; the GP
        posterior comes from the CUDA kernels (fp64, cast to float32 like the reference's state array)."""
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, self.D)
        v = self._vec
        eng = v.engine
        mean, std = self.eval_gp(X)
        inc = self.get_incumbent()
        eng.set_episode_scalars(np.array([inc], dtype=np.float32), np.array([float(v.t)], dtype=np.float32),
                                np.array([float(v.T)], dtype=np.float32))
        cs = ops.CandidateSet(self.D, tail=torch.as_tensor(X, device=v.dev))
        st = eng.state_tensor(cs, torch.as_tensor(mean[None].astype(np.float32), device=v.dev),
                              torch.as_tensor(std[None].astype(np.float32), device=v.dev))
        return st[0].cpu().numpy()

    def eval_gp(self, X_star):
        """metabo_gym.py:732-745 on the CUDA kernels -> (mean [M], std [M]) float64"""
        assert X_star.shape[1] == self.D
        v = self._vec
        cs = ops.CandidateSet(self.D, tail=torch.as_tensor(np.ascontiguousarray(X_star, dtype=np.float64), device=v.dev))
        if getattr(self, "_fit_t", None) != v.t:          # refit once per step (update_gp, metabo_gym.py:610-613)
            v.engine.fit()
            self._fit_t = v.t
        mean, std = ops.gp_posterior(v.engine.gp_state, v.T, cs, 1, kernel=v.kernel, precision=L.GP_FP64,
                                     n_active_max=v.t)
        return mean[0].cpu().numpy(), std[0].cpu().numpy()

    def get_incumbent(self):
        """metabo_gym.py:723-730"""
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
