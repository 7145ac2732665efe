"""VecMetaBO: B MetaBO episodes in lock-step, resident on one GPU.

This is the data-parallel axis of the reference (`n_workers` forked EnvRunner processes, one BO
episode each, metabo/ppo/batchrecorder.py:199-225) turned into a batch dimension: lane b owns the
RNG stream `RandomState(lane_seeds[b])` exactly like `MetaBO.seed` (metabo_gym.py:162-166) and its
episodes are the ones the reference worker with that seed would play.

Per episode step (MetaBO.step, metabo_gym.py:195-217) for all lanes at once:
    x = xi_t[action] -> y = f(x) (device) -> append (add_data :595-608) -> reward (:679-704)
    -> GP refit + AF-optimisation round (AFEngine.af_round) -> next state / logits.
The AF round whose state is never acted on (after the T-th step, metabo_gym.py:211-214) is skipped.
Only fixed horizons (T_min == T_max) are supported in lock-step mode.
"""
import os

import numpy as np
import torch

from .. import _lib as L
from ..engine import AFEngine
from ..sobol import i4_sobol_generate
from .functions import FunctionSampler


class VecMetaBO:
    def __init__(self, B, lane_seeds, gp_precision="fp64", mlp_mode="bf16x3", device=None, lane_rngs=None,
                 **kwargs):
        self.kwargs = kwargs
        self.B = B
        self.D = kwargs["D"]
        if "T" in kwargs:
            self.T = kwargs["T"]
        else:
            if kwargs["T_min"] != kwargs["T_max"]:
                raise NotImplementedError("VecMetaBO runs lanes in lock-step: T_min must equal T_max")
            self.T = kwargs["T_min"]
        self.n_init_samples = int(kwargs.get("n_init_samples", 0))       # Sobol initial design (metabo_gym.py:59-61, 202-204)
        assert self.n_init_samples <= self.T
        self.features = kwargs["features"]
        self.reward_transformation = kwargs["reward_transformation"]
        if self.reward_transformation not in ("none", "neg_linear", "neg_log10"):
            raise ValueError("Unknown reward transformation!")
        self.f_type, self.f_opts = kwargs["f_type"], kwargs["f_opts"]
        self.kernel = kwargs.get("kernel", "RBF")
        if self.kernel not in L.KERNEL_IDS:
            raise ValueError("Unknown kernel function for GP model!")
        self.use_prior_mean_function = bool(kwargs.get("use_prior_mean_function", False))
        self.do_local_af_opt = kwargs["local_af_opt"]
        if lane_rngs is not None:
            # several lanes may share one RandomState (the sequential episodes of one reference worker)
            assert len(lane_rngs) == B
            self.rngs = list(lane_rngs)
        else:
            assert len(lane_seeds) == B
            self.rngs = []
            for s in lane_seeds:
                r = np.random.RandomState()
                r.seed(int(s))
                self.rngs.append(r)
        gp_prec = {"fp64": L.GP_FP64, "fp32": L.GP_FP32}[gp_precision]
        mode = L.MLP_MODES[mlp_mode]
        common = dict(kernel=self.kernel, use_prior_mean=self.use_prior_mean_function, gp_precision=gp_prec,
                      mlp_mode=mode, device=device, T_training=kwargs.get("T_training"))
        if self.do_local_af_opt:
            self.engine = AFEngine(B, self.D, self.T, self.features, N_MS=kwargs["N_MS"], N_LS=kwargs["N_LS"],
                                   k=kwargs["k"], local_search_grid=i4_sobol_generate(self.D, kwargs["N_LS"]),
                                   **common)
            grid = None
        else:
            grid_np = i4_sobol_generate(self.D, kwargs["cardinality_domain"])
            self.engine = AFEngine(B, self.D, self.T, self.features, discrete_domain=grid_np, **common)
            grid = self.engine.xi_table
        self.dev = self.engine.dev
        self.initial_design = (torch.as_tensor(i4_sobol_generate(self.D, self.n_init_samples), device=self.dev)
                               if self.n_init_samples > 0 else None)
        self.engine.reuse_multistart = True
        self.cardinality_xi_t = self.engine.M_final
        self.n_features = self.engine.n_features
        self.sampler = FunctionSampler(self.f_type, self.f_opts, self.D, self.dev, discrete_grid=grid)
        self.fixed_hyper = (kwargs.get("kernel_lengthscale"), kwargs.get("kernel_variance"), kwargs.get("noise_variance"))
        self.policy = None
        self.t = 0
        self.fb = None
        self.cur = None          # output of the AF round that produced the current state
        self.sample_seed = 0
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)   # sampler step counter (device resident)
        self.want_states = True   # materialise mean/std of the final phase (needed by current_state())
        self.best = torch.full((self.B,), -float("inf"), dtype=torch.float64, device=self.dev)
        self.reward = torch.zeros((self.B,), dtype=torch.float64, device=self.dev)
        # CUDA graphs: the device work of "round at t=0" and of "step at time t" is captured once per t and replayed
        # (the loop is launch-bound for reference-sized batches); opt-in, the BatchRecorder turns it on
        self.use_graphs = False
        self._graphs = {}

    # -- policy ----------------------------------------------------------------------------------
    def set_policy(self, pi):
        """Bind a NeuralAF: its weights feed the engine's kernels (replaces set_af_functions +
        the per-step numpy callback of metabo_gym.py:168-174, 708, 718)."""
        self.policy = pi
        self._graphs = {}
        if hasattr(pi, "engine_spec"):
            # closed-form baseline AF (EI / UCB / PI, policies.py:181-307): no network, no value function
            self.engine.closed_form = pi.engine_spec()
            return
        self.engine.closed_form = None
        self.engine.codes_policy = [c for c, m in zip(self.engine.codes_all, pi.policy_mask()) if m]
        self.sync_policy()
        self.engine.policy.workspace(self.dev, 256)

    def sync_policy(self):
        pi = self.policy
        if hasattr(pi, "engine_spec"):
            return
        dev = self.dev
        self.engine.set_policy([l.weight.detach().to(dev) for l in pi.policy_net.fc],
                               [l.bias.detach().to(dev) for l in pi.policy_net.fc], pi.activation)
        if pi.use_value_network:
            self.engine.set_value_net([l.weight.detach().to(dev) for l in pi.value_net.fc],
                                      [l.bias.detach().to(dev) for l in pi.value_net.fc], pi.activation)
            self.engine.value.workspace(dev, 256)          # exists before any graph capture
        if not pi.tensor_core_ok():
            self.engine.mlp_mode = L.MLP_FP32

    # -- episode control -------------------------------------------------------------------------
    def _unit(self, key, fn):
        """Run the device work `fn` eagerly, or (use_graphs) capture it the second time `key` is seen and replay
        it afterwards.  `fn` must be pure stream work on persistent / graph-owned tensors."""
        if not self.use_graphs:
            return fn()
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs[key] = {}
            return fn()                                   # warm-up run (eager)
        if "graph" not in ent:
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = fn()
            ent["graph"], ent["out"] = g, out
        ent["graph"].replay()
        return ent["out"]

    def reset(self, deterministic=True):
        """New objective per lane, empty GP, AF round on the empty GP (MetaBO.reset, :176-193)."""
        eng = self.engine
        self.t = 0
        fb, hyper = self.sampler.draw(self.rngs)
        if self.fb is None:
            self.fb = fb
        else:
            self.fb.copy_from(fb)                         # persistent parameter tensors (graph replays read them)
        if hyper is not None:
            eng.set_hyperparameters(np.repeat(hyper[0][:, None], self.D, axis=1), hyper[1], hyper[2])
        else:
            eng.set_hyperparameters(self.fixed_hyper[0], self.fixed_hyper[1], self.fixed_hyper[2])
        self.cur = self._unit(("round0", deterministic, self.want_states), lambda: self._reset_device(deterministic))

    def _reset_device(self, deterministic):
        eng = self.engine
        eng.n.zero_()
        eng.n_active_max = 0
        self.best.fill_(-float("inf"))
        return self._round(0, deterministic)

    def _round(self, t, deterministic, scalars_set=False):
        eng = self.engine
        if not scalars_set:              # (the fused environment step has already written them)
            inc = torch.where(torch.isinf(self.best), self.fb.y_min, self.best)      # metabo_gym.py:723-730
            eng.set_episode_scalars(inc.float(), torch.full((self.B,), float(t), device=self.dev),
                                    torch.full((self.B,), float(self.T), device=self.dev))
        cur = eng.af_round(deterministic=deterministic, seed=self.sample_seed, step=0, step_dev=self.step_dev,
                           want_stats=False, want_posterior=self.want_states)
        self.step_dev += 1
        cur["value"] = self._values()
        if self.want_states:
            cur["state"] = eng.state_tensor(cur["cand"], cur["mean"], cur["std"])
        return cur

    def _values(self):
        """value net on (t, T) (policies.py:117-121) for the current state of every lane."""
        if self.engine.value is None or self.engine.closed_form is not None:
            return torch.zeros(self.B, device=self.dev)
        from .. import ops
        cols = []
        for idx in (self.policy.t_idx, self.policy.T_idx):       # state columns the value net reads (row 0)
            code = self.engine.codes_all[idx]
            if not (L.FEAT_INCUMBENT <= code <= L.FEAT_BUDGET):
                raise NotImplementedError("value net inputs must be per-episode scalar features")
            cols.append(code - L.FEAT_INCUMBENT)
        epi = self.engine.epi
        tT = torch.stack([epi[:, cols[0]], epi[:, cols[1]]], dim=1).contiguous()
        w = self.engine.value.widths
        if (self.policy.activation == "relu" and len(w) == 6 and len(set(w[1:5])) == 1 and 8 <= w[1] <= 208 and w[5] == 1
                and self.B >= int(os.environ.get("MBO_VALUE_TC_MIN_B", "1"))):
            # the B lanes as the candidate rows of ONE dense "episode" on the tensor-core kernel (bf16x3: fp32-class):
            # 8 tiles at B = 1024 instead of 130 us of the CUDA-core kernel per step (6 % of a rollout); also for few lanes
            # (40 episodes: one tile, ~20 us instead of 42 us of mlp_rows_kernel)
            return ops.af_logits_dense(self.engine.value, [0, 1], tT.view(1, self.B, 2), mode=L.MLP_BF16X3)[0]
        return ops.value_net(self.engine.value, tT)

    def values(self):
        return self.cur["value"]

    def current_state(self):
        """state [B, M, F] fp32 as get_state packs it"""
        if "state" in self.cur:
            return self.cur["state"]
        return self.engine.state_tensor(self.cur["cand"], self.cur["mean"], self.cur["std"])

    def step(self, deterministic=True, advance=True, actions=None):
        """Apply the action chosen by the last AF round (or `actions` [B] int32, indices into xi_t).
        Returns (reward [B] f64, done bool)."""
        t = self.t
        done = (t + 1) == self.T
        if actions is not None:
            out = self._step_device(t, deterministic, advance and not done, self.cur, actions)
        else:
            prev = self.cur
            out = self._unit(("step", t, deterministic, advance and not done, self.want_states),
                             lambda: self._step_device(t, deterministic, advance and not done, prev, None))
        self.t = t + 1
        self.engine.n_active_max = self.t
        if out.get("cur") is not None:
            self.cur = out["cur"]
        return out["reward"], done

    def _step_device(self, t, deterministic, do_round, prev, actions):
        eng = self.engine
        if t < self.n_init_samples:
            x = self.initial_design[t][None, :].expand(self.B, self.D)   # the action is ignored (metabo_gym.py:202-204)
        elif actions is None:
            x = prev["x_action"]                                  # [B, D] fp64
        else:
            from .. import ops
            a = torch.as_tensor(actions, dtype=torch.int32, device=self.dev).reshape(self.B, 1).contiguous()
            x = ops.candidates_gather(prev["cand"], a, self.B)[:, 0]
        fused = (getattr(self.fb, "family", None) is not None and self.reward_transformation in L.REWARDS
                 and os.environ.get("MBO_ENV_TORCH") is None)
        if fused:
            # ONE kernel: evaluate, append, incumbent, reward and the next state's per-episode scalars (mbo_env_step)
            from .. import ops
            reward = ops.env_step(self.fb.family, t, x.contiguous(), self.fb.t, self.fb.s, self.fb.y_max, self.fb.y_min, eng.X, eng.Y,
                                  eng.n, self.best, L.REWARDS[self.reward_transformation],
                                  torch.empty(self.B, dtype=torch.float64, device=self.dev),
                                  epi=eng.epi if do_round else None, t_next=t + 1, T_budget=self.T, T_training=eng.T_training)
            eng.n_active_max = t + 1
            out = {"reward": reward, "cur": None}
            if do_round:
                out["cur"] = self._round(t + 1, deterministic, scalars_set=True)
            return out
        y = self.fb.eval(x)
        eng.X[:, t] = x
        eng.Y[:, t] = y
        eng.n.fill_(t + 1)
        eng.n_active_max = t + 1
        torch.maximum(self.best, y, out=self.best)
        regret = self.fb.y_max - self.best                        # min(y_max - Y)
        if self.reward_transformation == "none":
            reward = regret
        elif self.reward_transformation == "neg_linear":
            reward = -regret
        else:
            reward = -torch.log10(torch.clamp(regret, min=1e-20))
        out = {"reward": reward, "cur": None}
        if do_round:
            out["cur"] = self._round(t + 1, deterministic)
        return out
