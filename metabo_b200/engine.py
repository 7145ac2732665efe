"""Batched AF-evaluation engine: B independent BO episodes in lock-step on one GPU.

One `af_round()` is what the reference executes per environment step and per episode in
`MetaBO.step` / `reset` (metabo/environment/metabo_gym.py:176-217):

    update_gp (:610)  ->  optimize_AF (:615): get_af_maxima (:706-721)
        phase 1  state on the multistart grid  -> AF -> top-k starts
        phase 2  state on k local Sobol grids   -> AF -> top-k maxima
    get_state(xi_t) (:214, xi_t = [af_maxima; multistart grid] :619)
        phase 3  -> NeuralAF.act (policies.py:135-148): argmax / categorical sample

All candidate sets are generated in-kernel from the shared meshgrid / Sobol tables (`CandidateSet`),
GP state lives in one fp64 blob per episode, and every step is enqueued on the current CUDA stream
without host synchronisation.  Discrete domains (`local_af_opt=False`, e.g. train_metabo_gps.py)
skip phases 1-2.
"""
import os

import numpy as np
import torch

from . import _lib as L
from . import ops

FEATURE_ORDER = ["posterior_mean", "posterior_std", "x", "incumbent", "timestep_perc", "timestep", "budget"]


def feature_codes(features, D):
    """MBO_FEAT_* code of every state column in the reference's fixed order (metabo_gym.py:630-671)."""
    codes = []
    for f in FEATURE_ORDER:
        if f not in features:
            continue
        if f == "x":
            codes += [L.FEAT_X0 + d for d in range(D)]
        else:
            codes.append({"posterior_mean": L.FEAT_MEAN, "posterior_std": L.FEAT_STD,
                          "incumbent": L.FEAT_INCUMBENT, "timestep_perc": L.FEAT_TPERC,
                          "timestep": L.FEAT_TIMESTEP, "budget": L.FEAT_BUDGET}[f])
    unknown = [f for f in features if f not in FEATURE_ORDER]
    if unknown:
        raise ValueError("Invalid feature specification!")       # metabo_gym.py:675
    return codes


def uniform_grid(D, n_per_dim):
    """util.py:24-32 create_uniform_grid on the unit cube (meshgrid order of numpy's default 'xy')."""
    axes = [np.linspace(0.0, 1.0, int(n_per_dim)) for _ in range(D)]
    mesh = np.meshgrid(*axes)
    return np.vstack(mesh).reshape((D, -1)).T.copy()


class AFEngine:
    def __init__(self, B, D, T, features, policy_mask=None, kernel="RBF", use_prior_mean=False,
                 N_MS=None, N_LS=None, k=None, local_search_grid=None, discrete_domain=None,
                 gp_precision=L.GP_FP64, mlp_mode=L.MLP_FP32, device=None, T_training=None):
        if not torch.cuda.is_available():
            raise L.MboError("metabo_b200 needs a CUDA device (no CPU fallback)")
        L.lib()
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.B, self.D, self.T = B, D, T
        self.features = list(features)
        self.kernel, self.use_prior_mean = kernel, use_prior_mean
        self.gp_precision, self.mlp_mode = gp_precision, mlp_mode
        self.T_training = T_training
        codes = feature_codes(features, D)
        self.n_features = len(codes)
        mask = [True] * len(codes) if policy_mask is None else list(policy_mask)
        self.codes_policy = [c for c, m in zip(codes, mask) if m]
        self.codes_all = codes
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.X = torch.zeros((B, T, D), **f64)
        self.Y = torch.zeros((B, T), **f64)
        self.n = torch.zeros((B,), dtype=torch.int32, device=self.dev)
        self.lengthscale = torch.ones((B, D), **f64)
        self.variance = torch.ones((B,), **f64)
        self.noise = torch.zeros((B,), **f64)
        self.gp_state = torch.zeros((B, ops.gp_state_doubles(T, D)), **f64)
        self.status = torch.zeros((B,), dtype=torch.int32, device=self.dev)
        self.epi = torch.zeros((B, 4), dtype=torch.float32, device=self.dev)
        self.discrete = discrete_domain is not None
        if self.discrete:
            self.xi_table = torch.as_tensor(np.asarray(discrete_domain, dtype=np.float64), device=self.dev).contiguous()
            self.M_final = self.xi_table.shape[0]
        else:
            per_dim = int(np.floor(N_MS ** (1 / D)))
            self.N_MS_per_dim = per_dim
            self.multistart = torch.as_tensor(uniform_grid(D, per_dim), device=self.dev).contiguous()
            self.N_MS = self.multistart.shape[0]
            self.k, self.N_LS = k, N_LS
            self.local_table = torch.as_tensor(np.asarray(local_search_grid, dtype=np.float64),
                                               device=self.dev).contiguous()
            assert self.local_table.shape == (N_LS, D)
            self.diam = 2 * 1 / per_dim                      # metabo_gym.py:86
            self.M_final = self.k + self.N_MS
            self.cs_ms = ops.CandidateSet(D, tail=self.multistart)
        self.policy = ops.PolicyHandle()
        self.value = None
        self.n_active_max = T
        self.launches = 0
        self.reuse_multistart = False   # phase 3 reuses the phase-1 results of the multistart grid (VecMetaBO turns it on)
        self.evals_skipped = 0
        self.episode_offset = 0         # global index of lane 0 (keys the sampler under sharding)
        # top-k / argmax selection in the tensor-core kernel's epilogue (mbo_af_topk): the logits of the search phases never
        # reach HBM.  Default since r2 (fp16: step 4.28 -> 4.23 ms, tf32 / bf16x3 +1.5 %; MBO_FUSE_TOPK=0: logits -> topk_kernel)
        self.fuse_topk = os.environ.get("MBO_FUSE_TOPK", "1") != "0"
        self.keep_phase_logits = False  # also materialise the phase-1/2 logits (parity tests, smoke)
        self.fuse_gp = False         # opt-in: mbo_af_eval_fused (correct, but slower than the two-kernel path so far, see profiles/r01_notes.md)
        self.tc_ok = False           # set by set_policy: architecture supported by the tensor-core kernel
        self.closed_form = None      # (kind, p0, p1): EI / UCB / PI of the reference's baselines instead of the NeuralAF
        self.kernel_events = None   # list -> (tag, start, end) CUDA events around GP / MLP launches

    # -- parameters -------------------------------------------------------------------------------
    def set_policy(self, weights, biases, activation="relu"):
        self.policy.set_weights(weights, biases, activation)
        w = [int(x.shape[0]) for x in weights]
        self.tc_ok = (activation == "relu" and len(weights) == 5 and len(set(w[:4])) == 1 and 8 <= w[0] <= 208
                      and w[4] == 1 and int(weights[0].shape[1]) <= 15)
        if self.mlp_mode == L.MLP_FP16 and w[0] > 206:
            self.mlp_mode = L.MLP_TF32    # fp16 mode needs two spare K columns; tf32 has the same significand

    def set_value_net(self, weights, biases, activation="relu"):
        # ONE handle for the life of the engine: captured CUDA graphs (VecMetaBO) hold its device pointers by value,
        # and mbo_policy_set_weights keeps the buffers while the architecture is unchanged
        if self.value is None:
            self.value = ops.PolicyHandle()
        self.value.set_weights(weights, biases, activation)

    def set_hyperparameters(self, lengthscale, variance, noise):
        """per-episode or broadcast GP hyper-parameters (metabo_gym.py:147-149, :250-252)."""
        ls = torch.as_tensor(np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (self.B, self.D)).copy())
        self.lengthscale.copy_(ls)
        self.variance.copy_(torch.as_tensor(np.broadcast_to(np.asarray(variance, dtype=np.float64), (self.B,)).copy()))
        self.noise.copy_(torch.as_tensor(np.broadcast_to(np.asarray(noise, dtype=np.float64), (self.B,)).copy()))

    def set_data(self, X, Y, n, n_active_max=None):
        """X [B,T,D], Y [B,T], n [B] device tensors (padding beyond n[b] ignored)."""
        self.X.copy_(X, non_blocking=True)
        self.Y.copy_(Y, non_blocking=True)
        self.n.copy_(n, non_blocking=True)
        self.n_active_max = self.T if n_active_max is None else int(n_active_max)

    def set_episode_scalars(self, incumbent, t, T_budget):
        """epi[b] = {incumbent, t/T, timestep, budget} of get_state (metabo_gym.py:642-671)."""
        inc = torch.as_tensor(incumbent, dtype=torch.float32, device=self.dev)
        t = torch.as_tensor(t, dtype=torch.float32, device=self.dev)
        Tb = torch.as_tensor(T_budget, dtype=torch.float32, device=self.dev)
        self.epi[:, 0] = inc
        self.epi[:, 1] = t / Tb
        if self.T_training is not None:
            self.epi[:, 2] = torch.clamp(t, max=float(self.T_training))
            self.epi[:, 3] = float(self.T_training)
        else:
            self.epi[:, 2] = t
            self.epi[:, 3] = Tb

    # -- the hot path -----------------------------------------------------------------------------
    def fit(self):
        ops.gp_fit(self.n, self.X, self.Y, self.lengthscale, self.variance, self.noise, kernel=self.kernel,
                   use_prior_mean=self.use_prior_mean, gp_state=self.gp_state, status=self.status)
        self.launches += 1

    def evaluate_topk(self, cand, k, want_logits=False):
        """GP posterior + NeuralAF + top-k; on the tensor-core path the selection is fused into the kernel's
        epilogue and the logits are only written when asked for.  Returns (mean, std, logits or None, idx)."""
        if self.closed_form is not None or self.mlp_mode == L.MLP_FP32 or not self.tc_ok or not self.fuse_topk or self.can_fuse():
            mean, std, lg = self.evaluate(cand, want_posterior=True)
            idx, _ = ops.logits_topk(lg, k)
            self.launches += 1
            return mean, std, lg, idx
        ev = self.kernel_events
        if ev is not None:
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
        mean, std = ops.gp_posterior(self.gp_state, self.T, cand, self.B, kernel=self.kernel,
                                     precision=self.gp_precision, out_dtype=torch.float32,
                                     n_active_max=self.n_active_max)
        if ev is not None:
            e1.record()
        lg = torch.empty((self.B, cand.M), dtype=torch.float32, device=self.dev) if want_logits else None
        idx, _ = ops.af_topk(self.policy, self.codes_policy, cand, mean, std, self.epi, self.B, k, mode=self.mlp_mode,
                             logits=lg)
        if ev is not None:
            e2.record()
            ev.append(("gp_M%d" % cand.M, e0, e1))
            ev.append(("mlp_M%d" % cand.M, e1, e2))
        self.launches += 3
        return mean, std, lg, idx

    def evaluate(self, cand, want_posterior=False):
        """GP posterior + NeuralAF logits on one candidate set -> (mean, std, logits)."""
        ev = self.kernel_events
        if ev is not None:
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
        if self.closed_form is not None:
            mean, std = ops.gp_posterior(self.gp_state, self.T, cand, self.B, kernel=self.kernel,
                                         precision=self.gp_precision, out_dtype=torch.float32,
                                         n_active_max=self.n_active_max)
            kind, p0, p1 = self.closed_form
            af = ops.af_closed_form(kind, mean, std, self.epi, p0=p0, p1=p1, D=self.D)
            self.launches += 2
            return mean, std, af
        if self.can_fuse():
            # one kernel: fp64 GP warps feed the tensor-core MLP through shared memory
            mean = std = None
            if want_posterior:
                mean = torch.empty((self.B, cand.M), dtype=torch.float32, device=self.dev)
                std = torch.empty((self.B, cand.M), dtype=torch.float32, device=self.dev)
            logits = ops.af_eval_fused(self.policy, self.codes_policy, cand, self.gp_state, self.T, self.n_active_max,
                                       self.epi, self.B, kernel=self.kernel, mode=self.mlp_mode, mean=mean, std=std)
            if ev is not None:
                e2.record()
                ev.append(("fused_M%d" % cand.M, e0, e2))
            self.launches += 1
            return mean, std, logits
        mean, std = ops.gp_posterior(self.gp_state, self.T, cand, self.B, kernel=self.kernel,
                                     precision=self.gp_precision, out_dtype=torch.float32,
                                     n_active_max=self.n_active_max)
        if ev is not None:
            e1.record()
        logits = ops.af_logits(self.policy, self.codes_policy, cand, mean, std, self.epi, self.B, mode=self.mlp_mode)
        if ev is not None:
            e2.record()
            ev.append(("gp_M%d" % cand.M, e0, e1))
            ev.append(("mlp_M%d" % cand.M, e1, e2))
        self.launches += 2
        return mean, std, logits

    def can_fuse(self):
        return (self.closed_form is None and self.fuse_gp and self.mlp_mode != L.MLP_FP32 and self.gp_precision == L.GP_FP64
                and self.n_active_max <= ops.FUSED_MAX_N and self.tc_ok)

    def af_round(self, deterministic=True, seed=0, step=0, refit=True, want_stats=False, want_posterior=False,
                 step_dev=None, want_logits=True):
        """One AF-optimisation round for all B episodes.  Returns a dict of device tensors:
        action [B] i32, x_action [B,D] f64, xi_t candidate set, logits [B,M] (phase 3), and with
        want_stats also logp / entropy [B].  want_logits=False (greedy evaluation, policies.py:127-136 only needs the
        argmax): phase 3 takes the action from the kernel's selection epilogue (k = 1: first-index ties like
        torch.argmax) and no logit of the round is written to HBM."""
        if refit:
            self.fit()
        out = {}
        if self.discrete:
            cs = ops.CandidateSet(self.D, tail=self.xi_table)
        else:
            mean1, std1, lg1, top1 = self.evaluate_topk(self.cs_ms, self.k, want_logits=self.reuse_multistart or self.keep_phase_logits)
            starts = ops.candidates_gather(self.cs_ms, top1, self.B)
            cubes = ops.cubes_around(starts, self.diam)
            cs_ls = ops.CandidateSet(self.D, tail=self.local_table, cubes=cubes)
            _, _, lg2, top2 = self.evaluate_topk(cs_ls, self.k, want_logits=self.keep_phase_logits)
            maxima = ops.candidates_gather(cs_ls, top2, self.B)
            cs = ops.CandidateSet(self.D, head=maxima, tail=self.multistart)
            self.launches += 3
            out.update(af_ms=lg1, top_ms=top1, af_ls=lg2, top_ls=top2, af_maxima=maxima, cubes=cubes)
        if self.reuse_multistart and not self.discrete:
            # xi_t = [af_maxima; multistart grid]: the grid part was evaluated in phase 1 with the same GP state,
            # so only the k new points are evaluated and the phase-1 results are reused (bitwise identical to
            # re-evaluating them: every candidate row is computed independently of its neighbours)
            cs_head = ops.CandidateSet(self.D, head=maxima)
            mh, sh, lgh = self.evaluate(cs_head, want_posterior=want_posterior)
            lg3 = torch.cat([lgh, lg1], dim=1)
            mean = torch.cat([mh, mean1], dim=1) if (want_posterior and mean1 is not None) else None
            std = torch.cat([sh, std1], dim=1) if (want_posterior and std1 is not None) else None
            self.evals_skipped += self.B * self.N_MS
        elif deterministic and not want_logits and not want_stats:
            mean, std, lg3, top3 = self.evaluate_topk(cs, 1)
            action = top3[:, 0].contiguous() if lg3 is None else None
        else:
            mean, std, lg3 = self.evaluate(cs, want_posterior=want_posterior)
        if deterministic and lg3 is None:
            pass                                     # the action came out of the selection epilogue
        elif deterministic:
            action, _ = ops.logits_argmax(lg3)
        else:
            action = ops.logits_sample(lg3, seed, step, episode_offset=self.episode_offset, step_dev=step_dev)
        x_action = ops.candidates_gather(cs, action[:, None].contiguous(), self.B)[:, 0]
        self.launches += 2
        out.update(cand=cs, mean=mean, std=std, logits=lg3, action=action, x_action=x_action)
        if want_stats:
            out["logp"], out["entropy"] = ops.logits_logp_entropy(lg3, action)
            self.launches += 1
        return out

    def state_tensor(self, cand, mean, std):
        """Materialise state [B,M,F] fp32 exactly as get_state packs it (for the PPO buffer): one kernel."""
        return ops.state_pack(cand, list(self.codes_all), mean.contiguous(), std.contiguous(), self.epi, mean.shape[0])
