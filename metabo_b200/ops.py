"""torch-tensor wrappers of the C ABI.  All tensors must live on the current CUDA device; every
call enqueues on torch's current stream and returns without synchronising."""
import ctypes as C

import torch

from . import _lib as L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _chk(t, dtype, name):
    if t is None:
        return
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError("%s must be a contiguous CUDA %s tensor" % (name, dtype))


class CandidateSet:
    """Device-side description of the M candidates of every episode (see mbo_candidates)."""

    def __init__(self, D, head=None, tail=None, cubes=None):
        self.D = D
        self.head, self.tail, self.cubes = head, tail, cubes
        n_head = 0 if head is None else head.shape[1]
        n_tail = 0 if tail is None else tail.shape[0]
        for t, nm in ((head, "head"), (tail, "tail")):
            if t is not None:
                if not (t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.float64)):
                    raise ValueError("%s must be a contiguous CUDA float32/float64 tensor" % nm)
                if t.shape[-1] != D:
                    raise ValueError("%s last dim must be D" % nm)
        local = cubes is not None
        if local:
            _chk(cubes, torch.float64, "cubes")
            n_cubes = cubes.shape[1]
            M = n_head + n_cubes * n_tail
        else:
            n_cubes = 0
            M = n_head + n_tail
        self.M = M
        self.B_head = None if head is None else head.shape[0]
        self.c = L.Candidates(M, D, n_head, n_tail, int(local), n_cubes,
                              int(head is not None and head.dtype == torch.float32),
                              int(tail is not None and tail.dtype == torch.float32),
                              0 if head is None else head.data_ptr(),
                              0 if tail is None else tail.data_ptr(),
                              0 if cubes is None else cubes.data_ptr())

    def ref(self):
        return C.byref(self.c)


def gp_state_doubles(nmax, D):
    return int(L.lib().mbo_gp_state_doubles(nmax, D))


def gp_fit(n, X, Y, lengthscale, variance, noise, kernel="RBF", use_prior_mean=False, gp_state=None, status=None):
    """MetaBO.update_gp for B episodes.  n [B] i32, X [B,nmax,D] f64, Y [B,nmax] f64,
    lengthscale [B,D], variance [B], noise [B] f64.  Returns (gp_state [B,S] f64, status [B] i32)."""
    B, nmax, D = X.shape
    _chk(n, torch.int32, "n"); _chk(X, torch.float64, "X"); _chk(Y, torch.float64, "Y")
    _chk(lengthscale, torch.float64, "lengthscale"); _chk(variance, torch.float64, "variance")
    _chk(noise, torch.float64, "noise")
    S = gp_state_doubles(nmax, D)
    if gp_state is None:
        gp_state = torch.empty((B, S), dtype=torch.float64, device=X.device)
    if status is None:
        status = torch.empty((B,), dtype=torch.int32, device=X.device)
    L.check(L.lib().mbo_gp_fit(B, nmax, D, L.KERNEL_IDS[kernel], _p(n), _p(X), _p(Y), _p(lengthscale), _p(variance),
                               _p(noise), int(bool(use_prior_mean)), _p(gp_state), _p(status), _stream()),
            "mbo_gp_fit")
    return gp_state, status


def gp_posterior(gp_state, nmax, cand, B, kernel="RBF", precision=L.GP_FP64, out_dtype=torch.float64,
                 n_active_max=None, mean=None, std=None):
    """MetaBO.eval_gp for B episodes x M candidates -> (mean [B,M], std [B,M])."""
    _chk(gp_state, torch.float64, "gp_state")
    if mean is None:
        mean = torch.empty((B, cand.M), dtype=out_dtype, device=gp_state.device)
    if std is None:
        std = torch.empty((B, cand.M), dtype=out_dtype, device=gp_state.device)
    L.check(L.lib().mbo_gp_posterior(B, nmax, nmax if n_active_max is None else int(n_active_max),
                                     L.KERNEL_IDS[kernel], precision, _p(gp_state), cand.ref(), _p(mean), _p(std),
                                     int(mean.dtype == torch.float32), _stream()), "mbo_gp_posterior")
    return mean, std


class PolicyHandle:
    """Owns an `mbo_policy` (device copies of one MLP in the kernels' layouts)."""

    def __init__(self):
        h = C.c_void_p()
        L.check(L.lib().mbo_policy_create(C.byref(h)), "mbo_policy_create")
        self.h = h
        self.widths = None
        self.ws = None       # device workspace of the tensor-core path (word 0 = sticky error flag)

    def set_weights(self, weights, biases, activation="relu"):
        """weights[l]: [out,in] fp32 CUDA tensors (nn.Linear.weight), biases[l]: [out]."""
        n = len(weights)
        ws = [w.detach().contiguous().float() for w in weights]
        bs = [b.detach().contiguous().float() for b in biases]
        on_dev = all(t.is_cuda for t in ws + bs)
        if not on_dev:
            ws = [t.cpu() for t in ws]
            bs = [t.cpu() for t in bs]
        widths = [ws[0].shape[1]] + [w.shape[0] for w in ws]
        W = (C.c_void_p * n)(*[t.data_ptr() for t in ws])
        Bv = (C.c_void_p * n)(*[t.data_ptr() for t in bs])
        wd = (C.c_int32 * (n + 1))(*widths)
        L.check(L.lib().mbo_policy_set_weights(self.h, n, wd, L.ACT[activation], W, Bv, int(on_dev), _stream()),
                "mbo_policy_set_weights")
        if not on_dev:
            torch.cuda.current_stream().synchronize()
        self.widths = widths
        self._keep = (ws, bs)

    def workspace(self, device, nbytes=256):
        """word 0 = sticky error flag of the tensor-core kernels; the rest holds top-k partial lists"""
        n = (int(nbytes) + 3) // 4
        if self.ws is None or self.ws.device != device or self.ws.numel() < n:
            self.ws = torch.zeros(max(n, 64), dtype=torch.int32, device=device)
        return self.ws

    def check(self):
        """Synchronises and raises if a tensor-core kernel reported a pipeline time-out."""
        if self.ws is not None:
            flag = int(self.ws[0].item())
            if flag != 0:
                self.ws.zero_()
                if flag & 1:
                    raise L.MboError("neural_af_tc kernel aborted: an mbarrier wait timed out (pipeline bug)")
                raise L.MboError("MBO_MLP_FP16: a hidden activation reached the fp16 range limit and was clamped; "
                                 "use mlp_mode 'tf32' or 'bf16x3' for this policy")

    def __del__(self):
        try:
            if self.h:
                L.lib().mbo_policy_destroy(self.h)
                self.h = None
        except Exception:
            pass


def af_logits(policy, feat_src, cand, mean, std, epi, B, mode=L.MLP_FP32, logits=None):
    """get_state + NeuralAF.forward (policy net) for B x M candidates -> logits [B,M] fp32."""
    _chk(mean, torch.float32, "mean"); _chk(std, torch.float32, "std"); _chk(epi, torch.float32, "epi")
    if logits is None:
        logits = torch.empty((B, cand.M), dtype=torch.float32, device=mean.device)
    F = len(feat_src)
    fs = (C.c_int32 * F)(*feat_src)
    ws = policy.workspace(mean.device) if mode != L.MLP_FP32 else None
    L.check(L.lib().mbo_af_logits(policy.h, mode, B, F, fs, cand.ref(), _p(mean), _p(std), _p(epi), _p(logits),
                                  _p(ws), 0 if ws is None else ws.numel() * 4, _stream()), "mbo_af_logits")
    return logits


FUSED_MAX_N = 32


def af_topk(policy, feat_src, cand, mean, std, epi, B, k, mode=L.MLP_TF32, logits=None, idx=None, val=None):
    """NeuralAF logits + top-k fused in the tensor-core kernel's epilogue -> (idx [B,k] i32, val [B,k] f32).
    `logits` (optional [B,M] buffer) is only written when given."""
    _chk(mean, torch.float32, "mean"); _chk(std, torch.float32, "std"); _chk(epi, torch.float32, "epi")
    _chk(logits, torch.float32, "logits")
    if idx is None:
        idx = torch.empty((B, k), dtype=torch.int32, device=mean.device)
    if val is None:
        val = torch.empty((B, k), dtype=torch.float32, device=mean.device)
    F = len(feat_src)
    fs = (C.c_int32 * F)(*feat_src)
    nbytes = int(L.lib().mbo_af_workspace_bytes(mode, B, cand.M))
    ws = policy.workspace(mean.device, nbytes)
    L.check(L.lib().mbo_af_topk(policy.h, mode, B, F, fs, cand.ref(), _p(mean), _p(std), _p(epi), int(k), _p(idx), _p(val),
                                _p(logits), _p(ws), ws.numel() * 4, _stream()), "mbo_af_topk")
    return idx, val


def af_eval_fused(policy, feat_src, cand, gp_state, nmax, n_active_max, epi, B, kernel="RBF", mode=L.MLP_TF32,
                  logits=None, mean=None, std=None):
    """GP posterior (fp64, fused warps) + features + NeuralAF logits in ONE kernel (n <= 32).
    mean/std [B,M] fp32 are only written when given (PPO state buffer)."""
    _chk(gp_state, torch.float64, "gp_state"); _chk(epi, torch.float32, "epi")
    _chk(mean, torch.float32, "mean"); _chk(std, torch.float32, "std")
    if logits is None:
        logits = torch.empty((B, cand.M), dtype=torch.float32, device=gp_state.device)
    F = len(feat_src)
    fs = (C.c_int32 * F)(*feat_src)
    ws = policy.workspace(gp_state.device)
    L.check(L.lib().mbo_af_eval_fused(policy.h, mode, B, nmax, int(n_active_max), L.KERNEL_IDS[kernel], _p(gp_state),
                                      F, fs, cand.ref(), _p(epi), _p(logits), _p(mean), _p(std), _p(ws),
                                      ws.numel() * 4, _stream()), "mbo_af_eval_fused")
    return logits


def af_logits_dense(policy, cols, states, mode=L.MLP_FP32, logits=None):
    """NeuralAF.forward (policy net) on materialised states [B,M,F_state] fp32 -> logits [B,M]."""
    _chk(states, torch.float32, "states")
    B, M, Fs = states.shape
    if logits is None:
        logits = torch.empty((B, M), dtype=torch.float32, device=states.device)
    F = len(cols)
    cc = (C.c_int32 * F)(*cols)
    ws = policy.workspace(states.device) if mode != L.MLP_FP32 else None
    L.check(L.lib().mbo_af_logits_dense(policy.h, mode, B, M, Fs, F, cc, _p(states), _p(logits), _p(ws),
                                        0 if ws is None else ws.numel() * 4, _stream()), "mbo_af_logits_dense")
    return logits


def value_net(policy, tT):
    _chk(tT, torch.float32, "tT")
    out = torch.empty((tT.shape[0],), dtype=torch.float32, device=tT.device)
    L.check(L.lib().mbo_value_net(policy.h, tT.shape[0], _p(tT), _p(out), _stream()), "mbo_value_net")
    return out


def af_closed_form(kind, mean, std, epi, p0=0.0, p1=0.0, D=0, out=None):
    """EI / UCB / PI of the reference's baselines (policies.py:181-307) on a GP posterior -> af [B,M] fp32.
    kind: L.AF_EI, L.AF_UCB (p0 = kappa, p1 = delta > 0 for the GP-UCB schedule), L.AF_PI (p0 = xi)."""
    _chk(mean, torch.float32, "mean"); _chk(std, torch.float32, "std"); _chk(epi, torch.float32, "epi")
    B, M = mean.shape
    if out is None:
        out = torch.empty((B, M), dtype=torch.float32, device=mean.device)
    L.check(L.lib().mbo_af_closed_form(int(kind), B, M, _p(mean), _p(std), _p(epi), float(p0), float(p1), int(D),
                                       _p(out), _stream()), "mbo_af_closed_form")
    return out


def logits_argmax(logits, a=None, v=None):
    _chk(logits, torch.float32, "logits")
    B, M = logits.shape
    if a is None:
        a = torch.empty((B,), dtype=torch.int32, device=logits.device)
    if v is None:
        v = torch.empty((B,), dtype=torch.float32, device=logits.device)
    L.check(L.lib().mbo_logits_argmax(B, M, _p(logits), _p(a), _p(v), _stream()), "mbo_logits_argmax")
    return a, v


def logits_topk(logits, k, idx=None, val=None):
    _chk(logits, torch.float32, "logits")
    B, M = logits.shape
    if idx is None:
        idx = torch.empty((B, k), dtype=torch.int32, device=logits.device)
    if val is None:
        val = torch.empty((B, k), dtype=torch.float32, device=logits.device)
    L.check(L.lib().mbo_logits_topk(B, M, k, _p(logits), _p(idx), _p(val), _stream()), "mbo_logits_topk")
    return idx, val


def logits_sample(logits, seed, step, a=None, episode_offset=0, step_dev=None):
    """Gumbel-max sample; `episode_offset` keys the noise by the GLOBAL episode index when `logits` is a slice of a
    larger batch; `step_dev` (int64 device scalar) is added to `step` on the device so that a captured CUDA graph
    draws fresh noise on every replay."""
    _chk(logits, torch.float32, "logits")
    B, M = logits.shape
    if a is None:
        a = torch.empty((B,), dtype=torch.int32, device=logits.device)
    L.check(L.lib().mbo_logits_sample(B, M, _p(logits), int(seed), int(step), _p(step_dev), int(episode_offset), _p(a),
                                      _stream()), "mbo_logits_sample")
    return a


def logits_logp_entropy(logits, actions):
    _chk(logits, torch.float32, "logits"); _chk(actions, torch.int32, "actions")
    B, M = logits.shape
    lp = torch.empty((B,), dtype=torch.float32, device=logits.device)
    en = torch.empty((B,), dtype=torch.float32, device=logits.device)
    L.check(L.lib().mbo_logits_logp_entropy(B, M, _p(logits), _p(actions), _p(lp), _p(en), _stream()),
            "mbo_logits_logp_entropy")
    return lp, en


def candidates_gather(cand, idx, B, out=None):
    _chk(idx, torch.int32, "idx")
    k = idx.shape[1]
    if out is None:
        out = torch.empty((B, k, cand.D), dtype=torch.float64, device=idx.device)
    L.check(L.lib().mbo_candidates_gather(B, cand.ref(), k, _p(idx), _p(out), _stream()), "mbo_candidates_gather")
    return out


def cubes_around(x, diam, cubes=None):
    _chk(x, torch.float64, "x")
    B, k, D = x.shape
    if cubes is None:
        cubes = torch.empty((B, k, 2, D), dtype=torch.float64, device=x.device)
    L.check(L.lib().mbo_cubes_around(B, k, D, _p(x), float(diam), _p(cubes), _stream()), "mbo_cubes_around")
    return cubes


def state_pack(cand, feat_src, mean, std, epi, B, out=None):
    """state [B, M, F] fp32 as get_state packs it (`mbo_state_pack`, one launch)."""
    _chk(mean, torch.float32, "mean"); _chk(std, torch.float32, "std"); _chk(epi, torch.float32, "epi")
    F = len(feat_src)
    if out is None:
        out = torch.empty((B, cand.M, F), dtype=torch.float32, device=mean.device)
    fs = (C.c_int32 * F)(*feat_src)
    L.check(L.lib().mbo_state_pack(B, cand.ref(), F, fs, _p(mean), _p(std), _p(epi), _p(out), _stream()), "mbo_state_pack")
    return out


def env_step(family, t_idx, x, trans, scale, y_max, y_min, X, Y, n, best, reward_kind, reward, epi=None, t_next=0.0,
             T_budget=1.0, T_training=None):
    """One BO step of all lanes on an analytic objective family (`mbo_env_step`): evaluate, append, incumbent, reward and
    the per-episode scalars of the next state in ONE launch."""
    for t_, nm in ((x, "x"), (trans, "trans"), (scale, "scale"), (y_max, "y_max"), (y_min, "y_min"), (X, "X"), (Y, "Y"),
                   (best, "best"), (reward, "reward")):
        _chk(t_, torch.float64, nm)
    _chk(n, torch.int32, "n"); _chk(epi, torch.float32, "epi")
    B, T, D = X.shape
    L.check(L.lib().mbo_env_step(int(family), B, T, D, int(t_idx), _p(x), _p(trans), _p(scale), _p(y_max), _p(y_min), _p(X), _p(Y),
                                 _p(n), _p(best), int(reward_kind), _p(reward), _p(epi), float(t_next), float(T_budget),
                                 -1.0 if T_training is None else float(T_training), _stream()), "mbo_env_step")
    return reward


class PPOPolicyGrad:
    """Workspace + call wrapper of `mbo_ppo_policy_grad` (the policy-net forward + backward of one PPO optimiser
    step, ppo.py:129-178) for a fixed minibatch shape [E, M, F_state]."""
    HP = 208

    def __init__(self, E, M, F_state, cols, H, device):
        self.E, self.M, self.F_state, self.cols, self.H = int(E), int(M), int(F_state), list(cols), int(H)
        self.nbytes = int(L.lib().mbo_ppo_workspace_bytes(self.E, self.M))
        self.ws = torch.zeros(self.nbytes // 4 + 64, dtype=torch.int32, device=device)      # zeroed once: sticky error word
        self.terms = torch.empty((self.E, 4), dtype=torch.float32, device=device)
        self._cols = (C.c_int32 * len(self.cols))(*self.cols)

    def __call__(self, states, weights, biases, actions, advs, logp_old, epsilon, ent_coeff, n_global, dW, db):
        """weights / biases / dW / db: lists of 5 contiguous fp32 CUDA tensors (torch layouts); the gradients of
        loss_ppo + ent_coeff * loss_ent are written into dW / db.  Returns terms [E, 4]."""
        _chk(states, torch.float32, "states"); _chk(actions, torch.int32, "actions")
        _chk(advs, torch.float32, "advs"); _chk(logp_old, torch.float32, "logp_old")
        E = states.shape[0]
        assert E <= self.E and tuple(states.shape[1:]) == (self.M, self.F_state), states.shape
        assert len(weights) == len(biases) == len(dW) == len(db) == 5
        for t in list(weights) + list(biases) + list(dW) + list(db):
            _chk(t, torch.float32, "parameter / gradient")
        W = (C.c_void_p * 5)(*[t.data_ptr() for t in weights])
        Bv = (C.c_void_p * 5)(*[t.data_ptr() for t in biases])
        gW = (C.c_void_p * 5)(*[t.data_ptr() for t in dW])
        gB = (C.c_void_p * 5)(*[t.data_ptr() for t in db])
        L.check(L.lib().mbo_ppo_policy_grad(E, self.M, self.F_state, len(self.cols), self._cols, self.H, _p(states),
                                            W, Bv, _p(actions), _p(advs), _p(logp_old), float(epsilon), float(ent_coeff),
                                            1.0 / float(n_global), gW, gB, _p(self.terms), _p(self.ws), self.nbytes,
                                            _stream()), "mbo_ppo_policy_grad")
        return self.terms[:E]

    def logp_entropy(self, states, weights, biases, actions):
        """Forward only: (log-prob of the action, entropy) [E] under the given weights, in the arithmetic of the
        update kernels (3xTF32 forward) -- ppo.py:146-147's logprobs_old of the frozen old policy."""
        _chk(states, torch.float32, "states"); _chk(actions, torch.int32, "actions")
        E = states.shape[0]
        assert E <= self.E and tuple(states.shape[1:]) == (self.M, self.F_state), states.shape
        W = (C.c_void_p * 5)(*[t.data_ptr() for t in weights])
        Bv = (C.c_void_p * 5)(*[t.data_ptr() for t in biases])
        L.check(L.lib().mbo_ppo_policy_grad(E, self.M, self.F_state, len(self.cols), self._cols, self.H, _p(states),
                                            W, Bv, _p(actions), None, None, 0.0, 0.0, 1.0, None, None, _p(self.terms),
                                            _p(self.ws), self.nbytes, _stream()), "mbo_ppo_policy_grad")
        return self.terms[:E, 0].clone(), self.terms[:E, 1].clone()

    def check(self):
        """Synchronises; raises if a pipeline wait of the update kernels timed out."""
        flag = int(self.ws[0].item())
        if flag:
            self.ws[0] = 0
            raise L.MboError("ppo_update kernels aborted: an mbarrier wait timed out (pipeline bug)")

    def debug_views(self):
        """h1..h4, dZ1..dZ4 [R, 208], logits, dlogits [R] as views of the workspace (tests / probes)."""
        off = (C.c_size_t * 12)()
        L.check(L.lib().mbo_ppo_debug_offsets(self.E, self.M, off), "mbo_ppo_debug_offsets")
        R = self.E * self.M
        f = self.ws.view(torch.float32)
        v = {}
        for i, name in enumerate(["h1", "h2", "h3", "h4", "dz1", "dz2", "dz3", "dz4"]):
            o = off[i] // 4
            v[name] = f[o:o + R * self.HP].view(R, self.HP)
        v["logits"] = f[off[8] // 4:off[8] // 4 + R]
        v["dlogits"] = f[off[9] // 4:off[9] // 4 + R]
        v["part_big"] = f[off[11] // 4:off[11] // 4 + 49 * 3 * self.HP * self.HP].view(49, 3, self.HP, self.HP)
        return v


class PPOPolicyGradSimt:
    """`mbo_ppo_policy_grad_simt`: the same call as PPOPolicyGrad for any MLP shape / activation (fp32 CUDA cores)."""

    def __init__(self, E, M, F_state, cols, widths, act, device):
        self.E, self.M, self.F_state, self.cols = int(E), int(M), int(F_state), list(cols)
        self.widths, self.act = [int(w) for w in widths], act
        assert self.widths[0] == len(self.cols) and self.widths[-1] == 1
        self.L = len(self.widths) - 1
        self._w = (C.c_int32 * (self.L + 1))(*self.widths)
        self._cols = (C.c_int32 * len(self.cols))(*self.cols)
        self.nbytes = int(L.lib().mbo_ppo_simt_workspace_bytes(self.E * self.M, self.L, self._w, self.E))
        if self.nbytes == 0:
            raise L.MboError("mbo_ppo_policy_grad_simt does not support this network: %s" % (self.widths,))
        self.ws = torch.zeros(self.nbytes // 4 + 64, dtype=torch.int32, device=device)
        self.terms = torch.empty((self.E, 4), dtype=torch.float32, device=device)

    def _call(self, states, weights, biases, actions, advs, logp_old, epsilon, ent_coeff, inv_n, dW, db):
        _chk(states, torch.float32, "states"); _chk(actions, torch.int32, "actions")
        _chk(advs, torch.float32, "advs"); _chk(logp_old, torch.float32, "logp_old")
        E = states.shape[0]
        assert E <= self.E and tuple(states.shape[1:]) == (self.M, self.F_state), states.shape
        assert len(weights) == len(biases) == self.L
        for t in list(weights) + list(biases) + list(dW or []) + list(db or []):
            _chk(t, torch.float32, "parameter / gradient")
        arr = lambda ts: None if ts is None else (C.c_void_p * self.L)(*[t.data_ptr() for t in ts])
        L.check(L.lib().mbo_ppo_policy_grad_simt(E, self.M, self.F_state, self.L, self._w, self._cols, L.ACT[self.act],
                                                 _p(states), arr(weights), arr(biases), _p(actions), _p(advs), _p(logp_old),
                                                 float(epsilon), float(ent_coeff), float(inv_n), arr(dW), arr(db),
                                                 _p(self.terms), _p(self.ws), self.nbytes, _stream()),
                "mbo_ppo_policy_grad_simt")
        return self.terms[:E]

    def __call__(self, states, weights, biases, actions, advs, logp_old, epsilon, ent_coeff, n_global, dW, db):
        assert len(dW) == len(db) == self.L
        return self._call(states, weights, biases, actions, advs, logp_old, epsilon, ent_coeff, 1.0 / float(n_global), dW, db)

    def logp_entropy(self, states, weights, biases, actions):
        t = self._call(states, weights, biases, actions, None, None, 0.0, 0.0, 1.0, None, None)
        return t[:, 0].clone(), t[:, 1].clone()

    def check(self):
        pass                        # no asynchronous pipeline in the CUDA-core kernels: nothing can time out


class PPOValueGradSimt:
    """`mbo_ppo_value_grad_simt`: value-net forward + backward for any MLP shape / activation (fp32 CUDA cores)."""

    def __init__(self, E, widths, act, device):
        self.E, self.widths, self.act = int(E), [int(w) for w in widths], act
        assert self.widths[0] == 2 and self.widths[-1] == 1
        self.L = len(self.widths) - 1
        self._w = (C.c_int32 * (self.L + 1))(*self.widths)
        self.nbytes = int(L.lib().mbo_ppo_simt_workspace_bytes(self.E, self.L, self._w, self.E))
        if self.nbytes == 0:
            raise L.MboError("mbo_ppo_value_grad_simt does not support this network: %s" % (self.widths,))
        self.ws = torch.zeros(self.nbytes // 4 + 64, dtype=torch.int32, device=device)
        self.values = torch.empty(self.E, dtype=torch.float32, device=device)
        self.vterms = torch.empty(self.E, dtype=torch.float32, device=device)

    def __call__(self, states, t_col, T_col, weights, biases, tdlamrets, coeff, dW, db):
        _chk(states, torch.float32, "states"); _chk(tdlamrets, torch.float32, "tdlamrets")
        E, M, F = states.shape
        assert E <= self.E and len(weights) == len(biases) == len(dW) == len(db) == self.L
        for t in list(weights) + list(biases) + list(dW) + list(db):
            _chk(t, torch.float32, "parameter / gradient")
        arr = lambda ts: (C.c_void_p * self.L)(*[t.data_ptr() for t in ts])
        L.check(L.lib().mbo_ppo_value_grad_simt(E, M * F, int(t_col) % F, int(T_col) % F, self.L, self._w, L.ACT[self.act],
                                                _p(states), arr(weights), arr(biases), _p(tdlamrets), float(coeff), arr(dW),
                                                arr(db), _p(self.values), _p(self.vterms), _p(self.ws), self.nbytes, _stream()),
                "mbo_ppo_value_grad_simt")
        return self.values[:E], self.vterms[:E]


def tensor_core_policy_update_ok(widths, act):
    """shape served by the tcgen05 update kernels (mbo_ppo_policy_grad): F -> H -> H -> H -> H -> 1 ReLU, 8 <= H <= 207, F <= 8"""
    w = list(widths)
    return act == "relu" and len(w) == 6 and len(set(w[1:5])) == 1 and 8 <= w[1] <= 207 and 1 <= w[0] <= 8 and w[5] == 1


def tensor_value_update_ok(widths, act):
    w = list(widths)
    return act == "relu" and len(w) == 6 and len(set(w[1:5])) == 1 and 1 <= w[1] <= 256 and w[0] == 2 and w[5] == 1


def make_policy_grad(E, M, F_state, cols, widths, act, device):
    """the library's kernel set for this policy-net shape: tcgen05 (3xTF32 forward / TF32 backward) for the shipped
    4 x H ReLU nets, fp32 CUDA cores for every other MLP -- both behind the same call"""
    if tensor_core_policy_update_ok(widths, act):
        return PPOPolicyGrad(E, M, F_state, cols, widths[1], device)
    return PPOPolicyGradSimt(E, M, F_state, cols, widths, act, device)


def make_value_grad(E, widths, act, device):
    if tensor_value_update_ok(widths, act):
        return PPOValueGrad(E, widths[1], device)
    return PPOValueGradSimt(E, widths, act, device)


def ppo_prepare_minibatch(sel, flat, normalize, out, adv_stats=None):
    """Gathers the transitions `sel` (int64 device indices) of the flattened batch `flat` (dict: states [n,M,F] f32,
    actions [n] i32, advs / tdlamrets / logprobs_old [n] f32) into the preallocated minibatch tensors `out` and
    normalises the advantages (ppo.py:141-144); `adv_stats` (3 doubles on the device): moments of the GLOBAL minibatch
    when it is sharded over ranks."""
    _chk(sel, torch.int64, "sel"); _chk(adv_stats, torch.float64, "adv_stats")
    E = sel.shape[0]
    st = flat["states"]
    lpo = flat.get("logprobs_old")
    L.check(L.lib().mbo_ppo_prepare_minibatch(E, st.shape[1] * st.shape[2], _p(sel), _p(st), _p(flat["actions"]),
                                              _p(flat["advs"]), _p(flat["tdlamrets"]), _p(lpo), int(bool(normalize)),
                                              _p(adv_stats), _p(out["states"]), _p(out["actions"]), _p(out["advs"]),
                                              _p(out["tdlamrets"]), _p(out["logprobs_old"]), _stream()),
            "mbo_ppo_prepare_minibatch")
    return out


def ppo_adv_stats(bounds, sel, advs_all, out):
    """out [n_steps, 3] f64 = {sum, sum of squares, count} of advs_all[sel[bounds[j]:bounds[j+1]]] (local shard of every
    minibatch of an epoch)."""
    _chk(bounds, torch.int32, "bounds"); _chk(sel, torch.int64, "sel"); _chk(advs_all, torch.float32, "advs_all")
    _chk(out, torch.float64, "out")
    n_steps = bounds.numel() - 1
    assert out.numel() >= 3 * n_steps
    L.check(L.lib().mbo_ppo_adv_stats(n_steps, _p(bounds), _p(sel), _p(advs_all), _p(out), _stream()), "mbo_ppo_adv_stats")
    return out


class Comm:
    """Peer-memory exchange of the data-parallel PPO update (`mbo_comm`, csrc/comm.cu): one region per rank, mapped into
    every process through CUDA IPC handles that travel over the torch.distributed process group on the host."""

    def __init__(self, rank, world, slot_bytes, device=None):
        self.rank, self.world, self.slot_bytes = int(rank), int(world), int(slot_bytes)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(L.lib().mbo_comm_create(self.rank, self.world, self.slot_bytes, C.byref(h)), "mbo_comm_create")
        self.h = h

    def export(self):
        buf = C.create_string_buffer(L.COMM_HANDLE_BYTES)
        L.check(L.lib().mbo_comm_export(self.h, buf), "mbo_comm_export")
        return bytes(buf.raw)

    def connect(self, handles):
        assert len(handles) == self.world and all(len(x) == L.COMM_HANDLE_BYTES for x in handles)
        blob = C.create_string_buffer(b"".join(handles), L.COMM_HANDLE_BYTES * self.world)
        with torch.cuda.device(self.device):
            L.check(L.lib().mbo_comm_connect(self.h, blob), "mbo_comm_connect")

    @classmethod
    def from_process_group(cls, slot_bytes, device=None, group=None):
        """collective: every rank of the (initialised) process group calls this"""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        c = cls(rank, world, slot_bytes, device)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, c.export(), group=group)
            c.connect(handles)
            dist.barrier(group)
        return c

    def allreduce(self, src, dst=None):
        """dst = sum over ranks of src (fp32 / fp64 contiguous CUDA tensors); in place when dst is None"""
        dst = src if dst is None else dst
        assert src.is_cuda and src.is_contiguous() and dst.is_contiguous() and src.dtype == dst.dtype
        assert src.numel() == dst.numel()
        dt = {torch.float32: 0, torch.float64: 1}[src.dtype]
        L.check(L.lib().mbo_comm_allreduce(self.h, dt, _p(src), _p(dst), src.numel(), _stream()), "mbo_comm_allreduce")
        return dst

    def status(self):
        ep, err = C.c_ulonglong(0), C.c_uint(0)
        L.check(L.lib().mbo_comm_status(self.h, C.byref(ep), C.byref(err), _stream()), "mbo_comm_status")
        return int(ep.value), int(err.value)

    def check(self):
        ep, err = self.status()
        if err:
            raise L.MboError("mbo_comm: a wait on a peer rank timed out (rank %d, %d completed calls): the ranks did not "
                             "enqueue the same sequence of exchange calls" % (self.rank, ep))
        return ep

    def __del__(self):
        try:
            if self.h:
                L.lib().mbo_comm_destroy(self.h)
                self.h = None
        except Exception:
            pass


def ppo_loss_sums(terms, vterms, n_global, value_coeff, ent_coeff, out):
    E = terms.shape[0]
    L.check(L.lib().mbo_ppo_loss_sums(E, _p(terms), _p(vterms), 1.0 / float(n_global), float(value_coeff), float(ent_coeff),
                                      _p(out), _stream()), "mbo_ppo_loss_sums")
    return out


class PPOValueGrad:
    """Workspace + call wrapper of `mbo_ppo_value_grad` (value-net forward + backward of one optimiser step)."""

    def __init__(self, E, Hv, device):
        self.E, self.Hv = int(E), int(Hv)
        self.nbytes = int(L.lib().mbo_ppo_value_workspace_bytes(self.E, self.Hv))
        self.ws = torch.zeros(self.nbytes // 4 + 64, dtype=torch.float32, device=device)
        self.values = torch.empty(self.E, dtype=torch.float32, device=device)
        self.vterms = torch.empty(self.E, dtype=torch.float32, device=device)

    def __call__(self, states, t_col, T_col, weights, biases, tdlamrets, coeff, dW, db):
        """states [E, M, F] fp32 contiguous; returns (values [E], (values - tdlamrets)^2 [E]); gradients of
        coeff * sum((values - tdlamrets)^2) are written into dW / db."""
        _chk(states, torch.float32, "states"); _chk(tdlamrets, torch.float32, "tdlamrets")
        E, M, F = states.shape
        assert E <= self.E and len(weights) == len(biases) == len(dW) == len(db) == 5
        for t in list(weights) + list(biases) + list(dW) + list(db):
            _chk(t, torch.float32, "parameter / gradient")
        arr = lambda ts: (C.c_void_p * 5)(*[t.data_ptr() for t in ts])
        L.check(L.lib().mbo_ppo_value_grad(E, M * F, int(t_col) % F, int(T_col) % F, self.Hv, _p(states), arr(weights),
                                           arr(biases), _p(tdlamrets), float(coeff), arr(dW), arr(db), _p(self.values),
                                           _p(self.vterms), _p(self.ws), self.nbytes, _stream()), "mbo_ppo_value_grad")
        return self.values[:E], self.vterms[:E]


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr) semantics (ppo.py:93: default betas / eps, no weight decay, no amsgrad) with the
    whole step in one kernel launch (`mbo_adam_step`); the step counters live on the device, so a captured CUDA graph of
    the step can be replayed.  state[p] has torch's keys: step (device int64, one per parameter), exp_avg, exp_avg_sq."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.comm = None          # ops.Comm: the step all-reduces the flat gradient over the ranks inside the Adam kernel
        self.flat_grad = None     # the ONE fp32 buffer all .grad tensors are views of (in parameter order)

    def set_comm(self, comm, flat_grad):
        """data-parallel mode: `flat_grad` is the concatenation of every parameter's gradient in param_groups order"""
        self.comm, self.flat_grad = comm, flat_grad

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                st = self.state[p]
                if not st:
                    st["_dev"] = torch.zeros(3, dtype=torch.int64, device=p.device)     # step, 1 - b1^t, 1 - b2^t (doubles)
                    st["step"] = st["_dev"][0:1]
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                _chk(p.data, torch.float32, "parameter"); _chk(p.grad, torch.float32, "gradient")
            n = len(ps)
            arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
            sizes = (C.c_int64 * n)(*[p.numel() for p in ps])
            b1, b2 = group["betas"]
            if self.comm is not None and self.comm.world > 1:
                off = 0
                for p in ps:        # the gradients must tile the flat buffer in order (PPO._alloc_native_grads)
                    assert p.grad.data_ptr() == self.flat_grad.data_ptr() + 4 * off, "gradients are not views of flat_grad"
                    off += p.numel()
                assert off == self.flat_grad.numel() and len(self.param_groups) == 1
                L.check(L.lib().mbo_comm_allreduce_adam(self.comm.h, _p(self.flat_grad), None, n, arr([p.data for p in ps]),
                                                        arr([self.state[p]["exp_avg"] for p in ps]),
                                                        arr([self.state[p]["exp_avg_sq"] for p in ps]), sizes,
                                                        arr([self.state[p]["_dev"] for p in ps]),
                                                        float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                                        _stream()), "mbo_comm_allreduce_adam")
                for p in ps:
                    p._mbo_version = getattr(p, "_mbo_version", 0) + 1
                continue
            L.check(L.lib().mbo_adam_step(n, arr([p.data for p in ps]), arr([p.grad for p in ps]),
                                          arr([self.state[p]["exp_avg"] for p in ps]),
                                          arr([self.state[p]["exp_avg_sq"] for p in ps]), sizes,
                                          arr([self.state[p]["_dev"] for p in ps]),
                                          float(group["lr"]), float(b1), float(b2), float(group["eps"]), _stream()),
                    "mbo_adam_step")
            for p in ps:        # the kernel writes through raw pointers: torch's version counter does not move, so
                p._mbo_version = getattr(p, "_mbo_version", 0) + 1      # NeuralAF._version() reads this one as well
        return None
