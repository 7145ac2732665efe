"""NeuralAF drop-in (metabo/policies/policies.py:33-178).

Same constructor, state_dict keys, shapes and return conventions as the reference; the no-grad
paths (`af`, `act`, `forward` under torch.no_grad, `predict_vals_logps_ents` of the frozen old
policy) run on the sm_100a kernels through the C ABI, the autograd path (PPO update) uses torch
ops on the GPU.  There is no CPU inference fallback: states are moved to the CUDA device."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib as L
from .. import ops
from .mlp import MLP

_MODES = L.MLP_MODES


class NeuralAF(nn.Module):
    """
    SHAPES (as the reference):
    forward():  states (N_batch, N_grid, N_features) -> logits (N_batch, N_grid), values (N_batch,)
    act():      state (N_grid, N_features) -> action (), value ()
    predict_vals_logps_ents(): states, actions (N_batch,) -> values, logprobs, entropies (N_batch,)
    """

    def __init__(self, observation_space, action_space, deterministic, options):
        super().__init__()
        self.N_features = None
        self.deterministic = deterministic
        self.init_structure(observation_space=observation_space, action_space=action_space, options=options)
        self.apply(self.init_weights)
        self.mlp_mode = options.get("mlp_mode", "bf16x3")   # kernel arithmetic of the no-grad paths
        self._h_policy = None
        self._h_value = None
        self._packed_version = None
        self.sample_seed = int(options.get("sample_seed", 0))
        self._sample_step = 0

    def init_structure(self, observation_space, action_space, options):
        self.N_features = observation_space.shape[1]
        if options["activations"] == "relu":
            f_act = F.relu
        elif options["activations"] == "tanh":
            f_act = torch.tanh
        else:
            raise NotImplementedError("Unknown activation function!")
        self.activation = options["activations"]
        self.N_features_policy = self.N_features
        if "exclude_t_from_policy" in options:
            self.exclude_t_from_policy = options["exclude_t_from_policy"]
            assert "t_idx" in options
            self.t_idx = options["t_idx"]
            self.N_features_policy -= 1 if self.exclude_t_from_policy else 0
        else:
            self.exclude_t_from_policy = False
        if "exclude_T_from_policy" in options:
            self.exclude_T_from_policy = options["exclude_T_from_policy"]
            assert "T_idx" in options
            self.T_idx = options["T_idx"]
            self.N_features_policy -= 1 if self.exclude_T_from_policy else 0
        else:
            self.exclude_T_from_policy = False
        self.policy_net = MLP(d_in=self.N_features_policy, d_out=1, arch_spec=options["arch_spec"], f_act=f_act)
        if "use_value_network" in options and options["use_value_network"]:
            self.use_value_network = True
            self.value_net = MLP(d_in=2, d_out=1, arch_spec=options["arch_spec_value"], f_act=f_act)
            self.t_idx = options["t_idx"]
            self.T_idx = options["T_idx"]
        else:
            self.use_value_network = False

    # -- helpers ---------------------------------------------------------------------------------
    def policy_mask(self):
        mask = [True] * self.N_features
        if self.exclude_t_from_policy:
            mask[self.t_idx] = False
        if self.exclude_T_from_policy:
            mask[self.T_idx] = False
        return mask

    def policy_columns(self):
        return [i for i, m in enumerate(self.policy_mask()) if m]

    def _version(self):
        # _mbo_version: bumped by ops.FusedAdam (its kernel updates the parameters through raw pointers)
        return tuple((p._version, getattr(p, "_mbo_version", 0), p.data_ptr()) for p in self.parameters())

    def __deepcopy__(self, memo):
        """copy.deepcopy(pi) (ppo.py:203 hands a copy to the workers): the kernel handles own device memory and are
        not copied -- the copy re-packs its own on first use."""
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_h_policy", "_h_value", "_packed_version"):
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def kernel_handles(self):
        """(policy handle, value handle or None), re-packed whenever a parameter changed."""
        dev = torch.device("cuda", torch.cuda.current_device())
        v = self._version()
        if self._h_policy is None or v != self._packed_version:
            if self._h_policy is None:
                self._h_policy = ops.PolicyHandle()
            self._h_policy.set_weights([l.weight.to(dev) for l in self.policy_net.fc],
                                       [l.bias.to(dev) for l in self.policy_net.fc], self.activation)
            if self.use_value_network:
                if self._h_value is None:
                    self._h_value = ops.PolicyHandle()
                self._h_value.set_weights([l.weight.to(dev) for l in self.value_net.fc],
                                          [l.bias.to(dev) for l in self.value_net.fc], self.activation)
            self._packed_version = v
        return self._h_policy, self._h_value

    def kernel_mode(self):
        h, _ = self.kernel_handles()
        mode = _MODES[self.mlp_mode]
        if mode != L.MLP_FP32 and not self.tensor_core_ok():
            mode = L.MLP_FP32
        if mode == L.MLP_FP16 and self.policy_net.arch_spec[0] > 206:
            mode = L.MLP_TF32          # fp16 mode needs two spare K columns; tf32 has the same significand
        return mode

    def tensor_core_ok(self):
        a = self.policy_net.arch_spec
        return (self.activation == "relu" and len(a) == 4 and len(set(a)) == 1 and 8 <= a[0] <= 208
                and self.N_features_policy <= 15)

    def _values_kernel(self, states):
        if not self.use_value_network:
            return torch.zeros(states.shape[0], device=states.device)
        _, hv = self.kernel_handles()
        tT = torch.stack([states[:, 0, self.t_idx], states[:, 0, self.T_idx]], dim=1).contiguous()
        return ops.value_net(hv, tT)

    # -- reference API ---------------------------------------------------------------------------
    def forward(self, states):
        assert states.dim() == 3
        assert states.shape[-1] == self.N_features
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            # autograd path (PPO update): torch ops on the parameters' device; no host->device index tensors are
            # created here so that the whole update step can be captured in a CUDA graph
            cols = self.policy_columns()
            if len(cols) == self.N_features:
                pin = states
            else:
                lo, hi = cols[0], cols[-1] + 1
                pin = states[:, :, lo:hi] if cols == list(range(lo, hi)) else torch.stack([states[:, :, c] for c in cols], dim=2)
            logits = self.policy_net.forward(pin)
            logits = logits.squeeze(2)
            if self.use_value_network:
                tT = torch.stack([states[:, 0, self.t_idx], states[:, 0, self.T_idx]], dim=1)
                values = self.value_net.forward(tT).squeeze(1)
            else:
                values = torch.zeros(states.shape[0]).to(logits.device)
            return logits, values
        dev = torch.device("cuda", torch.cuda.current_device())
        src_dev = states.device
        s = states.to(dev, dtype=torch.float32).contiguous()
        hp, _ = self.kernel_handles()
        logits = ops.af_logits_dense(hp, self.policy_columns(), s, mode=self.kernel_mode())
        values = self._values_kernel(s)
        return logits.to(src_dev), values.to(src_dev)

    def af(self, state):
        state = torch.from_numpy(state[None, :].astype(np.float32))
        with torch.no_grad():
            out = self.forward(state)
        return out[0].to("cpu").numpy().squeeze()

    def act(self, state):
        state = state.unsqueeze(0)
        with torch.no_grad():
            dev = torch.device("cuda", torch.cuda.current_device())
            s = state.to(dev, dtype=torch.float32).contiguous()
            hp, _ = self.kernel_handles()
            logits = ops.af_logits_dense(hp, self.policy_columns(), s, mode=self.kernel_mode())
            value = self._values_kernel(s)
            if self.deterministic:
                action, _ = ops.logits_argmax(logits)
            else:
                action = ops.logits_sample(logits, self.sample_seed, self._sample_step)
                self._sample_step += 1
        return action.to(torch.int64).squeeze(0).to(state.device), value.squeeze(0).to(state.device)

    def predict_vals_logps_ents(self, states, actions):
        assert actions.dim() == 1
        assert states.shape[0] == actions.shape[0]
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            logits, values = self.forward(states)
            norm = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
            logprobs = norm.gather(-1, actions.long()[:, None]).squeeze(-1)
            entropies = -(norm * torch.exp(norm)).sum(-1)
            return values, logprobs, entropies
        logits, values = self.forward(states)
        dev = torch.device("cuda", torch.cuda.current_device())
        lp, en = ops.logits_logp_entropy(logits.to(dev).contiguous(), actions.to(dev).to(torch.int32).contiguous())
        return values, lp.to(states.device), en.to(states.device)

    def set_requires_grad(self, requires_grad):
        for p in self.parameters():
            p.requires_grad = requires_grad

    def reset(self):
        pass

    @staticmethod
    def num_flat_features(x):
        return np.prod(x.size()[1:])

    @staticmethod
    def init_weights(m):
        if type(m) == nn.Linear:
            m.weight.data.normal_(mean=0.0, std=0.01)
            m.bias.data.fill_(0.0)


class _ClosedFormAF:
    """Shared part of the baseline AFs (metabo/policies/policies.py:181-307): same constructor arguments, `af`,
    `act`, `set_requires_grad`, `reset` as the reference; evaluated by `mbo_af_closed_form` on the GPU.  Bound to a
    VecMetaBO (set_policy) they replace the NeuralAF inside the AF-optimisation rounds; `af` / `act` on a
    materialised state keep the reference's calling convention.  Ties: lowest index (the reference draws uniformly
    among exact ties from numpy's global RNG)."""
    kind = None

    def engine_spec(self):
        raise NotImplementedError

    def _columns(self, state):
        fo = self.feature_order
        st = np.asarray(state, dtype=np.float32)
        dev = torch.device("cuda", torch.cuda.current_device())
        mean = torch.as_tensor(st[None, :, fo.index("posterior_mean")].copy(), device=dev)
        std = torch.as_tensor(st[None, :, fo.index("posterior_std")].copy(), device=dev)
        epi = torch.zeros((1, 4), dtype=torch.float32, device=dev)
        if "incumbent" in fo:
            epi[0, 0] = float(st[0, fo.index("incumbent")])
        if "timestep" in fo:
            epi[0, 2] = float(st[0, fo.index("timestep")])
        return mean, std, epi

    def _af_device(self, state):
        mean, std, epi = self._columns(state)
        kind, p0, p1 = self.engine_spec()
        return ops.af_closed_form(kind, mean, std, epi, p0=p0, p1=p1, D=getattr(self, "D", 0) or 0)

    def af(self, state):
        return self._af_device(state)[0].to("cpu").numpy().astype(self.out_dtype)

    def act(self, state):
        st = state.numpy() if isinstance(state, torch.Tensor) else state
        a, _ = ops.logits_argmax(self._af_device(st))
        return a[0].to("cpu", torch.int64), torch.tensor(0.0)

    def set_requires_grad(self, flag):
        pass

    def reset(self):
        pass


class UCB(_ClosedFormAF):
    out_dtype = np.float32

    def __init__(self, feature_order, kappa, D=None, delta=None):
        self.feature_order = list(feature_order)
        self.kappa = kappa
        self.D = D
        self.delta = delta
        assert not (self.kappa == "gp_ucb" and self.D is None)
        assert not (self.kappa == "gp_ucb" and self.delta is None)

    def engine_spec(self):
        if self.kappa == "gp_ucb":
            return L.AF_UCB, 0.0, float(self.delta)
        return L.AF_UCB, float(self.kappa), 0.0


class EI(_ClosedFormAF):
    out_dtype = np.float64

    def __init__(self, feature_order):
        self.feature_order = list(feature_order)

    def engine_spec(self):
        return L.AF_EI, 0.0, 0.0


class PI(_ClosedFormAF):
    out_dtype = np.float64

    def __init__(self, feature_order, xi):
        self.feature_order = list(feature_order)
        self.xi = xi

    def engine_spec(self):
        return L.AF_PI, float(self.xi), 0.0
