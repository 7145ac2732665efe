"""GPU parity of the CUDA-core PPO update kernels (csrc/ppo_update_simt.cu: `mbo_ppo_policy_grad_simt`,
`mbo_ppo_value_grad_simt`) -- the optimiser step of ppo.py:129-178 for every network the tensor-core kernels do not cover
(train_metabo_gps.py:62-94's 4 -> 20 -> 20 -> 1 policy / 2 -> 20 -> 20 -> 1 value net, tanh policies policies.py:69-74) --
against torch-CPU fp64 autograd of the same loss (policies.py:150-161 Categorical terms, ppo.py:148-168).
Tolerance: fp32 arithmetic, 1e-4 of max|.| per tensor; log-probs 2e-5 absolute-relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def _mlp64(x, Ws, bs, act):
    f = torch.relu if act == "relu" else torch.tanh
    for W, b in zip(Ws[:-1], bs[:-1]):
        x = f(x @ W.t() + b)
    return x @ Ws[-1].t() + bs[-1]


CASES = [("relu", [4, 20, 20, 1], [0, 1, 2, 3], 6, 5, 333), ("tanh", [4, 20, 20, 1], [0, 1, 2, 3], 6, 5, 333),
         ("tanh", [6, 17, 9, 33, 1], [0, 1, 2, 3, 4, 5], 6, 3, 2100), ("relu", [3, 64, 1], [0, 2, 5], 7, 7, 96),
         ("tanh", [7, 200, 200, 200, 200, 1], list(range(7)), 7, 2, 300)]


@pytest.mark.parametrize("act,widths,cols,F_state,E,M", CASES)
def test_simt_policy_grad_matches_fp64_autograd(act, widths, cols, F_state, E, M):
    from metabo_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(len(widths) * 100 + E)
    L = len(widths) - 1
    W = [torch.randn(widths[i + 1], widths[i], generator=g) * (0.6 / np.sqrt(widths[i])) for i in range(L)]
    b = [torch.randn(widths[i + 1], generator=g) * 0.1 for i in range(L)]
    states = torch.randn(E, M, F_state, generator=g)
    actions = torch.randint(0, M, (E,), generator=g)
    advs = torch.randn(E, generator=g)
    eps, entc, n_global = 0.15, 0.01, E + 2
    Wr = [w.double().clone().requires_grad_() for w in W]
    br = [x.double().clone().requires_grad_() for x in b]

    def terms64(lpo):
        logits = _mlp64(states[:, :, cols].double().reshape(E * M, len(cols)), Wr, br, act).reshape(E, M)
        norm = logits - torch.logsumexp(logits, -1, keepdim=True)
        lp = norm.gather(-1, actions[:, None]).squeeze(-1)
        ent = -(norm * norm.exp()).sum(-1)
        if lpo is None:
            return lp, ent, None
        ratio = torch.exp(lp - lpo)
        loss = (-torch.min(ratio * advs.double(), ratio.clamp(1 - eps, 1 + eps) * advs.double()).sum() - entc * ent.sum()) / n_global
        return lp, ent, loss

    lp0, _, _ = terms64(None)
    lpo = (lp0.detach() + 0.2 * torch.randn(E, generator=g).double())
    lp, ent, loss = terms64(lpo)
    loss.backward()
    pg = ops.make_policy_grad(E + 1, M, F_state, cols, widths, act, dev)
    assert isinstance(pg, ops.PPOPolicyGradSimt)
    Wd, bd = [w.to(dev).contiguous() for w in W], [x.to(dev).contiguous() for x in b]
    dW, db = [torch.full_like(w, 5.0) for w in Wd], [torch.full_like(x, 5.0) for x in bd]
    a32 = actions.to(torch.int32).to(dev)
    lp_k, ent_k = pg.logp_entropy(states.to(dev), Wd, bd, a32)
    terms = pg(states.to(dev), Wd, bd, a32, advs.to(dev), lpo.float().to(dev), eps, entc, n_global, dW, db).clone()
    torch.cuda.synchronize()
    assert float((lp_k.double().cpu() - lp.detach()).abs().max()) <= 2e-5 * max(1.0, float(lp.abs().max()))
    assert float((ent_k.double().cpu() - ent.detach()).abs().max()) <= 2e-5 * max(1.0, float(ent.abs().max()))
    assert torch.equal(terms[:, 0], lp_k) and torch.equal(terms[:, 1], ent_k)      # forward-only == gradient call
    for l in range(L):
        assert _rel(dW[l], Wr[l].grad) <= 1e-4, ("dW", l, _rel(dW[l], Wr[l].grad))
        if l < L - 1:
            assert _rel(db[l], br[l].grad) <= 1e-4, ("db", l)
    # the logit bias shifts every logit of an episode alike: its gradient is analytically 0 (softmax invariance) and
    # the kernels' value is the fp32 round-off of a sum of R cancelling dlogits
    assert float(db[L - 1].abs().max()) <= 1e-5 * float(dW[L - 1].abs().max()) + 1e-7
    # bitwise reproducible (fixed-order partial sums, no atomics)
    dW2, db2 = [torch.zeros_like(w) for w in Wd], [torch.zeros_like(x) for x in bd]
    pg(states.to(dev), Wd, bd, a32, advs.to(dev), lpo.float().to(dev), eps, entc, n_global, dW2, db2)
    torch.cuda.synchronize()
    assert all(torch.equal(x, y) for x, y in zip(dW + db, dW2 + db2))


@pytest.mark.parametrize("act,widths,E", [("relu", [2, 20, 20, 1], 60), ("tanh", [2, 20, 20, 1], 7), ("tanh", [2, 200, 200, 200, 200, 1], 33),
                                          ("relu", [2, 300, 1], 2500)])
def test_simt_value_grad_matches_fp64_autograd(act, widths, E):
    from metabo_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(E)
    L = len(widths) - 1
    M, F = 9, 6
    W = [torch.randn(widths[i + 1], widths[i], generator=g) * (0.5 / np.sqrt(widths[i])) for i in range(L)]
    b = [torch.randn(widths[i + 1], generator=g) * 0.1 for i in range(L)]
    states = torch.randn(E, M, F, generator=g)
    states[:, :, -2] = torch.randint(0, 30, (E, 1), generator=g).float() / 10
    states[:, :, -1] = 3.0
    ret = torch.randn(E, generator=g)
    coeff = 1.0 / (2 * E)
    Wr = [w.double().clone().requires_grad_() for w in W]
    br = [x.double().clone().requires_grad_() for x in b]
    v = _mlp64(torch.stack([states[:, 0, -2], states[:, 0, -1]], 1).double(), Wr, br, act).squeeze(1)
    (coeff * ((v - ret.double()) ** 2).sum()).backward()
    vg = ops.make_value_grad(E + 3, widths, act, dev)
    assert isinstance(vg, ops.PPOValueGradSimt)
    Wd, bd = [w.to(dev).contiguous() for w in W], [x.to(dev).contiguous() for x in b]
    dW, db = [torch.full_like(w, 3.0) for w in Wd], [torch.full_like(x, 3.0) for x in bd]
    vals, vterms = vg(states.to(dev), -2, -1, Wd, bd, ret.to(dev), coeff, dW, db)
    torch.cuda.synchronize()
    assert _rel(vals, v.detach()) <= 1e-5
    assert _rel(vterms, ((v - ret.double()) ** 2).detach()) <= 1e-4
    for l in range(L):
        assert _rel(dW[l], Wr[l].grad) <= 1e-4, ("dW", l)
        assert _rel(db[l], br[l].grad) <= 1e-4, ("db", l)


def test_ppo_training_gps_config_runs_on_native_kernels(tmp_path):
    """train_metabo_gps.py (C4): 4 -> 20 -> 20 -> 1 ReLU policy with t / T masked from the policy, 2 -> 20 -> 20 -> 1 value
    net, Matern-5/2 GP-sample objectives on a discrete Sobol domain.  PPO.train must run the optimiser step on the
    library's CUDA-core kernels (no torch-autograd path exists) and follow the all-torch comparison implementation
    (tests/autograd_ppo.py) from the same seeds: same rollouts in iteration 0, logged losses within 2 %."""
    from metabo_b200 import gym_compat as gym, ops
    from metabo_b200.policies.policies import NeuralAF
    from metabo_b200.ppo.ppo import PPO
    from autograd_ppo import AutogradPPO
    spec = {"env_id": "MetaBO-GP-C4-v0", "D": 3, "f_type": "GP",
            "f_opts": {"kernel": "Matern52", "lengthscale_low": 0.05, "lengthscale_high": 0.5, "noise_var_low": 0.1,
                       "noise_var_high": 0.1, "signal_var_low": 1.0, "signal_var_high": 1.0, "min_regret": 1e-6},
            "features": ["posterior_mean", "posterior_std", "incumbent", "timestep_perc", "timestep", "budget"],
            "T": 6, "n_init_samples": 0, "pass_X_to_pi": False, "kernel": "Matern52", "kernel_lengthscale": None,
            "kernel_variance": None, "noise_variance": None, "use_prior_mean_function": True, "local_af_opt": False,
            "cardinality_domain": 300, "reward_transformation": "neg_log10"}
    gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=6, kwargs=spec)

    def run(cls, name):
        params = {"batch_size": 96, "max_steps": 192, "minibatch_size": 24, "n_epochs": 2, "lr": 1e-3, "epsilon": 0.15,
                  "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98, "loss_type": "GAElam",
                  "normalize_advs": True, "n_workers": 4, "env_id": spec["env_id"], "seed": 0, "env_seeds": [0, 1, 2, 3],
                  "policy_options": {"activations": "relu", "arch_spec": [20, 20], "use_value_network": True,
                                     "arch_spec_value": [20, 20], "exclude_t_from_policy": True, "exclude_T_from_policy": True,
                                     "t_idx": -2, "T_idx": -1}}
        ppo = cls(policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=params["policy_options"]), params=params,
                  logpath=str(tmp_path / name), save_interval=10, verbose=True)
        w0 = torch.cat([p.detach().reshape(-1).clone() for p in ppo.pi.parameters()])
        ppo.train()
        w1 = torch.cat([p.detach().reshape(-1) for p in ppo.pi.parameters()])
        losses = [[float(x.split("=")[1]) for x in line.split(",")] for line in ppo.iter_log_str.splitlines() if "loss_ppo" in line]
        return ppo, w0, w1, np.array(losses)

    pa, w0a, w1a, la = run(AutogradPPO, "autograd")
    pn, w0n, w1n, ln = run(PPO, "native")
    assert isinstance(pn._pg, ops.PPOPolicyGradSimt) and isinstance(pn._vg, ops.PPOValueGradSimt) and pa._pg is None
    assert pn._ep is not None                                   # second iteration ran as a captured epoch graph
    assert torch.equal(w0a, w0n)
    assert la.shape == ln.shape == (2, 4)
    np.testing.assert_allclose(ln, la, rtol=2e-2, atol=2e-4)
    da, dn = (w1a - w0a).double(), (w1n - w0n).double()
    assert float(torch.nn.functional.cosine_similarity(da, dn, dim=0)) > 0.9
