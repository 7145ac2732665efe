"""GPU parity tests of the hand-written PPO policy-gradient kernels (mbo_ppo_policy_grad)
against the oracle's autograd restatement of ppo.py:148-168 / policies.py:150-161.

Tolerances.  Forward: 3xTF32 (hi/lo split of both operands, fp32 accumulation), i.e. fp32-level: logits, log-probs,
ratios and dlogits within 2e-5 of max|.| of the fp64 oracle (north-star bar for logits: 1e-3); the stored
activations are the tf32-rounded hi parts (5e-4).  Backward: single-pass TF32 GEMMs with fp32 accumulation;
gradients are compared (a) element-wise against an fp64 backward under the KERNEL'S OWN ReLU decisions (the handful
of pre-activations within fp32 rounding error of zero may flip their ReLU against the fp64 forward and change dZ by
its full value) at 3e-3 of max|grad| per tensor for random nets and 2e-2 for the trained policy (its gradient is a
small sum of large cancelling terms: the error is 2^-12 relative to the terms, not to the sum), and (b) against the
plain fp64 autograd gradient by cosine similarity (> 0.9995)."""
import os

import numpy as np
import pytest
import torch

from oracle import metabo_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def _run_case(E, M, F_state, cols, H, layers, states, actions, advs, logp_old, eps=0.15, entc=0.01, n_global=None):
    from metabo_b200 import ops
    dev = torch.device("cuda:0")
    n_global = n_global or E
    pg = ops.PPOPolicyGrad(E, M, F_state, cols, H, dev)
    Wd = [torch.as_tensor(w, dtype=torch.float32, device=dev).contiguous() for w, _ in layers]
    bd = [torch.as_tensor(b, dtype=torch.float32, device=dev).contiguous() for _, b in layers]
    dW = [torch.full_like(w, 7.0) for w in Wd]        # outputs are overwritten, not accumulated
    db = [torch.full_like(b, 7.0) for b in bd]
    args = (torch.as_tensor(states, dtype=torch.float32, device=dev).contiguous(), Wd, bd,
            torch.as_tensor(actions, dtype=torch.int32, device=dev), torch.as_tensor(advs, dtype=torch.float32, device=dev),
            torch.as_tensor(logp_old, dtype=torch.float32, device=dev), eps, entc, n_global)
    terms = pg(*args, dW, db).clone()
    torch.cuda.synchronize()
    pg.check()
    v = {k: t.clone() for k, t in pg.debug_views().items() if k != "part_big"}
    dW2 = [torch.zeros_like(w) for w in Wd]
    db2 = [torch.zeros_like(b) for b in bd]
    pg(*args, dW2, db2)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(dW + db, dW2 + db2)), "run-to-run difference (split-K must be deterministic)"
    return terms.cpu(), v, [g.cpu() for g in dW], [g.cpu() for g in db]


def _check_against_oracle(E, M, cols, H, layers, states, actions, advs, logp_old, terms, v, dW, db, eps=0.15, entc=0.01,
                          n_global=None, fwd_tol=2e-5, grad_tol=3e-3, ratio_tol=1e-4, cos_min=0.9995):
    n_global = n_global or E
    pol_states = np.asarray(states)[:, :, cols]
    ref = O.ppo_policy_grads(layers, pol_states, actions, advs, logp_old, eps, entc, n_global)
    # forward
    for l in range(4):
        h = v["h%d" % (l + 1)].cpu()
        assert _rel(h[:, :H], ref["hs"][l]) <= 5e-4, ("h", l + 1)        # stored hi part: tf32 rounding of the fp32 value
        assert bool((h[:, H] == 1).all()) and bool((h[:, H + 1:] == 0).all())
    assert _rel(v["logits"].cpu(), ref["logits"].reshape(-1)) <= fwd_tol
    assert _rel(terms[:, 0], ref["logp"]) <= fwd_tol and _rel(terms[:, 1], ref["entropy"]) <= 1e-4
    assert _rel(terms[:, 3], ref["ratio"]) <= ratio_tol      # exp(logp - logp_old): absolute logit error, so |logit|-scaled
    assert abs(float(terms[:, 2].sum()) - ref["loss_ppo"]) <= ratio_tol * max(1.0, abs(ref["loss_ppo"]))
    assert _rel(v["dlogits"].cpu(), ref["dlogits"].reshape(-1)) <= ratio_tol
    # backward under the kernel's own ReLU decisions: element-wise
    masks = [(v["h%d" % (l + 1)][:, :H] > 0).cpu().numpy() for l in range(4)]
    own = O.ppo_policy_grads(layers, pol_states, actions, advs, logp_old, eps, entc, n_global, masks=masks)
    for l in range(5):
        assert _rel(dW[l], own["dW"][l]) <= grad_tol, ("dW", l, _rel(dW[l], own["dW"][l]))
        # a bias gradient can vanish analytically (a unit that is on for every candidate of every episode sums dlogits,
        # which is 0): its error is measured on the scale of the layer's gradient, not of itself
        scale = max(float(own["db"][l].abs().max()), float(own["dW"][l].abs().max()))
        if l < 4:           # db4 = sum of dlogits = 0 analytically: only rounding noise is left
            err = float((db[l].double() - own["db"][l]).abs().max()) / scale
            assert err <= grad_tol, ("db", l, err)
        else:
            assert float((db[l].double() - own["db"][l]).abs().max()) <= 1e-4 * float(ref["dlogits"].abs().sum()) + 1e-12
    # against the exact autograd gradient: direction
    flat = lambda gs: torch.cat([torch.as_tensor(g).double().reshape(-1) for g in gs])
    cos = torch.nn.functional.cosine_similarity(flat(dW + db[:4]), flat(ref["dW"] + ref["db"][:4]), dim=0)
    assert float(cos) > cos_min, float(cos)
    flips = [int((torch.as_tensor(masks[l]) != (ref["hs"][l] > 0)).sum()) for l in range(4)]
    return flips, float(cos)


def _random_case(seed, E, M, F_state, cols, H):
    rng = np.random.RandomState(seed)
    F = len(cols)
    dims = [F, H, H, H, H, 1]
    layers = [(rng.normal(0, 0.08, (dims[i + 1], dims[i])).astype(np.float32), rng.normal(0, 0.05, dims[i + 1]).astype(np.float32))
              for i in range(5)]
    states = rng.randn(E, M, F_state).astype(np.float32)
    actions = rng.randint(0, M, E)
    advs = rng.randn(E).astype(np.float32)
    lp0 = O.ppo_policy_grads(layers, states[:, :, cols], actions, advs, np.zeros(E), 0.15, 0.01, E)["logp"].numpy()
    logp_old = (lp0 + 0.2 * rng.randn(E)).astype(np.float32)       # some ratios outside the clip range, either side
    return layers, states, actions, advs, logp_old


@pytest.mark.parametrize("E,M,F_state,cols,H", [
    (7, 301, 6, [0, 1, 2, 3, 4, 5], 200),      # ragged row count (2107 = 16 * 128 + 59), the shipped width
    (3, 50, 6, [0, 1, 2, 3], 64),              # masked features (exclude_t / exclude_T, policies.py:109-114), narrow net
    (1, 1, 7, [0, 1, 2, 3, 4, 5, 6], 200),     # a single candidate: softmax over one logit, dlogits = 0
    (16, 966, 6, [0, 1, 2, 3, 4, 5], 200),     # Branin minibatch rows (16 x 966)
    (5, 130, 8, [0, 1, 2, 3, 4, 5, 6, 7], 200),  # all eight state columns at D = 2: layer 0 takes two K = 8 slices (K0S = 2)
])
def test_policy_grad_kernels_against_oracle(E, M, F_state, cols, H):
    layers, states, actions, advs, logp_old = _random_case(E * 1000 + M, E, M, F_state, cols, H)
    terms, v, dW, db = _run_case(E, M, F_state, cols, H, layers, states, actions, advs, logp_old)
    if M == 1:
        assert all(float(g.abs().max()) == 0.0 for g in dW[:4])      # one candidate: p = 1, every gradient vanishes
        assert float(terms[0, 0]) == 0.0 and float(terms[0, 1]) == 0.0
        return
    _check_against_oracle(E, M, cols, H, layers, states, actions, advs, logp_old, terms, v, dW, db)


def test_policy_grad_on_golden_states_with_shipped_weights(golden_dir):
    """Six states of the reference's own Branin episode under the shipped weights_1635: the forward logits of the
    update kernels against the logits the reference NeuralAF produced (golden), gradients against the oracle."""
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, "weights_bra_1635.npz"))
    layers = [(w.numpy(), b.numpy()) for w, b in pw.layers("policy_net")]
    ep = np.load(os.path.join(golden_dir, "episode_bra.npz"), allow_pickle=True)
    ts = ["t00", "t01", "t02", "t07", "t15", "t29"]
    states = np.stack([ep[t + "_state"] for t in ts]).astype(np.float32)
    gold_logits = np.stack([ep[t + "_logits"] for t in ts])
    E, M, F = states.shape
    rng = np.random.RandomState(3)
    actions = np.array([int(ep[t + "_action"]) for t in ts])
    actions[::2] = rng.randint(0, M, len(actions[::2]))
    advs = rng.randn(E).astype(np.float32)
    lp0 = O.ppo_policy_grads(layers, states, actions, advs, np.zeros(E), 0.15, 0.01, E)["logp"].numpy()
    logp_old = (lp0 + 0.1 * rng.randn(E)).astype(np.float32)
    cols = list(range(F))
    terms, v, dW, db = _run_case(E, M, F, cols, 200, layers, states, actions, advs, logp_old, n_global=2 * E)
    got = v["logits"].cpu().numpy().reshape(E, M)
    assert np.max(np.abs(got - gold_logits)) <= 2e-5 * np.max(np.abs(gold_logits))      # north-star logits bar: 1e-3
    flips, cos = _check_against_oracle(E, M, cols, 200, layers, states, actions, advs, logp_old, terms, v, dW, db,
                                       n_global=2 * E, grad_tol=8e-2, ratio_tol=4e-3, cos_min=0.998)
    # ratio_tol: |logit| up to 400, 4e-6 relative = 2e-3 absolute.  grad_tol: with the trained policy whole gradient
    # entries are ~1e-14 in exact arithmetic (sum_r dlogit_r = 0 per episode times near-constant factors); single-pass
    # TF32 rounds every term to 2^-12 and leaves ~1e-4 of the cancelled mass behind (torch's TF32 autograd does the same)
    assert sum(flips) <= 4


def test_policy_grad_argument_errors():
    from metabo_b200 import ops, _lib as L
    dev = torch.device("cuda:0")
    with pytest.raises(L.MboError):
        pg = ops.PPOPolicyGrad(2, 10, 6, list(range(6)), 300, dev)         # hidden width above 207
        z = torch.zeros(2, 10, 6, device=dev)
        W = [torch.zeros(1, device=dev)] * 5
        pg(z, W, W, torch.zeros(2, dtype=torch.int32, device=dev), torch.zeros(2, device=dev), torch.zeros(2, device=dev),
           0.1, 0.0, 2, W, W)


def test_ppo_native_update_follows_autograd_update(tmp_path):
    """One PPO iteration (rollout + 2 epochs of minibatch updates) with the native policy-gradient kernels against
    the all-torch fp32 autograd backend from the same seeds: same rollout (bit-identical policies at the start), the
    logged losses agree and the parameter update points the same way."""
    from metabo_b200 import gym_compat as gym
    from metabo_b200.policies.policies import NeuralAF
    from metabo_b200.ppo.ppo import PPO
    from autograd_ppo import AutogradPPO
    from test_dropin_gpu import BRA_SPEC
    spec = dict(BRA_SPEC, T=5, N_MS=100, N_LS=50, k=3, reward_transformation="neg_log10", env_id="MetaBO-NATIVE-v0",
                f_opts={"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50})
    gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=5, kwargs=spec)

    def run(backend, name):
        params = {"batch_size": 80, "max_steps": 80, "minibatch_size": 20, "n_epochs": 2, "lr": 1e-4, "epsilon": 0.15,
                  "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98, "loss_type": "GAElam",
                  "normalize_advs": True, "n_workers": 4, "env_id": spec["env_id"], "seed": 0, "env_seeds": [0, 1, 2, 3],
                  "policy_options": {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2,
                                     "T_idx": -1, "arch_spec_value": [200] * 4}}
        cls = AutogradPPO if backend == "autograd" else PPO
        ppo = cls(policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=params["policy_options"]), params=params,
                  logpath=str(tmp_path / name), save_interval=10, verbose=True)
        w0 = torch.cat([p.detach().reshape(-1).clone() for p in ppo.pi.parameters()])
        ppo.train()
        w1 = torch.cat([p.detach().reshape(-1) for p in ppo.pi.parameters()])
        losses = [[float(x.split("=")[1]) for x in line.split(",")] for line in ppo.iter_log_str.splitlines() if "loss_ppo" in line]
        return ppo, w0, w1, np.array(losses)

    pa, w0a, w1a, la = run("autograd", "autograd")
    pn, w0n, w1n, ln = run("native", "native")
    assert pn._pg is not None and pa._pg is None
    assert torch.equal(w0a, w0n)
    assert la.shape == ln.shape == (2, 4)
    np.testing.assert_allclose(ln, la, rtol=2e-2, atol=2e-4)
    da, dn = (w1a - w0a).double(), (w1n - w0n).double()
    cos = float(torch.nn.functional.cosine_similarity(da, dn, dim=0))
    assert cos > 0.9, cos


def test_fused_adam_matches_torch_adam():
    """mbo_adam_step against torch.optim.Adam (the reference's optimiser, ppo.py:93) over several steps and ragged
    tensor sizes; a parameter without a gradient is left alone."""
    from metabo_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    shapes = [(200, 6), (200,), (200, 200), (1, 200), (1,), (7, 3)]
    pa = [torch.randn(s, generator=g).to(dev).requires_grad_() for s in shapes]
    pb = [p.detach().clone().requires_grad_() for p in pa]
    oa = torch.optim.Adam(pa, lr=1e-2)
    ob = ops.FusedAdam(pb, lr=1e-2)
    for step in range(5):
        for i, (x, y) in enumerate(zip(pa, pb)):
            if i == 5 and step < 2:
                x.grad = y.grad = None
                continue
            gr = torch.randn(x.shape, generator=g).to(dev) * (10.0 ** (i - 3))
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    for x, y in zip(pa, pb):
        assert float((x - y).abs().max()) <= 2e-6 * float(x.abs().max()) + 1e-7
    assert int(ob.state[pb[0]]["step"].item()) == 5


def test_old_policy_logprobs_are_bitwise_the_new_policys_at_equal_weights():
    """ratio == 1 exactly while theta == theta_old: the forward-only call and the gradient call share their arithmetic."""
    from metabo_b200 import ops
    dev = torch.device("cuda:0")
    layers, states, actions, advs, _ = _random_case(5, 6, 400, 6, list(range(6)), 200)
    pg = ops.PPOPolicyGrad(8, 400, 6, list(range(6)), 200, dev)        # capacity above the call's E
    W = [torch.as_tensor(w, device=dev).contiguous() for w, _ in layers]
    b = [torch.as_tensor(x, device=dev).contiguous() for _, x in layers]
    st = torch.as_tensor(states, device=dev)
    ac = torch.as_tensor(actions, dtype=torch.int32, device=dev)
    lp, en = pg.logp_entropy(st, W, b, ac)
    dW = [torch.zeros_like(w) for w in W]
    db = [torch.zeros_like(x) for x in b]
    terms = pg(st, W, b, ac, torch.as_tensor(advs, device=dev), lp, 0.15, 0.01, 6, dW, db)
    torch.cuda.synchronize()
    pg.check()
    assert torch.equal(terms[:, 0], lp) and torch.equal(terms[:, 1], en)
    assert bool((terms[:, 3] == 1.0).all())


@pytest.mark.parametrize("E,Hv", [(60, 200), (7, 20), (119, 200)])
def test_value_net_grad_matches_autograd(E, Hv):
    """mbo_ppo_value_grad (ppo.py:150-156 value loss through the 2 -> Hv x 4 -> 1 value net on states[:, 0, [t, T]],
    policies.py:116-121) against torch-CPU fp64 autograd: values 1e-5, gradients 1e-4 of max|grad| per tensor."""
    from metabo_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(E)
    M, F = 11, 6
    dims = [2, Hv, Hv, Hv, Hv, 1]
    W = [torch.randn(dims[i + 1], dims[i], generator=g) * (0.5 if i == 0 else 0.15) for i in range(5)]
    b = [torch.randn(dims[i + 1], generator=g) * 0.1 for i in range(5)]
    states = torch.randn(E, M, F, generator=g)
    states[:, :, -2] = torch.randint(0, 30, (E, 1), generator=g).float()
    states[:, :, -1] = 30.0
    ret = torch.randn(E, generator=g)
    coeff = 1.0 / (2 * E)
    Wr = [w.double().clone().requires_grad_() for w in W]
    br = [x.double().clone().requires_grad_() for x in b]
    h = torch.stack([states[:, 0, -2], states[:, 0, -1]], 1).double()
    for l in range(4):
        h = torch.relu(h @ Wr[l].t() + br[l])
    v = (h @ Wr[4].t() + br[4]).squeeze(1)
    (coeff * ((v - ret.double()) ** 2).sum()).backward()
    vg = ops.PPOValueGrad(E + 3, Hv, dev)
    Wd, bd = [w.to(dev).contiguous() for w in W], [x.to(dev).contiguous() for x in b]
    dW, db = [torch.full_like(w, 3.0) for w in Wd], [torch.full_like(x, 3.0) for x in bd]
    vals, vterms = vg(states.to(dev), -2, -1, Wd, bd, ret.to(dev), coeff, dW, db)
    torch.cuda.synchronize()
    assert _rel(vals, v.detach()) <= 1e-5
    assert _rel(vterms, ((v - ret.double()) ** 2).detach()) <= 1e-4
    for l in range(5):
        assert _rel(dW[l], Wr[l].grad) <= 1e-4, ("dW", l)
        assert _rel(db[l], br[l].grad) <= 1e-4, ("db", l)


def test_gemm_kernel_variants_agree_bitwise(monkeypatch):
    """The persistent double-buffered dgrad kernel (default) and the one-tile-per-CTA kernel, the TS-form forward (A operand
    in tensor memory, default) and the SS-form forward (A split in shared memory) issue the same MMA sequence per output
    tile: every activation, dZ and gradient is bitwise identical whichever variants run, on a ragged row count that gives
    the persistent CTAs partial and empty tiles."""
    layers, states, actions, advs, logp_old = _random_case(11, 9, 517, 6, list(range(6)), 200)
    monkeypatch.setenv("MBO_UPD_FWD_K2", "0")               # the GEMM-per-layer forward (the default is the fused tcgen05 policy kernel)
    outs = []
    for dgr, ts, dts in (("1", "1", "0"), ("0", "0", "0"), ("1", "0", "1"), ("0", "1", "1")):
        monkeypatch.setenv("MBO_UPD_PERSIST_DGRAD", dgr)
        monkeypatch.setenv("MBO_UPD_FWD_TS", ts)            # forward: A operand in TMEM (TS form) or in shared memory
        monkeypatch.setenv("MBO_UPD_DGRAD_TS", dts)         # dgrad: TS form instead of the persistent / one-tile SS kernels
        terms, v, dW, db = _run_case(9, 517, 6, list(range(6)), 200, layers, states, actions, advs, logp_old)
        outs.append((terms, v, dW, db))
    t0, v0, dW0, db0 = outs[0]
    for terms, v, dW, db in outs[1:]:
        assert torch.equal(terms, t0)
        for k in ("h1", "h2", "h3", "h4", "dz2", "dz3", "dz4", "logits", "dlogits"):
            assert torch.equal(v[k], v0[k]), k
        # the SS forward has no ReLU words of h1: its last dgrad stores dH1 and the layer-0 wgrad applies the mask
        assert torch.equal(v["dz1"] * (v["h1"] > 0), v0["dz1"] * (v0["h1"] > 0))
        assert all(torch.equal(a, b) for a, b in zip(dW + db, dW0 + db0))


def test_prepare_minibatch_and_loss_sums_match_torch():
    """mbo_ppo_prepare_minibatch (gather + ppo.py:141-144 advantage normalisation) and mbo_ppo_loss_sums (ppo.py:158-168)
    against the torch expressions of the drop-in layer."""
    from metabo_b200 import ops
    from metabo_b200.ppo.ppo import normalize_advantages
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(7)
    n, M, F, E = 50, 13, 6, 17
    flat = dict(states=torch.randn(n, M, F, generator=g).to(dev), actions=torch.randint(0, M, (n,), generator=g).to(dev).to(torch.int32),
                advs=torch.randn(n, generator=g).to(dev) * 3 + 1, tdlamrets=torch.randn(n, generator=g).to(dev),
                logprobs_old=torch.randn(n, generator=g).to(dev))
    sel = torch.randperm(n, generator=g)[:E].to(dev)
    out = dict(states=torch.zeros(E, M, F, device=dev), actions=torch.zeros(E, dtype=torch.int32, device=dev),
               advs=torch.zeros(E, device=dev), tdlamrets=torch.zeros(E, device=dev), logprobs_old=torch.zeros(E, device=dev))
    for normalize in (True, False):
        ops.ppo_prepare_minibatch(sel, flat, normalize, out)
        torch.cuda.synchronize()
        for k in ("states", "actions", "tdlamrets", "logprobs_old"):
            assert torch.equal(out[k], flat[k].index_select(0, sel)), k
        ref = flat["advs"].index_select(0, sel)
        ref = normalize_advantages(ref) if normalize else ref
        assert float((out["advs"] - ref).abs().max()) <= 2e-6 * float(ref.abs().max())
    flat["advs"][:] = 2.5                                            # std == 0: advantages are left alone
    ops.ppo_prepare_minibatch(sel, flat, True, out)
    assert bool((out["advs"] == 2.5).all())
    terms = torch.randn(E, 4, generator=g).to(dev)
    vterms = torch.rand(E, generator=g).to(dev)
    out4 = torch.zeros(4, device=dev)
    ops.ppo_loss_sums(terms, vterms, 2 * E, 0.7, 0.01, out4)
    l_ppo, l_val, l_ent = terms[:, 2].sum(), vterms.sum() / (2 * E), -terms[:, 1].sum() / (2 * E)
    ref4 = torch.stack([l_ppo, l_val, l_ent, l_ppo + 0.7 * l_val + 0.01 * l_ent])
    assert float((out4 - ref4).abs().max()) <= 1e-5 * float(ref4.abs().max())


def test_fused_policy_forward_agrees_with_gemm_forward(monkeypatch):
    """Default forward of the update = ONE launch of the tcgen05 policy kernel (bf16x3, activations resident in tensor
    memory, h1..h4 + ReLU bit words written for the backward) vs the GEMM-per-layer 3xTF32 forward (MBO_UPD_FWD_K2=0):
    both are fp32-class -- logits within 2e-5 of max|logit|, stored activations within 2e-5 of max|h|, ReLU decisions
    identical except at pre-activations within rounding of 0, gradients pointing the same way (cosine > 0.9999) and the
    ones column / zero padding of every stored activation in place."""
    layers, states, actions, advs, logp_old = _random_case(21, 7, 966, 6, list(range(6)), 200)
    outs = {}
    for k2 in ("0", "1"):
        monkeypatch.setenv("MBO_UPD_FWD_K2", k2)
        outs[k2] = _run_case(7, 966, 6, list(range(6)), 200, layers, states, actions, advs, logp_old)
    (t0, v0, dW0, db0), (t1, v1, dW1, db1) = outs["0"], outs["1"]
    R = 7 * 966
    scale = float(v0["logits"][:R].abs().max())
    assert float((v0["logits"][:R] - v1["logits"][:R]).abs().max()) <= 2e-5 * scale
    for k in ("h1", "h2", "h3", "h4"):
        a, b = v0[k][:R], v1[k][:R]
        assert float((a - b).abs().max()) <= 2e-5 * float(a.abs().max()), k
        assert bool((b[:, 200] == 1.0).all()) and float(b[:, 201:].abs().max()) == 0.0, k
        assert float(((a > 0) != (b > 0)).float().mean()) < 1e-4, k
    np.testing.assert_allclose(t1[:, 0].cpu().numpy(), t0[:, 0].cpu().numpy(), atol=2e-5 * scale)       # log-probs
    g0 = torch.cat([x.reshape(-1) for x in dW0 + db0]).double()
    g1 = torch.cat([x.reshape(-1) for x in dW1 + db1]).double()
    assert float(torch.nn.functional.cosine_similarity(g0, g1, dim=0)) > 0.9999
