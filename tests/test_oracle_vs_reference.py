"""Authoring-container-only check: the oracle's policy restatement equals the reference's own
NeuralAF (policies.py:104-161) on the shipped weights, bit for bit (same torch CPU kernels)."""
import os

import numpy as np
import pytest
import torch

from oracle import metabo_oracle as O
from oracle.ref_loader import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("name,wfile,F,M", [("bra", "weights_bra_1635.npz", 6, 966),
                                            ("gps", "weights_gps_1161.npz", 6, 2000)])
def test_policy_forward_equals_reference(golden_dir, name, wfile, F, M):
    from oracle.ref_loader import load_reference
    load_reference()
    from metabo.policies.policies import NeuralAF
    import gym
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, wfile))
    obs = gym.spaces.Box(-1e5, 1e5, shape=(M, F), dtype=np.float32)
    pi = NeuralAF(obs, gym.spaces.Discrete(M), deterministic=True, options=pw.options)
    pi.load_state_dict({k: v.clone() for k, v in pw.sd.items()})
    rng = np.random.RandomState(0)
    states = rng.randn(3, M, F).astype(np.float32)
    states[:, :, -2] = 7.0
    states[:, :, -1] = 30.0
    with torch.no_grad():
        lr, vr = pi.forward(torch.from_numpy(states))
        acts = torch.tensor([1.0, 5.0, 9.0])
        v2, lp, en = pi.predict_vals_logps_ents(torch.from_numpy(states), acts)
    lo, vo = O.policy_forward(pw, states)
    assert torch.equal(lr, lo) and torch.equal(vr, vo)
    lpo, eno = O.categorical_logp_entropy(lo, acts.numpy())
    assert torch.allclose(lp, lpo, atol=0, rtol=0) and torch.allclose(en, eno, rtol=1e-6)


def test_policy_gradient_restatement_equals_reference_autograd(golden_dir):
    """oracle.ppo_policy_grads (the checker of the PPO update kernels) against the reference's own
    NeuralAF.predict_vals_logps_ents + the loss lines of ppo.py:148-168 under torch autograd, shipped Branin weights,
    golden states: same gradient for every policy-net tensor (fp32 autograd vs the fp64 restatement)."""
    from oracle.ref_loader import load_reference
    load_reference()
    from metabo.policies.policies import NeuralAF
    import gym
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, "weights_bra_1635.npz"))
    ep = np.load(os.path.join(golden_dir, "episode_bra.npz"), allow_pickle=True)
    states = np.stack([ep[t + "_state"] for t in ("t01", "t07", "t15", "t29")]).astype(np.float32)
    E, M, F = states.shape
    obs = gym.spaces.Box(-1e5, 1e5, shape=(M, F), dtype=np.float32)
    pi = NeuralAF(obs, gym.spaces.Discrete(M), deterministic=False, options=pw.options)
    pi.load_state_dict({k: v.clone() for k, v in pw.sd.items()})
    rng = np.random.RandomState(1)
    actions = rng.randint(0, M, E)
    advs = rng.randn(E).astype(np.float32)
    eps, entc = 0.15, 0.01
    st, ac = torch.from_numpy(states), torch.from_numpy(actions.astype(np.float32))
    with torch.no_grad():
        _, lp0, _ = pi.predict_vals_logps_ents(states=st, actions=ac)
    logp_old = (lp0.numpy() + 0.2 * rng.randn(E)).astype(np.float32)        # ratios on both sides of the clip range
    # ppo.py:148-168, policy terms only (the value loss touches the value net alone)
    _, logprobs, entropies = pi.predict_vals_logps_ents(states=st, actions=ac)
    ratios = torch.exp(logprobs - torch.from_numpy(logp_old))
    clipped = ratios.clamp(1 - eps, 1 + eps)
    a = torch.from_numpy(advs)
    loss = -torch.mean(torch.min(ratios * a, clipped * a)) + entc * (-torch.mean(entropies))
    pi.zero_grad()
    loss.backward()
    layers = [(w.numpy(), b.numpy()) for w, b in pw.layers("policy_net")]
    ref = O.ppo_policy_grads(layers, states, actions, advs, logp_old, eps, entc, E)
    assert ((ratios.detach().numpy() < 1 - eps) | (ratios.detach().numpy() > 1 + eps)).any()
    for l, fc in enumerate(pi.policy_net.fc):
        gw, gb = fc.weight.grad.double(), fc.bias.grad.double()
        assert float((gw - ref["dW"][l]).abs().max()) <= 2e-4 * float(ref["dW"][l].abs().max()) + 1e-12, l
        if l < 4:
            assert float((gb - ref["db"][l]).abs().max()) <= 2e-4 * float(ref["db"][l].abs().max()) + 1e-12, l
    assert abs(float(logprobs.detach()[0]) - float(ref["logp"][0])) < 1e-4
