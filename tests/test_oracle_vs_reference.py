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
