"""Mints tests/golden/baseline_af.npz: outputs of the reference's OWN EI / UCB / PI classes
(metabo/policies/policies.py:181-307, imported from /root/reference through oracle/ref_loader.py) on seeded float32
states with the evaluation feature order (metabo_gym.py:95).  Test infrastructure; run in the authoring container
(the reference is not available on the GPU box): python oracle/make_golden_baselines.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

FEATURE_ORDER = ["posterior_mean", "posterior_std", "incumbent", "timestep"]


def main():
    load_reference()
    from metabo.policies.policies import EI, PI, UCB
    rng = np.random.RandomState(11)
    M = 700
    rec = {}
    for case, (t, inc) in enumerate([(0, 0.0), (7, 0.35), (29, 1.9)]):
        state = np.zeros((M, 4), dtype=np.float32)
        state[:, 0] = rng.randn(M) * 0.8 + 0.2
        state[:, 1] = np.abs(rng.randn(M)) * 0.5 + 1e-4
        state[:8, 1] = 1e-7                                  # nearly observed points: |z| up to 1e7
        state[:, 2] = inc
        state[:, 3] = t
        p = "c%d_" % case
        rec[p + "state"] = state
        rec[p + "ei"] = EI(feature_order=FEATURE_ORDER).af(state)
        rec[p + "pi"] = PI(feature_order=FEATURE_ORDER, xi=0.01).af(state)
        rec[p + "ucb_fixed"] = UCB(feature_order=FEATURE_ORDER, kappa=2.0).af(state)
        rec[p + "ucb_gp"] = UCB(feature_order=FEATURE_ORDER, kappa="gp_ucb", D=2, delta=0.1).af(state)
    rec["feature_order"] = np.array(FEATURE_ORDER)
    out = os.path.join(ROOT, "tests", "golden", "baseline_af.npz")
    np.savez_compressed(out, **rec)
    print(out, "%.1f KB" % (os.path.getsize(out) / 1024), {k: (v.dtype, v.shape) for k, v in rec.items() if k.startswith("c0_")})


if __name__ == "__main__":
    main()
