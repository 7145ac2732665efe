"""The oracle restatement (oracle/metabo_oracle.py) must reproduce the vectors minted from the
reference's own code (oracle/make_golden.py -> tests/golden/episode_*.npz)."""
import json
import os

import numpy as np
import pytest

from oracle import metabo_oracle as O

CASES = [("bra", "weights_bra_1635.npz"), ("hm3", "weights_hm3_1988.npz"), ("gps", "weights_gps_1161.npz")]


def load_case(golden_dir, name, wfile):
    ep = np.load(os.path.join(golden_dir, "episode_%s.npz" % name))
    spec = json.loads(str(ep["__env_spec__"]))
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, wfile))
    if spec["local_af_opt"]:
        grids = O.Grids(spec["D"], N_MS=spec["N_MS"], N_LS=spec["N_LS"], k=spec["k"])
    else:
        grids = O.Grids(spec["D"], cardinality_domain=spec["cardinality_domain"])
    return ep, spec, pw, grids


def make_gp(ep, spec, X, Y):
    gp = O.EpisodeGP(spec["D"], spec.get("kernel", "RBF"), float(ep["kernel_variance"]), ep["kernel_lengthscale"],
                     float(ep["noise_variance"]), spec["use_prior_mean_function"], float(ep["y_min"]))
    gp.set_data(X, Y)
    return gp


@pytest.mark.parametrize("name,wfile", CASES)
def test_af_round_reproduces_reference(golden_dir, name, wfile):
    """Teacher-forced, stage by stage: GP mean/std and states are bitwise (same numpy fp64
    code path as the stand-in GPy); logits carry 1-ulp fp32 noise from torch's thread-count
    dependent summation order, which can permute near-tied top-k entries (SURVEY section 0
    finding 2), so selections are compared by value, not by index."""
    ep, spec, pw, grids = load_case(golden_dir, name, wfile)
    feats, T = spec["features"], spec["T"]
    tol = dict(rtol=2e-6, atol=1e-5)
    for t in ep["keep"]:
        t = int(t)
        p = "t%02d_" % t
        gp = make_gp(ep, spec, ep[p + "X"], ep[p + "Y"])
        if spec["local_af_opt"]:
            mean, std = gp.eval_gp(grids.multistart_grid)
            assert np.array_equal(mean, ep[p + "mean_ms"]) and np.array_equal(std, ep[p + "std_ms"])
            s_ms = O.get_state(gp, grids.multistart_grid, feats, t, T)
            af_ms = O.policy_forward(pw, s_ms[None])[0][0].numpy()
            np.testing.assert_allclose(af_ms, ep[p + "af_ms"], **tol)
            # the reference's local grids are k affine images of the Sobol table around its top-k starts
            ref_local = ep[p + "local"]
            top = np.argsort(-ep[p + "af_ms"])[:grids.k]
            assert np.array_equal(grids.local_grids(grids.multistart_grid[top]), ref_local)
            s_ls = O.get_state(gp, ref_local, feats, t, T)
            af_ls = O.policy_forward(pw, s_ls[None])[0][0].numpy()
            np.testing.assert_allclose(af_ls, ep[p + "af_ls"], **tol)
            top_ls = np.argsort(-ep[p + "af_ls"])[:grids.k]
            assert np.array_equal(np.concatenate([ref_local[top_ls], grids.multistart_grid]), ep[p + "xi_t"])
        xi_t = ep[p + "xi_t"]
        state = O.get_state(gp, xi_t, feats, t, T)
        assert np.array_equal(state, ep[p + "state"])
        logits, values = O.policy_forward(pw, state[None])
        np.testing.assert_allclose(logits[0].numpy(), ep[p + "logits"], **tol)
        assert abs(float(values[0]) - float(ep[p + "value"])) < 1e-5
        a_ref = int(ep[p + "action"])
        lg = logits[0].numpy()
        assert lg[a_ref] >= lg.max() - 2e-6 * abs(lg.max()) - 1e-5
        # the free-running oracle round picks points whose AF value ties the reference's picks
        r = O.af_round(gp, grids, pw, feats, t, T)
        assert r["logits"].max() >= ep[p + "logits"].max() - 2e-6 * abs(ep[p + "logits"].max()) - 1e-5


def test_objectives_and_rewards(golden_dir):
    # known-answer: regret -> ~0 at the analytic optimum (objectives.py:58-68, 199-206)
    ymax, pos = O.bra_max(np.zeros(2), 1.0)
    assert abs(O.bra_var(pos, np.zeros(2), 1.0)[0, 0] - ymax) < 1e-12
    g = O.create_uniform_grid(np.array([[0, 1.0], [0, 1.0]]), 200)
    assert O.bra_var(g, np.zeros(2), 1.0).max() <= ymax + 1e-3
    ymax3, pos3 = O.hm3_max(np.zeros(3), 1.0)
    g3 = O.create_uniform_grid(np.array([[0, 1.0]] * 3), 40)
    assert O.hm3_var(g3, np.zeros(3), 1.0).max() <= ymax3 + 1e-6


def test_gae_matches_closed_form():
    r = np.array([1.0, 2.0, 3.0])
    v = np.array([0.5, 0.4, 0.3])
    adv, ret = O.gae(r, v, [1, 0, 0], next_new=1, next_value=0.3, gamma=0.9, lam=0.8)
    d2 = -0.3 + 3.0
    d1 = -0.4 + 2.0 + 0.9 * 0.3
    d0 = -0.5 + 1.0 + 0.9 * 0.4
    a1 = d1 + 0.72 * d2
    a0 = d0 + 0.72 * a1
    np.testing.assert_allclose(adv, [a0, a1, d2])
    np.testing.assert_allclose(ret, adv + v)
