"""Closed-form baseline AFs: oracle vs the reference's own classes (golden vectors minted by
oracle/make_golden_baselines.py), CUDA kernel vs both, and an end-to-end EI episode batch."""
import os

import numpy as np
import pytest
import torch

from oracle import metabo_oracle as O

KINDS = ["ei", "pi", "ucb_fixed", "ucb_gp"]


def _oracle(kind, st):
    if kind == "ei":
        return O.ei_af(st[:, 0], st[:, 1], st[:, 2])
    if kind == "pi":
        return O.pi_af(st[:, 0], st[:, 1], st[:, 2], 0.01)
    if kind == "ucb_fixed":
        return O.ucb_af(st[:, 0], st[:, 1], 2.0)
    return O.ucb_af(st[:, 0], st[:, 1], "gp_ucb", timestep=st[:, 3], D=2, delta=0.1)


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_reference_classes(golden_dir, kind):
    g = np.load(os.path.join(golden_dir, "baseline_af.npz"))
    for c in range(3):
        st, ref = g["c%d_state" % c], g["c%d_%s" % (c, kind)]
        got = _oracle(kind, st)
        # UCB is float32 arithmetic in the reference, the restatement widens first: 1 float32 ulp
        tol = 2e-7 if kind.startswith("ucb") else 1e-12
        assert np.max(np.abs(got - ref)) <= tol * max(1.0, np.max(np.abs(ref))), (kind, c)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_closed_form_af_kernel_matches_reference(golden_dir, kind):
    from metabo_b200 import ops, _lib as L
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "baseline_af.npz"))
    sts = [g["c%d_state" % c] for c in range(3)]
    mean = torch.as_tensor(np.stack([s[:, 0] for s in sts]), device=dev).contiguous()
    std = torch.as_tensor(np.stack([s[:, 1] for s in sts]), device=dev).contiguous()
    epi = torch.zeros((3, 4), dtype=torch.float32, device=dev)
    epi[:, 0] = torch.as_tensor([float(s[0, 2]) for s in sts])
    epi[:, 2] = torch.as_tensor([float(s[0, 3]) for s in sts])
    args = {"ei": (L.AF_EI, 0.0, 0.0, 0), "pi": (L.AF_PI, 0.01, 0.0, 0), "ucb_fixed": (L.AF_UCB, 2.0, 0.0, 0),
            "ucb_gp": (L.AF_UCB, 0.0, 0.1, 2)}[kind]
    af = ops.af_closed_form(args[0], mean, std, epi, p0=args[1], p1=args[2], D=args[3]).cpu().numpy()
    for c in range(3):
        ref = g["c%d_%s" % (c, kind)]
        assert np.max(np.abs(af[c] - ref)) <= 2e-7 * max(1.0, np.max(np.abs(ref))), (kind, c)      # fp32 output
        assert int(np.argmax(af[c])) in np.flatnonzero(ref >= ref.max() - 2e-7 * max(1.0, abs(ref.max())))


@pytest.mark.gpu
@pytest.mark.parametrize("policy", ["EI", "GP-UCB", "PI"])
def test_baseline_af_episodes_on_the_gpu_engine(policy):
    """A batch of Branin episodes driven by a closed-form AF inside the AF-optimisation rounds (multistart -> local
    grids -> xi_t), with one Sobol initial-design point: at every step the action the engine takes is the argmax of
    the oracle's AF on the state the engine reports (unless near-tied), and the regret shrinks."""
    from metabo_b200 import _lib as L
    from metabo_b200.environment.vec_env import VecMetaBO
    from metabo_b200.policies.policies import EI, PI, UCB
    fo = ["posterior_mean", "posterior_std", "incumbent", "timestep"]
    spec = {"D": 2, "f_type": "BRA-var", "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1},
            "features": fo, "T": 8, "n_init_samples": 1, "pass_X_to_pi": False, "kernel": "RBF",
            "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0, "noise_variance": 8.9e-16,
            "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 400, "N_LS": 300, "k": 3,
            "reward_transformation": "none"}
    B = 6
    env = VecMetaBO(B, list(range(50, 50 + B)), **spec)
    pi = {"EI": EI(fo), "PI": PI(fo, xi=0.01), "GP-UCB": UCB(fo, kappa="gp_ucb", D=2, delta=0.1)}[policy]
    env.set_policy(pi)
    env.reset(deterministic=True)
    first = last = None
    for t in range(spec["T"]):
        st = env.current_state().cpu().numpy()                    # [B, M, 4] float32
        act = env.cur["action"].cpu().numpy()
        for b in range(B):
            s = st[b]
            ref = {"EI": lambda: O.ei_af(s[:, 0], s[:, 1], s[:, 2]),
                   "PI": lambda: O.pi_af(s[:, 0], s[:, 1], s[:, 2], 0.01),
                   "GP-UCB": lambda: O.ucb_af(s[:, 0], s[:, 1], "gp_ucb", timestep=s[:, 3], D=2, delta=0.1)}[policy]()
            tol = 1e-6 * max(1.0, float(np.max(np.abs(ref))))
            assert ref[act[b]] >= ref.max() - tol, (policy, t, b)
        reward, done = env.step(deterministic=True)
        r = reward.cpu().numpy()
        first = r if first is None else first
        last = r
    assert done
    assert np.all(last <= first + 1e-12) and np.median(last) < np.median(first)      # simple regret never grows
    # the first point of every lane is the Sobol initial design, whatever the AF proposed
    x0 = env.engine.X[:, 0].cpu().numpy()
    assert np.allclose(x0, x0[0])
