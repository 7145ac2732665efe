"""ORACLE (test infrastructure): mint the golden fixtures under tests/golden/ by running the
REFERENCE'S OWN code (metabo.environment.metabo_gym.MetaBO + metabo.policies.policies.NeuralAF,
imported unmodified through oracle/ref_loader.py) on the shipped iclr2020 weights.

Run in the authoring container only (needs /root/reference):

    python -m oracle.make_golden

Writes
  tests/golden/weights_{bra_1635,hm3_1988,gps_1161}.npz   shipped state_dicts + policy_options
  tests/golden/episode_{bra,hm3,gps}.npz                  per-step traces of one deterministic
                                                          episode (seed 100), reference actions

Every array in episode_*.npz is produced by reference code; the only non-reference numerics
underneath are the GPy stand-in (oracle/gp_oracle.py::GPyRoute, "parity unpinned", see its
header) and the sobol_seq stand-in (oracle/sobol.py).
"""
import json
import os
import pickle as pkl

import numpy as np
import torch

from .ref_loader import load_reference, shipped

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CONFIGS = {
    # evaluate_metabo_branin.py:91-111
    "bra": dict(
        env_spec={
            "env_id": "MetaBO-BRA-v0", "D": 2, "f_type": "BRA-var",
            "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1},
            "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"],
            "T": 30, "n_init_samples": 0, "pass_X_to_pi": False,
            "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0, "noise_variance": 8.9e-16,
            "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 1000, "N_LS": 1000, "k": 5,
            "reward_transformation": "none"},
        logdir=("gobfcts", "bra", "M_50", "MetaBO-BRA-v0"), load_iter=1635, keep=[0, 1, 2, 7, 15, 29]),
    # evaluate_metabo_hm3.py:92-110
    "hm3": dict(
        env_spec={
            "env_id": "MetaBO-HM3-v0", "D": 3, "f_type": "HM3-var",
            "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1},
            "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"],
            "T": 30, "n_init_samples": 0, "pass_X_to_pi": False,
            "kernel_lengthscale": [0.716, 0.298, 0.186], "kernel_variance": 0.83, "noise_variance": 1.688e-11,
            "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 2000, "N_LS": 2000, "k": 5,
            "reward_transformation": "none"},
        logdir=("gobfcts", "hm3", "M_50", "MetaBO-HM3-v0"), load_iter=1988, keep=[0, 1, 4, 12, 29]),
    # train_metabo_gps.py:34-60 (evaluate_metabo_gps.py uses the same family)
    "gps": dict(
        env_spec={
            "env_id": "MetaBO-GP-v0", "D": 3, "f_type": "GP",
            "f_opts": {"kernel": "Matern52", "lengthscale_low": 0.05, "lengthscale_high": 0.5,
                       "noise_var_low": 0.1, "noise_var_high": 0.1, "signal_var_low": 1.0,
                       "signal_var_high": 1.0, "min_regret": 1e-6},
            "features": ["posterior_mean", "posterior_std", "incumbent", "timestep_perc", "timestep", "budget"],
            "T": 30, "n_init_samples": 0, "pass_X_to_pi": False, "kernel": "Matern52",
            "kernel_lengthscale": None, "kernel_variance": None, "noise_variance": None,
            "use_prior_mean_function": True, "local_af_opt": False, "cardinality_domain": 2000,
            "reward_transformation": "neg_log10"},
        logdir=("gps", "MetaBO-GP-v0", "Matern52"), load_iter=1161, keep=[0, 1, 5, 14, 29]),
}


def export_weights(name, cfg, n_features):
    logdir = shipped(*cfg["logdir"])
    with open(os.path.join(logdir, "params_%d" % cfg["load_iter"]), "rb") as f:
        params = pkl.load(f)
    sd = torch.load(os.path.join(logdir, "weights_%d" % cfg["load_iter"]), weights_only=False, map_location="cpu")
    opts = {k: v for k, v in params["policy_options"].items()
            if k in ("activations", "arch_spec", "arch_spec_value", "use_value_network", "t_idx", "T_idx",
                     "exclude_t_from_policy", "exclude_T_from_policy")}
    arrs = {k: v.detach().cpu().numpy().astype(np.float32) for k, v in sd.items()}
    arrs["__policy_options__"] = np.array(json.dumps(opts))
    arrs["__n_features__"] = np.array(n_features)
    out = os.path.join(GOLDEN, "weights_%s_%d.npz" % (name, cfg["load_iter"]))
    np.savez_compressed(out, **arrs)
    return params["policy_options"], sd, out


def trace_episode(name, cfg, seed=100):
    metabo = load_reference()
    from metabo.environment.metabo_gym import MetaBO
    from metabo.policies.policies import NeuralAF
    torch.set_num_threads(1)
    env = MetaBO(**cfg["env_spec"])
    opts, sd, _ = export_weights(name, cfg, env.n_features)
    pi = NeuralAF(env.observation_space, env.action_space, deterministic=True, options=opts)
    pi.load_state_dict(sd)
    pi.set_requires_grad(False)
    env.set_af_functions(pi.af)
    env.seed(seed)

    calls = []
    orig_get_state, orig_af = env.get_state, env.af

    def get_state(X):
        s = orig_get_state(X)
        mean, std = env.eval_gp(X)
        calls.append(dict(kind="state", X=np.array(X), state=s, mean=mean, std=std))
        return s

    def af(state):
        a = orig_af(state)
        calls.append(dict(kind="af", af=np.array(a)))
        return a

    env.get_state, env.af = get_state, af
    rec = {}
    state = env.reset()
    T = env.T
    rewards = []
    for t in range(T + 1):
        # `calls` now holds the AF-optimisation round that produced `state`
        with torch.no_grad():
            action, value = pi.act(torch.from_numpy(state.astype(np.float32)))
            logits, _ = pi.forward(torch.from_numpy(state.astype(np.float32))[None])
        if t in cfg["keep"]:
            p = "t%02d_" % t
            rec[p + "X"] = np.zeros((0, env.D)) if env.X is None else env.X.copy()
            rec[p + "Y"] = np.zeros((0,)) if env.Y is None else env.Y[:, 0].copy()
            rec[p + "xi_t"] = np.array(env.xi_t)
            rec[p + "state"] = state.copy()
            rec[p + "logits"] = logits[0].numpy().copy()
            rec[p + "action"] = np.array(int(action))
            rec[p + "value"] = np.array(float(value))
            rec[p + "incumbent"] = np.array(env.get_incumbent())
            st = [c for c in calls if c["kind"] == "state"]
            afs = [c for c in calls if c["kind"] == "af"]
            rec[p + "mean_final"] = st[-1]["mean"]
            rec[p + "std_final"] = st[-1]["std"]
            if env.do_local_af_opt:
                rec[p + "mean_ms"] = st[0]["mean"]
                rec[p + "std_ms"] = st[0]["std"]
                rec[p + "af_ms"] = afs[0]["af"]
                rec[p + "local"] = st[1]["X"]
                rec[p + "mean_ls"] = st[1]["mean"]
                rec[p + "std_ls"] = st[1]["std"]
                rec[p + "af_ls"] = afs[1]["af"]
        if t == T:
            break
        calls.clear()
        state, reward, done, _ = env.step(action.numpy())
        rewards.append(reward)
    rec["rewards"] = np.array(rewards)
    rec["actions_all_X"] = env.X.copy()
    rec["Y_all"] = env.Y[:, 0].copy()
    rec["y_max"] = np.array(float(np.asarray(env.y_max).reshape(-1)[0]))
    rec["y_min"] = np.array(float(np.asarray(env.y_min).reshape(-1)[0]))
    rec["kernel_lengthscale"] = np.broadcast_to(np.asarray(env.kernel_lengthscale, dtype=np.float64), (env.D,)).copy()
    rec["kernel_variance"] = np.array(float(env.kernel_variance))
    rec["noise_variance"] = np.array(float(env.noise_variance))
    rec["keep"] = np.array(cfg["keep"])
    rec["__env_spec__"] = np.array(json.dumps(cfg["env_spec"]))
    out = os.path.join(GOLDEN, "episode_%s.npz" % name)
    np.savez_compressed(out, **rec)
    return out, rewards


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for name, cfg in CONFIGS.items():
        out, rewards = trace_episode(name, cfg)
        print(name, "->", out, "%.1f KB" % (os.path.getsize(out) / 1024), "final reward", rewards[-1])


if __name__ == "__main__":
    main()
