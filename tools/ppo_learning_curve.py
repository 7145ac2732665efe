"""train_metabo_branin.py configuration (C3) for N iterations with the library's update kernels (metabo_b200 PPO) and with the
all-torch comparison implementation (tests/autograd_ppo.py) from the same seeds: the reward curves (avg_step_rews per
iteration) side by side.  python tools/ppo_learning_curve.py [iterations]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from metabo_b200 import gym_compat as gym
from metabo_b200.policies.policies import NeuralAF
from metabo_b200.ppo.ppo import PPO
from autograd_ppo import AutogradPPO
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 30
spec = {"env_id": "MetaBO-BRA-v0", "D": 2, "f_type": "BRA-var", "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50},
        "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": 30, "n_init_samples": 0,
        "pass_X_to_pi": False, "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0, "noise_variance": 8.9e-16,
        "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 1000, "N_LS": 1000, "k": 5,
        "reward_transformation": "neg_log10"}
gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=30, kwargs=spec)
out = {}
for backend in ("native", "autograd"):
    params = {"batch_size": 1200, "max_steps": n_it * 1200, "minibatch_size": 60, "n_epochs": 4, "lr": 1e-4, "epsilon": 0.15,
              "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98, "loss_type": "GAElam", "normalize_advs": True,
              "n_workers": 10, "env_id": spec["env_id"], "seed": 0, "env_seeds": list(range(10)),
              "policy_options": {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2, "T_idx": -1,
                                 "arch_spec_value": [200] * 4}}
    ppo = (PPO if backend == "native" else AutogradPPO)(policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=params["policy_options"]), params=params,
              logpath=os.path.join(tempfile.mkdtemp(), "log"), save_interval=1000, verbose=True)
    t0 = time.time(); ppo.train(); wall = time.time() - t0
    out[backend] = dict(wall_s=wall, avg_step_rews=[round(float(x), 4) for x in ppo.stats["avg_step_rews"]],
                        avg_term_rews=[round(float(x), 4) for x in ppo.stats["avg_term_rews"]])
a, b = np.array(out["native"]["avg_step_rews"]), np.array(out["autograd"]["avg_step_rews"])
out["max_abs_diff_avg_step_rew"] = float(np.abs(a - b).max())
out["mean_last5"] = {"native": float(a[-5:].mean()), "autograd": float(b[-5:].mean()), "first": float(a[0])}
print(json.dumps(out))
