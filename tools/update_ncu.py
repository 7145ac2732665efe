"""Two PPO iterations of the C3 configuration (train_metabo_branin.py: batch 1200, minibatch 60 x 966 candidates,
80 optimiser steps per iteration) with the optimiser step NOT graph-captured, for an ncu launch list of the update:
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv ... python tools/update_ncu.py [high|highest]"""
import os, sys, time, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabo_b200 import gym_compat as gym
from metabo_b200.policies.policies import NeuralAF
from metabo_b200.ppo.ppo import PPO
prec = sys.argv[1] if len(sys.argv) > 1 else "highest"
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
spec = {"env_id": "MetaBO-BRA-v0", "D": 2, "f_type": "BRA-var", "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50},
        "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": 30, "n_init_samples": 0,
        "pass_X_to_pi": False, "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0, "noise_variance": 8.9e-16,
        "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 1000, "N_LS": 1000, "k": 5,
        "reward_transformation": "neg_log10"}
gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=30, kwargs=spec)
params = {"batch_size": 1200, "max_steps": 2 * 1200, "minibatch_size": 60, "n_epochs": 4, "lr": 1e-4, "epsilon": 0.15,
          "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98, "loss_type": "GAElam", "normalize_advs": True,
          "n_workers": 10, "env_id": spec["env_id"], "seed": 0, "env_seeds": list(range(10)), "mlp_mode": "fp16",
          "update_matmul_precision": prec, "update_graph": graph,
          "policy_options": {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2, "T_idx": -1,
                             "arch_spec_value": [200] * 4, "mlp_mode": "fp16"}}
ppo = PPO(policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=params["policy_options"]), params=params,
          logpath=os.path.join(tempfile.mkdtemp(), "log"), save_interval=10, verbose=False)
ppo.train()
print(json.dumps({"prec": prec, "graph": graph, "t_optim_s": ppo.stats["t_optim"], "t_batch_s": ppo.stats["batch_stats"]["t_batch"]}))
