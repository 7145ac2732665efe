"""Kernel-time table of one PPO optimize_on_batch (C3: 80 optimiser steps) from torch.profiler, eager (no graph) so
that every kernel is listed: python tools/update_profile.py [native|autograd]"""
import os, sys, json, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from metabo_b200 import gym_compat as gym
from metabo_b200.policies.policies import NeuralAF
from metabo_b200.ppo.ppo import PPO
backend = sys.argv[1] if len(sys.argv) > 1 else "native"
spec = {"env_id": "MetaBO-BRA-v0", "D": 2, "f_type": "BRA-var", "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50},
        "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": 30, "n_init_samples": 0,
        "pass_X_to_pi": False, "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0, "noise_variance": 8.9e-16,
        "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 1000, "N_LS": 1000, "k": 5,
        "reward_transformation": "neg_log10"}
gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=30, kwargs=spec)
params = {"batch_size": 1200, "max_steps": 1200, "minibatch_size": 60, "n_epochs": 4, "lr": 1e-4, "epsilon": 0.15,
          "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98, "loss_type": "GAElam", "normalize_advs": True,
          "n_workers": 10, "env_id": spec["env_id"], "seed": 0, "env_seeds": list(range(10)), "mlp_mode": "fp16",
          "update_graph": False, "update_backend": backend,
          "policy_options": {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2, "T_idx": -1,
                             "arch_spec_value": [200] * 4, "mlp_mode": "fp16"}}
ppo = PPO(policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=params["policy_options"]), params=params,
          logpath=os.path.join(tempfile.mkdtemp(), "log"), save_interval=10, verbose=True)
ppo.iter_log_str = ""
ppo.batch_recorder.set_worker_weights(ppo.pi)
ppo.batch_recorder.record_batch(gamma=0.98, lam=0.98)
ppo.optimize_on_batch()          # warm-up
ppo.iter_log_str = ""
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ppo.optimize_on_batch()
    torch.cuda.synchronize()
ev = [(e.key, e.count, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total) for e in prof.key_averages()]
ev.sort(key=lambda t: -t[2])
tot = sum(t[2] for t in ev)
print("backend %s: %d kernels, %.1f ms of kernel time for 80 optimiser steps (%.3f ms / step)" % (backend, sum(t[1] for t in ev), tot / 1e3, tot / 80e3))
for k, c, t in ev[:28]:
    print("%9.1f us %6d  %s" % (t, c, k[:110]))
