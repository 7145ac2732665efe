"""One PPO rollout batch (bench.py's C3 environment, 1024 episodes x T=30, stochastic fp16 NeuralAF) for an ncu launch
list: MBO_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum ... python tools/rollout_ncu.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metabo_b200 import gym_compat as gym
from metabo_b200.policies.policies import NeuralAF
from metabo_b200.ppo.batchrecorder import BatchRecorder
spec = {"env_id": "MetaBO-BRA-v0", "D": 2, "f_type": "BRA-var",
        "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50},
        "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": 30, "n_init_samples": 0,
        "pass_X_to_pi": False, "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0,
        "noise_variance": 8.9e-16, "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 1000,
        "N_LS": 1000, "k": 5, "reward_transformation": "neg_log10"}
gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=30, kwargs=spec)
opts = {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2, "T_idx": -1,
        "arch_spec_value": [200] * 4, "mlp_mode": "fp16"}
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_ep = int(sys.argv[2]) if len(sys.argv) > 2 else 1024          # episodes per batch (40 = train_metabo_branin.py's batch of 1200)
n_w = 64 if n_ep % 64 == 0 else n_ep
br = BatchRecorder(size=n_ep * 30, env_id=spec["env_id"], env_seeds=list(range(n_w)),
                   policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=opts), n_workers=n_w,
                   deterministic=False, store_states=True, mlp_mode="fp16", device=torch.device("cuda:0"))
for i in range(nb):
    t = br.record_batch(0.98, 0.98)
    print("batch", i, "t_batch %.4f s  %.0f env-steps/s" % (t, n_ep * 30 / t), flush=True)
