"""torchrun smoke of the sharded PPO: every rank owns n_workers/world rollout lanes, minibatches are sharded and the
gradient is all-reduced over NCCL; checks that the weights stay bit-identical on all ranks after training.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ppo_multi_gpu.py"""
import os, sys, tempfile, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabo_b200 import gym_compat as gym
from metabo_b200.policies.policies import NeuralAF
from metabo_b200.ppo.ppo import PPO

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
spec = {"env_id": "MetaBO-BRA-v0", "D": 2, "f_type": "BRA-var", "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50},
        "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": 30, "n_init_samples": 0,
        "pass_X_to_pi": False, "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0, "noise_variance": 8.9e-16,
        "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 1000, "N_LS": 1000, "k": 5, "reward_transformation": "neg_log10"}
gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=30, kwargs=spec)
nw = 2 * world
params = {"batch_size": nw * 120, "max_steps": 3 * nw * 120, "minibatch_size": 12 * world, "n_epochs": 2, "lr": 1e-4, "epsilon": 0.15,
          "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98, "loss_type": "GAElam", "normalize_advs": True,
          "n_workers": nw, "env_id": spec["env_id"], "seed": 0, "env_seeds": list(range(nw)), "mlp_mode": "tf32",
          "policy_options": {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2, "T_idx": -1,
                             "arch_spec_value": [200] * 4, "mlp_mode": "tf32"}}
lp = os.path.join(tempfile.gettempdir(), "mbo_ppo_%d" % int(time.time() // 60), "log")
ppo = PPO(policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=params["policy_options"]), params=params, logpath=lp,
          save_interval=10, verbose=True)
t0 = time.time(); ppo.train(); wall = time.time() - t0
flat = torch.cat([p.detach().reshape(-1) for p in ppo.pi.parameters()])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], g) for g in gathered)
if rank == 0:
    st = ppo.stats
    print("world", world, "weights identical on all ranks:", same, "| iterations", st["n_iters"], "| global batch", st["batch_stats"]["size"],
          "| rollout sps %.0f" % st["batch_stats"]["sps"], "| t_optim %.3f s" % st["t_optim"], "| wall %.1f s" % wall)
assert same
dist.destroy_process_group()
