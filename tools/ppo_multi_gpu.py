"""torchrun driver of the data-parallel PPO update: every rank owns n_workers/world rollout lanes, every minibatch of the
reference's size (60) is sharded over the ranks, the flat gradient is exchanged inside the Adam kernel over NVLink peer
memory (`--exchange peer`, default, whole epoch = one CUDA graph per rank) or by an eager NCCL all-reduce
(`--exchange nccl`, the round-1 path, needs minibatch % world == 0).  Checks that the weights stay bit-identical on all
ranks and prints one JSON line with the per-iteration times (max over ranks).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ppo_multi_gpu.py
"""
import argparse
import copy
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabo_b200 import gym_compat as gym
from metabo_b200.policies.policies import NeuralAF
from metabo_b200.ppo.ppo import PPO

ap = argparse.ArgumentParser()
ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"])
ap.add_argument("--episodes", type=int, default=40, help="episodes per iteration over ALL ranks (reference C3: 40)")
ap.add_argument("--minibatch", type=int, default=60)
ap.add_argument("--epochs", type=int, default=4)
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
spec = {"env_id": "MetaBO-BRA-v0", "D": 2, "f_type": "BRA-var", "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50},
        "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": 30, "n_init_samples": 0,
        "pass_X_to_pi": False, "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0, "noise_variance": 8.9e-16,
        "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 1000, "N_LS": 1000, "k": 5, "reward_transformation": "neg_log10"}
gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=30, kwargs=spec)
nw = args.episodes if args.episodes <= 64 * world else 64 * world
params = {"batch_size": args.episodes * 30, "max_steps": 10 ** 12, "minibatch_size": args.minibatch, "n_epochs": args.epochs, "lr": 1e-4,
          "epsilon": 0.15, "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98, "loss_type": "GAElam",
          "normalize_advs": True, "n_workers": nw, "env_id": spec["env_id"], "seed": 0, "env_seeds": list(range(nw)), "mlp_mode": "fp16",
          "dp_exchange": args.exchange,
          "policy_options": {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2, "T_idx": -1,
                             "arch_spec_value": [200] * 4, "mlp_mode": "fp16"}}
lp = os.path.join(tempfile.gettempdir(), "mbo_ppo_%s_%d" % (os.environ.get("MASTER_PORT", "0"), int(time.time() // 600)), args.exchange)
if rank == 0:
    import shutil
    shutil.rmtree(lp, ignore_errors=True)
if world > 1:
    dist.barrier()
ppo = PPO(policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=params["policy_options"]), params=params, logpath=os.path.join(lp, "log"),
          save_interval=10 ** 9, verbose=True)
br = ppo.batch_recorder
times = []
for it in range(args.iters):
    ppo.iter_log_str = ""
    if world > 1:
        dist.barrier()
    tw = br.set_worker_weights(copy.deepcopy(ppo.pi))
    tb = br.record_batch(gamma=0.98, lam=0.98)
    to = ppo.optimize_on_batch()
    tt = torch.tensor([tw, tb, to], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    times.append(tt.tolist())
flat = torch.cat([p.detach().reshape(-1) for p in ppo.pi.parameters()])
same = True
if world > 1:
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
if rank == 0:
    n_steps = ppo.stats["n_optsteps"] // args.iters
    tw, tb, to = [float(np.mean([t[i] for t in times[2:]])) for i in range(3)]
    print(json.dumps({"world": world, "exchange": args.exchange if world > 1 else "none", "weights_identical_on_all_ranks": bool(same),
                      "global_batch_steps": params["batch_size"], "minibatch": args.minibatch, "epochs": args.epochs,
                      "optimizer_steps_per_iteration": n_steps, "t_weights_s": tw, "t_batch_s": tb, "t_optim_s": to,
                      "ms_per_optimizer_step": 1e3 * to / n_steps, "rollout_sps": params["batch_size"] / tb,
                      "end_to_end_sps": params["batch_size"] / (tw + tb + to),
                      "epoch_graph": bool(ppo._ep is not None and ppo._ep.get("graph") is not None),
                      "iteration_times": times}))
assert same
if world > 1:
    dist.destroy_process_group()
