"""eval_experiment drop-in for the MetaBO policy (metabo/eval/evaluate.py:59-167): loads a trained
NeuralAF from the reference's checkpoint files (params_N / weights_N / stats_N, shipped iclr2020
runs included), plays T * n_episodes steps through the GPU BatchRecorder and stores the same
`Result` pickle.  Baselines: EI / GP-UCB / PI run as closed-form epilogues of the same GPU AF rounds, TAF-ME / TAF-RANKING
(policies.py:310-468) through the host-driven env with their 50 source GPs on the batched GP kernel; "Random" is
`MetaBO.get_random_sampling_reward` (evaluate.py:154-160); the data-file baselines EPS-GREEDY / GMM-UCB need the authors'
data files, are outside the scope and raise."""
import os
import pickle as pkl
from collections import namedtuple
from datetime import datetime

import numpy as np
import torch

from .. import gym_compat as gym
from ..policies.policies import NeuralAF, EI, PI, UCB
from ..policies.taf import TAF
from ..ppo.batchrecorder import BatchRecorder

Result = namedtuple("Result",
                    "logpath env_id env_specs policy policy_specs deterministic load_iter T n_episodes rewards")


def load_metabo_policy(logpath, load_iter, env, device, deterministic):
    with open(os.path.join(logpath, "params_" + str(load_iter)), "rb") as f:
        train_params = pkl.load(f)
    pi = NeuralAF(observation_space=env.observation_space, action_space=env.action_space,
                  deterministic=deterministic, options=train_params["policy_options"])
    with open(os.path.join(logpath, "weights_" + str(load_iter)), "rb") as f:
        pi.load_state_dict(torch.load(f, map_location="cpu", weights_only=False))
    with open(os.path.join(logpath, "stats_" + str(load_iter)), "rb") as f:
        stats = pkl.load(f)
    return pi, train_params, stats


def eval_experiment(eval_spec):
    env_id = eval_spec["env_id"]
    policy = eval_spec["policy"]
    if policy not in ("MetaBO", "EI", "GP-UCB", "PI", "TAF-ME", "TAF-RANKING", "Random"):
        raise ValueError("Unknown policy!" if policy not in ("EPS-GREEDY", "GMM-UCB")
                         else "baseline '%s' needs the authors' data files and is outside the scope" % policy)
    logpath, savepath = eval_spec["logpath"], eval_spec["savepath"]
    n_workers, n_episodes, T = eval_spec["n_workers"], eval_spec["n_episodes"], eval_spec["T"]
    assert n_episodes % n_workers == 0
    load_iter, deterministic = eval_spec.get("load_iter"), eval_spec.get("deterministic", True)
    os.makedirs(savepath, exist_ok=True)
    env_seeds = eval_spec["env_seed_offset"] + np.arange(n_workers)
    env_specs = gym.registry.env_specs[env_id]._kwargs
    feature_order = ["posterior_mean", "posterior_std", "incumbent", "timestep"]     # metabo_gym.py:95
    if policy == "MetaBO":
        with open(os.path.join(logpath, "params_" + str(load_iter)), "rb") as f:
            policy_specs = pkl.load(f)

        def policy_fn(osp, asp, det):
            return NeuralAF(observation_space=osp, action_space=asp, deterministic=det,
                            options=policy_specs["policy_options"])
    else:
        # closed-form baselines on the same GP kernels (evaluate.py:102-118)
        policy_specs = eval_spec.get("policy_specs", {})
        policy_fn = None
        if policy == "TAF-ME":
            policy_fn = lambda *_: TAF(datafile=policy_specs["TAF_datafile"], mode="me")               # evaluate.py:112-113
        elif policy == "TAF-RANKING":
            policy_fn = lambda *_: TAF(datafile=policy_specs["TAF_datafile"], mode="ranking", rho=1.0)  # :114-115
        elif policy == "EI":
            policy_fn = lambda *_: EI(feature_order=feature_order)
        elif policy == "PI":
            policy_fn = lambda *_: PI(feature_order=feature_order, xi=policy_specs["xi"])
        elif policy == "GP-UCB":
            policy_fn = lambda *_: UCB(feature_order=feature_order, kappa=policy_specs["kappa"], D=env_specs["D"],
                                       delta=policy_specs["delta"])

    if policy == "Random":                       # evaluate.py:154-160
        env = gym.make(env_id)
        env.seed(eval_spec["env_seed_offset"])
        rewards = []
        for _ in range(n_episodes):
            rewards = rewards + env.unwrapped.get_random_sampling_reward()
        env.close()
        rewards = tuple(rewards)
    else:
        br = BatchRecorder(size=T * n_episodes, env_id=env_id, env_seeds=list(env_seeds), policy_fn=policy_fn,
                           n_workers=n_workers, deterministic=deterministic, store_states=False)
        if policy == "MetaBO":
            pi, _, _ = load_metabo_policy(logpath, load_iter, br, "cpu", deterministic)
            br.set_worker_weights(pi=pi)
        br.record_batch(gamma=1.0, lam=1.0)
        rewards = tuple(t.reward for t in br.memory)
        br.cleanup()
    with open(os.path.join(savepath, "000_eval_overview_{}.txt".format(policy)), "w") as f:
        print("Evaluation timestamp: {}\nEnvironment-ID: {}\nEnvironment-seeds:\n{}".format(
            datetime.strftime(datetime.now(), "%Y-%m-%d-%H-%M-%S"), env_id, env_seeds), file=f)
    result = Result(logpath=logpath, env_id=env_id, env_specs=env_specs, policy=policy, policy_specs=policy_specs,
                    deterministic=deterministic, load_iter=load_iter, T=T, n_episodes=n_episodes, rewards=rewards)
    fn = "result_metabo_iter_{:04d}".format(load_iter) if policy == "MetaBO" else "result_{}".format(policy)   # evaluate.py:165
    with open(os.path.join(savepath, fn), "wb") as f:
        pkl.dump(result, f)
    return result
