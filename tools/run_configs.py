"""End-to-end runs of the BASELINE.json configs through the drop-in layer on one GPU:
C1 evaluate_metabo_branin (100 episodes, deterministic, shipped weights), C2 evaluate_metabo_hm3 at B=1024,
C3 one PPO iteration of train_metabo_branin (batch 1200, 80 optimiser steps), C4 train_metabo_gps rollouts
(T=30 as shipped and T=100).  Prints one JSON object per config (wall time, env-steps/s, regret statistics)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metabo_b200 import gym_compat as gym
from metabo_b200.policies.policies import NeuralAF
from metabo_b200.ppo.batchrecorder import BatchRecorder
from metabo_b200.ppo.ppo import PPO

G = os.path.join(ROOT, "tests", "golden")


def load(npz, mode):
    z = np.load(os.path.join(G, npz), allow_pickle=False)
    opts = json.loads(str(z["__policy_options__"])); opts["mlp_mode"] = mode
    sd = {k: torch.as_tensor(z[k]) for k in z.files if not k.startswith("__")}
    return opts, sd


def recorder(spec, npz, n_workers, n_episodes, seeds0, deterministic, mode, store_states=False):
    gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=spec["T"], kwargs=spec)
    opts, sd = load(npz, mode)
    fn = lambda o, a, d: NeuralAF(o, a, d, options=opts)
    br = BatchRecorder(size=spec["T"] * n_episodes, env_id=spec["env_id"], env_seeds=list(range(seeds0, seeds0 + n_workers)),
                       policy_fn=fn, n_workers=n_workers, deterministic=deterministic, store_states=store_states, mlp_mode=mode)
    pi = fn(br.observation_space, br.action_space, deterministic); pi.load_state_dict(sd)
    br.set_worker_weights(pi)
    return br


BRA = {"env_id": "MetaBO-BRA-v0", "D": 2, "f_type": "BRA-var", "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1},
       "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": 30, "n_init_samples": 0,
       "pass_X_to_pi": False, "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0, "noise_variance": 8.9e-16,
       "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 1000, "N_LS": 1000, "k": 5, "reward_transformation": "none"}
HM3 = {"env_id": "MetaBO-HM3-v0", "D": 3, "f_type": "HM3-var", "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1},
       "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": 30, "n_init_samples": 0,
       "pass_X_to_pi": False, "kernel_lengthscale": [0.716, 0.298, 0.186], "kernel_variance": 0.83, "noise_variance": 1.688e-11,
       "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 2000, "N_LS": 2000, "k": 5, "reward_transformation": "none"}
GPS = {"env_id": "MetaBO-GP-v0", "D": 3, "f_type": "GP",
       "f_opts": {"kernel": "Matern52", "lengthscale_low": 0.05, "lengthscale_high": 0.5, "noise_var_low": 0.1, "noise_var_high": 0.1,
                  "signal_var_low": 1.0, "signal_var_high": 1.0, "min_regret": 1e-6},
       "features": ["posterior_mean", "posterior_std", "incumbent", "timestep_perc", "timestep", "budget"], "T": 30,
       "n_init_samples": 0, "pass_X_to_pi": False, "kernel": "Matern52", "kernel_lengthscale": None, "kernel_variance": None,
       "noise_variance": None, "use_prior_mean_function": True, "local_af_opt": False, "cardinality_domain": 2000,
       "reward_transformation": "neg_log10"}


def run_eval(name, spec, npz, n_workers, n_episodes, mode, cands):
    br = recorder(spec, npz, n_workers, n_episodes, 100, True, mode)
    br.record_batch(1.0, 1.0)                     # eager warm-up
    t_capture = br.record_batch(1.0, 1.0)         # CUDA-graph capture of the T+1 per-step units
    br.record_batch(1.0, 1.0)
    t = br.record_batch(1.0, 1.0)                 # steady state: graph replays
    r = br.buf["rewards"].cpu().numpy()
    out = {"config": name, "episodes": n_episodes, "T": spec["T"], "mlp_mode": mode, "t_batch_s": t, "t_capture_s": t_capture,
           "env_steps_per_s": spec["T"] * n_episodes / t, "af_evals_per_s": spec["T"] * n_episodes * cands / t,
           "final": {"median": float(np.median(r[:, -1])), "p30": float(np.percentile(r[:, -1], 30)),
                     "p70": float(np.percentile(r[:, -1], 70)), "mean": float(r[:, -1].mean())}}
    print(json.dumps(out), flush=True)


def main():
    which = sys.argv[1:] or ["C1", "C2", "C3", "C4"]
    if "C1" in which:
        run_eval("C1 evaluate_metabo_branin (100 episodes, 10 workers, deterministic, weights_1635)", BRA, "weights_bra_1635.npz", 10, 100, "bf16x3", 6927)
        run_eval("C1 same, tf32", BRA, "weights_bra_1635.npz", 10, 100, "tf32", 6927)
        run_eval("C1 same, fp16", BRA, "weights_bra_1635.npz", 10, 100, "fp16", 6927)
    if "C2" in which:
        run_eval("C2 evaluate_metabo_hm3 at B=1024 (64 workers x 16 episodes, deterministic, weights_1988)", HM3, "weights_hm3_1988.npz", 64, 1024, "bf16x3", 13461)
        run_eval("C2 same, tf32", HM3, "weights_hm3_1988.npz", 64, 1024, "tf32", 13461)
        run_eval("C2 same, fp16", HM3, "weights_hm3_1988.npz", 64, 1024, "fp16", 13461)
    if "C4" in which:
        run_eval("C4 train_metabo_gps env, T=30, M=2000, 2x20 policy (SIMT fp32), 1000 episodes", GPS, "weights_gps_1161.npz", 10, 1000, "fp32", 2000)
        g100 = dict(GPS, T=100)
        run_eval("C4 T=100 observations, M=2000, 1000 episodes", g100, "weights_gps_1161.npz", 10, 1000, "fp32", 2000)
        g16k = dict(GPS, T=100, cardinality_domain=16000)
        run_eval("C4 T=100, M=16000, 200 episodes", g16k, "weights_gps_1161.npz", 10, 200, "fp32", 16000)
    for cfg in ((["C3"] if "C3" in which else []) + (["C4train"] if "C4" in which else [])):
        import tempfile
        if cfg == "C3":
            spec = dict(BRA, f_opts={"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50}, reward_transformation="neg_log10")
            popts = {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2, "T_idx": -1,
                     "arch_spec_value": [200] * 4, "mlp_mode": "fp16"}
            desc = ("C3 train_metabo_branin: 6 PPO iterations (batch 1200 = 40 episodes, 80 optimiser steps each, minibatch 60x966 "
                    "candidates; policy / value / Adam on the tcgen05 update kernels)")
            ref = {"rollout_sps": 151.1, "t_optim_s": 2.05, "e2e_steps_per_s": 78.4, "source": "stats_1635"}
        else:
            spec = dict(GPS, env_id="MetaBO-GP-train-v0")
            popts = {"activations": "relu", "arch_spec": [20, 20], "use_value_network": True, "arch_spec_value": [20, 20],
                     "exclude_t_from_policy": True, "exclude_T_from_policy": True, "t_idx": -2, "T_idx": -1, "mlp_mode": "fp32"}
            desc = ("C4 train_metabo_gps: 6 PPO iterations (batch 1200, minibatch 60x2000 candidates, 4->20->20->1 policy and "
                    "2->20->20->1 value net on the CUDA-core update kernels mbo_ppo_*_simt)")
            ref = {"rollout_sps": 397.5, "t_optim_s": 4.33, "e2e_steps_per_s": 149.7, "source": "stats_1161"}
        gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=30, kwargs=spec)
        params = {"batch_size": 1200, "max_steps": 6 * 1200, "minibatch_size": 60, "n_epochs": 4, "lr": 1e-4, "epsilon": 0.15,
                  "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98, "loss_type": "GAElam", "normalize_advs": True,
                  "n_workers": 10, "env_id": spec["env_id"], "seed": 0, "env_seeds": list(range(10)), "mlp_mode": popts["mlp_mode"],
                  "policy_options": popts}
        lp = os.path.join(tempfile.mkdtemp(), "log")
        ppo = PPO(policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=params["policy_options"]), params=params, logpath=lp,
                  save_interval=10, verbose=True)
        t0 = time.time(); ppo.train(); wall = time.time() - t0
        st = ppo.stats
        print(json.dumps({"config": desc, "update_kernels": type(ppo._pg).__name__,
                          "steady_state_e2e_steps_per_s": 1200 / (st["batch_stats"]["t_batch"] + st["t_optim"]),
                          "wall_s": wall, "t_batch_s": st["batch_stats"]["t_batch"],
                          "rollout_sps": st["batch_stats"]["sps"], "t_optim_s": st["t_optim"],
                          "e2e_steps_per_s": st["n_timesteps"] / st["t_train"], "avg_step_rews": list(map(float, st["avg_step_rews"])),
                          "reference_shipped_stats": ref}), flush=True)

if __name__ == "__main__":
    main()
