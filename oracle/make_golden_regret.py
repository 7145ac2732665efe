"""ORACLE (test infrastructure): mint the 100-episode regret fixtures by running the REFERENCE'S OWN
classes (metabo.environment.metabo_gym.MetaBO + metabo.policies.policies.NeuralAF, imported unmodified through
oracle/ref_loader.py) exactly as evaluate_metabo_{branin,hm3}.py / evaluate.py:146-152 deploy them:
10 workers with env seeds 100..109 (evaluate_metabo_branin.py:127), 10 sequential episodes each, deterministic policy,
shipped weights; reward = simple regret (reward_transformation "none").  The per-episode regret curves are what
plot_results.py:45-63 reduces to median / 30th / 70th percentile per step.

    python -m oracle.make_golden_regret          (authoring container only: needs /root/reference)

Writes tests/golden/regret_{bra,hm3,gps}.npz: regret [100, T] (worker-major, episodes of a worker in order),
y_max [100] (one per episode: pins that the objective draws are the same), seeds [10].
"""
import multiprocessing as mp
import os

import numpy as np
import torch

from .make_golden import CONFIGS, GOLDEN, export_weights
from .ref_loader import load_reference

N_WORKERS, EP_PER_WORKER, SEED0 = 10, 10, 100


def _worker(args):
    name, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    torch.set_num_threads(1)
    load_reference()
    from metabo.environment.metabo_gym import MetaBO
    from metabo.policies.policies import NeuralAF
    cfg = CONFIGS[name]
    env = MetaBO(**cfg["env_spec"])
    opts, sd, _ = export_weights(name, cfg, env.n_features)
    pi = NeuralAF(env.observation_space, env.action_space, deterministic=True, options=opts)
    pi.load_state_dict(sd)
    pi.set_requires_grad(False)
    env.set_af_functions(pi.af)
    env.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    T = cfg["env_spec"]["T"]
    regs, ymaxs = [], []
    for _ in range(EP_PER_WORKER):                 # EnvRunner.record_batch (batchrecorder.py:88-113)
        state = env.reset()
        ymaxs.append(float(np.asarray(env.y_max).reshape(-1)[0]))
        rs = []
        for t in range(T):
            with torch.no_grad():
                action, _ = pi.act(torch.from_numpy(state.astype(np.float32)))
            state, reward, done, _ = env.step(action.numpy())
            rs.append(float(reward))
        assert done
        regs.append(rs)
    return np.array(regs), np.array(ymaxs)


def main():
    import sys
    os.makedirs(GOLDEN, exist_ok=True)
    ctx = mp.get_context("spawn")         # a forked third pool deadlocked in OpenMP after the parent had used torch
    for name in (sys.argv[1:] or ("bra", "hm3", "gps")):
        with ctx.Pool(min(N_WORKERS, os.cpu_count())) as pool:
            res = pool.map(_worker, [(name, SEED0 + i) for i in range(N_WORKERS)], chunksize=1)
        regret = np.concatenate([r for r, _ in res], axis=0)
        y_max = np.concatenate([y for _, y in res])
        out = os.path.join(GOLDEN, "regret_%s.npz" % name)
        np.savez_compressed(out, regret=regret, y_max=y_max, seeds=np.arange(SEED0, SEED0 + N_WORKERS),
                            episodes_per_worker=np.array(EP_PER_WORKER))
        fin = regret[:, -1]
        print(name, regret.shape, "final: median %.3g p30 %.3g p70 %.3g" % (np.median(fin), np.percentile(fin, 30),
                                                                             np.percentile(fin, 70)))


if __name__ == "__main__":
    main()
