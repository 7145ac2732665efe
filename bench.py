#!/usr/bin/env python
"""bench.py -- AF evals/s of the MetaBO hot path on B200 (see BASELINE.json).

A "step" is one AF-optimisation round (what the reference runs per environment step,
metabo_gym.py:195-217) for B parallel BO episodes: GP refit, then three dependent candidate phases
(multistart grid -> k local Sobol grids -> xi_t), each = GP posterior + NeuralAF MLP + reduction,
ending in the greedy action.  Workload = configs[1] of BASELINE.json: evaluate_metabo_hm3.py
(Hartmann-3, D=3, F=7, 4x200 ReLU policy `weights_1988`, N_MS=1728, N_LS=2000, k=5 ->
1728 + 10000 + 1733 = 13461 candidates per episode and round) at B=1024 episodes per GPU with
n = T = 30 observations per episode (the most expensive round of an episode).

  value   AF evals/s with the observations already resident in HBM
  e2e     same metric through the public engine API with HOST buffers (H2D of X/Y/n/scalars and
          D2H of the chosen actions inside the timed region)
  --impl reference   the reference algorithm's CPU path (oracle port, all host cores)

Multi-GPU: episodes are independent -> every rank runs its own B episodes (weak scaling, no
data-path collective); rank 0 prints one JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HM3 = dict(D=3, T=30, N_MS=2000, N_LS=2000, k=5, kernel="RBF",
           features=["posterior_mean", "posterior_std", "timestep", "budget", "x"],
           lengthscale=[0.716, 0.298, 0.186], variance=0.83, noise=1.688e-11)
MLP_FLOPS = 2 * (7 * 200 + 3 * 200 * 200 + 200)       # 243 200 per candidate


def gp_flops(n, D):
    return n * n + n * (3 * D + 8)


def load_weights():
    """shipped HM3 policy if the fixture is present, else N(0, 0.01) weights of the same shape."""
    p = os.path.join(ROOT, "tests", "golden", "weights_hm3_1988.npz")
    if os.path.exists(p):
        z = np.load(p)
        W = [z["policy_net.fc.%d.weight" % i] for i in range(5)]
        b = [z["policy_net.fc.%d.bias" % i] for i in range(5)]
        return W, b, "shipped iclr2020 hm3/M_50 weights_1988"
    rng = np.random.RandomState(0)
    dims = [7, 200, 200, 200, 200, 1]
    W = [rng.normal(0, 0.01, (dims[i + 1], dims[i])).astype(np.float32) for i in range(5)]
    b = [np.zeros(dims[i + 1], dtype=np.float32) for i in range(5)]
    return W, b, "random N(0,0.01) init"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on host cores (oracle port; the reference's GPy is not
# installable).  One forked worker per core, every worker single-threaded exactly as
# the reference pins its workers (OMP_NUM_THREADS=1 at import: ppo.py:23, batchrecorder.py:24;
# torch.set_num_threads(1): batchrecorder.py:162).  The pinning has to happen BEFORE numpy / torch
# are imported, so the arm runs in a fresh child interpreter (`bench.py --cpu-child`) whose
# environment is set first; the child keeps one persistent pool and prints one JSON line per step.
# ------------------------------------------------------------------------------------------------
PIN_ENV = {"OMP_NUM_THREADS": "1", "OPENBLAS_NUM_THREADS": "1", "MKL_NUM_THREADS": "1", "NUMEXPR_NUM_THREADS": "1",
           "VECLIB_MAXIMUM_THREADS": "1"}
_W = {}


def _cpu_worker_init():
    import torch
    torch.set_num_threads(1)
    from oracle import metabo_oracle as O
    _W["O"] = O
    _W["pw"] = O.PolicyWeights.from_npz(os.path.join(ROOT, "tests", "golden", "weights_hm3_1988.npz"))
    _W["grids"] = O.Grids(HM3["D"], N_MS=HM3["N_MS"], N_LS=HM3["N_LS"], k=HM3["k"])


def _cpu_worker(args):
    seed, rounds, n = args
    O, pw, grids = _W["O"], _W["pw"], _W["grids"]
    rng = np.random.RandomState(seed)
    cands = 0
    t0 = time.perf_counter()
    for _ in range(rounds):
        X = rng.rand(n, HM3["D"])
        Y = O.hm3_var(X, rng.uniform(-0.1, 0.1, 3), rng.uniform(0.9, 1.1))[:, 0]
        gp = O.EpisodeGP(HM3["D"], "RBF", HM3["variance"], HM3["lengthscale"], HM3["noise"])
        gp.set_data(X, Y)
        r = O.af_round(gp, grids, pw, HM3["features"], n, HM3["T"])
        cands += grids.N_MS + grids.k * grids.N_LS + r["xi_t"].shape[0]
    return cands, time.perf_counter() - t0


def cpu_child_main(argv):
    """child interpreter of the CPU arm: `bench.py --cpu-child N_OBS CORES ROUNDS_PER_CORE REPS` (environment already
    pinned by the parent, asserted here); prints one JSON line per repetition."""
    n, cores, rounds, reps = (int(a) for a in argv)
    for k, v in PIN_ENV.items():
        assert os.environ.get(k) == v, "CPU arm must start with %s=%s" % (k, v)
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_worker_init) as pool:
        pool.map(_cpu_worker, [(i, 0, n) for i in range(cores)])         # all workers up, weights loaded
        for rep in range(reps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(1000 + rep * cores + i, rounds, n) for i in range(cores)], chunksize=1)
            wall = time.perf_counter() - t0
            print(json.dumps({"cands": sum(r[0] for r in res), "wall": wall, "busy_max": max(r[1] for r in res)}), flush=True)


def cpu_arm(rounds_per_core, n, reps=1, cores=None):
    """-> list of (candidates, seconds) per repetition, and the description dict.  seconds = wall time of the
    parallel map over the persistent single-threaded workers (includes task dispatch, excludes interpreter start)."""
    cores = cores or os.cpu_count()
    env = dict(os.environ)
    env.update(PIN_ENV)
    env.pop("CUDA_VISIBLE_DEVICES", None)
    env["CUDA_VISIBLE_DEVICES"] = ""                      # the CPU arm never touches a GPU
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-child", str(n), str(cores),
                          str(rounds_per_core), str(reps)], env=env, check=True, stdout=subprocess.PIPE, cwd=ROOT)
    reps_out = [json.loads(l) for l in out.stdout.decode().splitlines() if l.startswith("{")]
    assert len(reps_out) == reps
    desc = dict(unit="AF evals/s", cores=cores, kind="port",
                sample="%d episode-rounds per step (n=%d obs, 13461 candidates each; %d per worker) on %d forked "
                       "single-threaded workers (OMP/OPENBLAS/MKL_NUM_THREADS=1 set before import, torch 1 thread, as "
                       "ppo.py:23 / batchrecorder.py:24,162): oracle port of the GPy-route GP + torch-CPU fp32 NeuralAF"
                       % (rounds_per_core * cores, n, rounds_per_core, cores))
    return [(r["cands"], r["wall"]) for r in reps_out], desc


def run_reference_arm(args, rank):
    if rank != 0:
        return
    rounds = 4
    reps, desc = cpu_arm(rounds, args.n_obs, reps=args.warmup + args.steps)
    timed = reps[args.warmup:]
    t = sum(w for _, w in timed)
    v = sum(c for c, _ in timed) / t
    line = {"metric": METRIC, "impl": "reference", "value": v,
            "unit": "AF evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 GP + f32 MLP", "data": "synthetic",
            "config": workload_config(args),
            "reference_step": "%d episode-rounds of the workload per step (bounded sample; the GPU arm's step is "
                              "episodes_per_gpu rounds per GPU)" % (rounds * desc["cores"]),
            "cpu_baseline": dict(desc, value=v),
            "e2e": {"value": v, "unit": "AF evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


METRIC = "AF evals/sec (GP posterior+NeuralAF per candidate)"


def workload_config(args):
    """identical in both arms (the driver compares the two lines' `config`)"""
    return {"workload": "evaluate_metabo_hm3.py AF-optimisation round (BASELINE configs[1]): D=3, F=7, 4x200 ReLU "
                        "policy, candidates 1728+10000+1733 per episode, n=%d observations" % args.n_obs,
            "episodes_per_gpu": args.B, "candidates_per_episode_round": 13461, "n_obs": args.n_obs,
            "gp_precision": args.gp, "mlp_mode": args.mlp,
            "l2": "flushed between timed steps (256 MiB write)"}


class ClockSampler:
    """SM clock + throttle reasons sampled (NVML, every 10 ms) DURING the timed region."""
    BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
            0x80: "hw_power_brake_slowdown"}

    def __init__(self, device):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = threading.Event()
        self.h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(device).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device.index or 0)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:          # pragma: no cover
            sys.stderr.write("[bench] NVML unavailable: %s\n" % e)
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        if self.h is None:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self.stop_flag.set()
        self.th.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_min_mhz": float(min(self.samples)),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def ppo_rollout_rate(dev, world, rank):
    """Supplementary BASELINE metric: PPO env-steps/s exactly as the reference defines `sps`
    (batch size / t_batch, rollout only, ppo.py:213) on the train_metabo_branin.py environment
    (stochastic policy, M=50 training functions, neg_log10 reward) with 1024 episodes per GPU."""
    import torch
    import torch.distributed as dist
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
    n_workers, eps = 64, 16
    size = n_workers * eps * 30
    br = BatchRecorder(size=size, env_id=spec["env_id"], env_seeds=list(range(rank * n_workers, (rank + 1) * n_workers)),
                       policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=opts), n_workers=n_workers,
                       deterministic=False, store_states=True, mlp_mode="fp16", device=dev)
    br.record_batch(0.98, 0.98)                       # eager warm-up
    br.record_batch(0.98, 0.98)                       # CUDA-graph capture of the per-step units
    br.record_batch(0.98, 0.98)                       # first replay
    if world > 1:
        dist.barrier()
    t = br.record_batch(0.98, 0.98)
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt[0])
    return {"env_steps_per_s": world * size / t, "t_batch_s": t, "batch_steps": world * size,
            "config": "train_metabo_branin.py env (C3), stochastic NeuralAF 4x200, 1024 episodes x T=30 per GPU, "
                      "6927 candidates per step (phase 3 reuses the multistart results of phase 1: 5199 evaluated), fp16 MLP + fp64 GP, "
                      "CUDA-graph replay of the episode loop, every transition's state [966, 6] fp32 stored on the device "
                      "(store_states=True, what PPO trains on; 712 MB per batch per GPU)"}


def ppo_train_rate(dev, world, rank, episodes_per_gpu, n_epochs, label, minibatch=60):
    """BASELINE configs[2] (train_metabo_branin.py:34-113) end to end under world = N: PPO iterations = weight sync +
    rollouts sharded over the ranks (record_batch) + optimize_on_batch with every minibatch of the reference's size
    (60) sharded over the ranks and the gradient exchanged inside the Adam kernel over NVLink peer memory.  Two warm
    iterations (eager, graph capture), then two timed ones; times are the max over ranks."""
    import copy
    import shutil
    import torch
    import torch.distributed as dist
    from metabo_b200 import gym_compat as gym
    from metabo_b200.policies.policies import NeuralAF
    from metabo_b200.ppo.ppo import PPO
    T = 30
    spec = {"env_id": "MetaBO-BRA-train-%s-v0" % episodes_per_gpu, "D": 2, "f_type": "BRA-var",
            "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50},
            "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": T, "n_init_samples": 0,
            "pass_X_to_pi": False, "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0,
            "noise_variance": 8.9e-16, "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 1000,
            "N_LS": 1000, "k": 5, "reward_transformation": "neg_log10"}
    gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=T, kwargs=spec)
    if episodes_per_gpu is None:            # the reference batch: 1200 transitions = 40 episodes in total
        n_workers, batch = 40, 1200
    else:
        n_workers, batch = 64 * world, episodes_per_gpu * T * world
    params = {"batch_size": batch, "max_steps": 10 ** 12, "minibatch_size": minibatch, "n_epochs": n_epochs, "lr": 1e-4,
              "epsilon": 0.15, "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98,
              "loss_type": "GAElam", "normalize_advs": True, "n_workers": n_workers, "env_id": spec["env_id"], "seed": 0,
              "env_seeds": list(range(n_workers)), "mlp_mode": "fp16",
              "policy_options": {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2,
                                 "T_idx": -1, "arch_spec_value": [200] * 4, "mlp_mode": "fp16"}}
    logdir = os.path.join(tempfile.gettempdir(), "mbo_bench_ppo_%d_%s" % (os.getpid() if world == 1 else int(os.environ.get("MASTER_PORT", "0")), label))
    if rank == 0:
        shutil.rmtree(logdir, ignore_errors=True)
    if world > 1:
        dist.barrier()
    ppo = PPO(policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=params["policy_options"]), params=params,
              logpath=os.path.join(logdir, "log"), save_interval=10 ** 9, verbose=True)
    br = ppo.batch_recorder
    times = []
    for it in range(4):
        ppo.iter_log_str = ""
        if world > 1:
            dist.barrier()
        tw = br.set_worker_weights(copy.deepcopy(ppo.pi))
        tb = br.record_batch(gamma=params["gamma"], lam=params["lambda"])
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
    tw, tb, to = [float(np.mean([t[i] for t in times[2:]])) for i in range(3)]
    n_steps = n_epochs * len(ppo._ep["bounds"]) if ppo._ep is not None else None
    if rank == 0:
        shutil.rmtree(logdir, ignore_errors=True)
    return {"config": label, "global_batch_steps": batch, "minibatch_size": minibatch, "n_epochs": n_epochs,
            "optimizer_steps_per_iteration": n_steps, "t_weights_s": tw, "t_batch_s": tb, "t_optim_s": to,
            "rollout_env_steps_per_s": batch / tb, "end_to_end_env_steps_per_s": batch / (tw + tb + to),
            "ms_per_optimizer_step": 1e3 * to / n_steps if n_steps else None,
            "gradient_exchange": ("none (1 rank)" if world == 1 else
                                  "mbo_comm_allreduce_adam: one-shot pull over NVLink peer memory fused into the Adam kernel, "
                                  "inside the captured epoch graph" if ppo.comm is not None else "NCCL all_reduce, eager"),
            "weights_identical_on_all_ranks": bool(same)}


def ppo_update_rate(dev):
    """Supplementary: one PPO optimiser step's policy-net forward + backward on a train_metabo_branin.py minibatch
    (60 transitions x 966 candidates, 4x200 ReLU policy): the hand-written kernels behind mbo_ppo_policy_grad
    (3xTF32 forward, TF32 backward) timed with CUDA events, next to torch's fp32 autograd of the same loss."""
    import torch
    from metabo_b200 import ops
    E, M, F, H = 60, 966, 6, 200
    g = torch.Generator().manual_seed(0)
    dims = [F, H, H, H, H, 1]
    W = [(torch.randn(dims[i + 1], dims[i], generator=g) * 0.08).to(dev).requires_grad_() for i in range(5)]
    b = [(torch.randn(dims[i + 1], generator=g) * 0.05).to(dev).requires_grad_() for i in range(5)]
    st = torch.randn(E, M, F, generator=g).to(dev)
    ac = torch.randint(0, M, (E,), generator=g).to(dev)
    adv = torch.randn(E, generator=g).to(dev)
    pg = ops.PPOPolicyGrad(E, M, F, list(range(F)), H, dev)
    Wd, bd = [w.detach() for w in W], [x.detach() for x in b]
    dW, db = [torch.zeros_like(w) for w in Wd], [torch.zeros_like(x) for x in bd]
    a32 = ac.to(torch.int32)
    lpo, _ = pg.logp_entropy(st, Wd, bd, a32)

    def native():
        pg(st, Wd, bd, a32, adv, lpo, 0.15, 0.01, E, dW, db)

    def autograd():
        for p_ in W + b:
            p_.grad = None
        h = st.reshape(E * M, F)
        for l in range(4):
            h = torch.relu(h @ W[l].t() + b[l])
        logits = (h @ W[4].t() + b[4]).reshape(E, M)
        norm = logits - torch.logsumexp(logits, -1, keepdim=True)
        lp = norm.gather(-1, ac[:, None]).squeeze(-1)
        ent = -(norm * norm.exp()).sum(-1)
        ratio = torch.exp(lp - lpo)
        loss = -torch.min(ratio * adv, ratio.clamp(0.85, 1.15) * adv).sum() / E - 0.01 * ent.sum() / E
        loss.backward()

    def timeit(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    prec = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("highest")
    t_nat, t_ref = timeit(native, 30), timeit(autograd, 10)
    torch.set_float32_matmul_precision(prec)
    pg.check()
    flops = 3 * 2 * E * M * (F * H + 3 * H * H + H)
    return {"policy_grad_ms": t_nat, "algorithmic_tflops": flops / t_nat / 1e9, "torch_fp32_autograd_ms": t_ref,
            "launches_per_step": 9,
            "config": "train_metabo_branin.py minibatch (C3): 60 x 966 candidates, F=6, 4x200 ReLU policy; forward = one launch "
                      "of the tcgen05 policy kernel in bf16x3 (fp32-level logits; activations + ReLU words stored for the "
                      "backward), backward single-pass TF32, deterministic split-K"}


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--cpu-child":
        return cpu_child_main(sys.argv[2:])
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--B", type=int, default=1024, help="episodes per GPU")
    ap.add_argument("--n-obs", dest="n_obs", type=int, default=30)
    ap.add_argument("--gp", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--mlp", default="fp16", choices=["fp32", "tf32", "bf16x3", "fp16"],
                    help="arithmetic of the NeuralAF kernel for the headline line; an unavailable mode is an error")
    ap.add_argument("--no-modes", action="store_true", help="skip the tf32 / bf16x3 step times printed next to the headline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ppo", action="store_true", help="skip the supplementary PPO rollout env-steps/s figure")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from metabo_b200 import _lib as L
    from metabo_b200.engine import AFEngine
    from metabo_b200.sobol import i4_sobol_generate
    from metabo_b200 import objectives

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, D, T, n = args.B, HM3["D"], HM3["T"], args.n_obs
    gp_prec = {"fp64": L.GP_FP64, "fp32": L.GP_FP32}[args.gp]
    W, bvec, wdesc = load_weights()
    Wt = [torch.as_tensor(w, device=dev) for w in W]
    bt = [torch.as_tensor(x, device=dev) for x in bvec]

    def make_engine(mode):
        e = AFEngine(B, D, T, HM3["features"], kernel=HM3["kernel"], N_MS=HM3["N_MS"], N_LS=HM3["N_LS"], k=HM3["k"],
                     local_search_grid=i4_sobol_generate(D, HM3["N_LS"]), gp_precision=gp_prec, mlp_mode=mode,
                     device=dev)
        e.set_policy(Wt, bt, "relu")
        e.set_hyperparameters(HM3["lengthscale"], HM3["variance"], HM3["noise"])
        return e

    # synthetic observations, generated on the device for the resident arm and mirrored in pinned
    # host memory for the e2e arm
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    X = torch.zeros((B, T, D), dtype=torch.float64, device=dev)
    X[:, :n] = torch.rand((B, n, D), generator=g, dtype=torch.float64, device=dev)
    tr = (torch.rand((B, D), generator=g, dtype=torch.float64, device=dev) - 0.5) * 0.2
    sc = 0.9 + 0.2 * torch.rand((B,), generator=g, dtype=torch.float64, device=dev)
    Y = torch.zeros((B, T), dtype=torch.float64, device=dev)
    Y[:, :n] = objectives.hm3_var(X[:, :n], tr[:, None, :], sc[:, None])
    nvec = torch.full((B,), n, dtype=torch.int32, device=dev)
    inc = Y[:, :n].max(dim=1).values.float() if n > 0 else torch.zeros(B, device=dev)
    tvec = torch.full((B,), float(n), device=dev)
    Tvec = torch.full((B,), float(T), device=dev)

    modes = L.MLP_MODES

    def ready_engine(m):
        """no fallback ladder: a mode the library cannot run raises (MboError) and the bench fails"""
        e = make_engine(modes[m])
        if e.mlp_mode != modes[m]:
            raise L.MboError("mlp mode %s is not available for this policy" % m)
        e.set_data(X, Y, nvec, n_active_max=n)
        e.set_episode_scalars(inc, tvec, Tvec)
        e.af_round(want_logits=False)
        torch.cuda.synchronize()
        e.policy.check()
        return e

    eng = ready_engine(args.mlp)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False, eng=eng):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(dev) if profile else None
        evs = []
        l0 = eng.launches
        for _ in range(steps):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = sum(s.elapsed_time(e) for s, e in evs)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, clocks, eng.launches - l0

    # ---- resident arm ----
    eng.kernel_events = []          # (tag, start, end) around the MLP launches, see engine.evaluate
    ms, clocks, launches = timed(lambda: eng.af_round(want_logits=False), args.steps, args.warmup, profile=True)
    kev = eng.kernel_events
    eng.kernel_events = None
    cands_step = B * 13461
    value = world * cands_step * args.steps / (ms / 1e3)

    # dominant kernel: the NeuralAF MLP launch on the 10 000-candidate local phase
    mlp_big = [s.elapsed_time(e) for tag, s, e in kev if tag == "mlp_M%d" % (HM3["k"] * HM3["N_LS"])]
    gp_big = [s.elapsed_time(e) for tag, s, e in kev if tag == "gp_M%d" % (HM3["k"] * HM3["N_LS"])]
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    # the kernel is timed alone (CUDA events around the one launch) inside a sub-second region at max clock: the burst
    # figure is the denominator; the sustained one is given beside it
    peak_tf = peaks.get("bf16_tflops", 1620.0)
    peak_sus = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone by CUDA events)" if peaks else \
        "fallback 1.62 PFLOP/s burst / 1.4 sustained (B200_PROFILING.md)"
    roof = None
    if mlp_big:
        dur = float(np.mean(mlp_big)) / 1e3
        ach = B * HM3["k"] * HM3["N_LS"] * MLP_FLOPS / dur / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tp) and B == 1024:
            tj = json.load(open(tp)).get(args.mlp)
            if tj:
                traffic = tj["dram_read_bytes"] + tj["dram_write_bytes"]     # one ncu --set full capture of this launch shape
        roof = {"bound": "tensor", "kernel": "NeuralAF MLP (%s), B x 10000 local-grid candidates" % args.mlp,
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                "peak_sustained": peak_sus, "frac_of_sustained": ach / peak_sus,
                "traffic_unit": "bytes per launch (dram read+write, one ncu --set full capture of this build: "
                                "profiles/r02_traffic.json)",
                "peak_source": peak_src, "avg_launch_ms": dur * 1e3,
                "format_ceiling_frac": {"fp16": 1.0, "tf32": 0.5, "bf16x3": 1.0 / 3.0, "fp32": None}[args.mlp]}
        if gp_big:
            gdur = float(np.mean(gp_big)) / 1e3
            g_ach = B * HM3["k"] * HM3["N_LS"] * gp_flops(n, D) / gdur / 1e12
            fp64_peak = 148 * 64 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12      # derived: not in MEASURED_PEAKS
            roof["gp_posterior"] = {"kernel": "GP posterior (fp64, DMMA triangular mat-vec), B x 10000 local-grid candidates",
                                    "avg_launch_ms": gdur * 1e3, "precision": args.gp, "achieved_tflops": g_ach,
                                    "bound": "fp64 pipe (DFMA and DMMA share it: tools/micro/dmma_bw.cu)",
                                    "peak_tflops_derived": fp64_peak, "frac": g_ach / fp64_peak,
                                    "algorithmic_flops_per_candidate": gp_flops(n, D)}

    # ---- e2e arm: host buffers -> engine -> host actions ----
    hX = X.cpu().pin_memory(); hY = Y.cpu().pin_memory(); hn = nvec.cpu().pin_memory()
    hs = torch.stack([inc.cpu(), tvec.cpu(), Tvec.cpu()]).pin_memory()
    h_act = torch.empty((B,), dtype=torch.int32).pin_memory()
    h_x = torch.empty((B, D), dtype=torch.float64).pin_memory()
    dX = torch.empty_like(X); dY = torch.empty_like(Y); dn = torch.empty_like(nvec)
    ds = torch.empty((3, B), device=dev)

    def e2e_step():
        dX.copy_(hX, non_blocking=True); dY.copy_(hY, non_blocking=True); dn.copy_(hn, non_blocking=True)
        ds.copy_(hs, non_blocking=True)
        eng.set_data(dX, dY, dn, n_active_max=n)
        eng.set_episode_scalars(ds[0], ds[1], ds[2])
        out = eng.af_round(want_logits=False)
        h_act.copy_(out["action"], non_blocking=True)
        h_x.copy_(out["x_action"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_e, _, _ = timed(e2e_step, args.steps, args.warmup)
    h2d = hX.numel() * 8 + hY.numel() * 8 + hn.numel() * 4 + hs.numel() * 4
    d2h = h_act.numel() * 4 + h_x.numel() * 8
    e2e_val = world * cands_step * args.steps / (ms_e / 1e3)

    # ---- the other arithmetic modes of the same step (bf16x3 = the fp32-class default of the drop-in classes) ----
    modes_info = None
    if not args.no_modes:
        modes_info = {args.mlp: {"ms_per_step": ms / args.steps, "AF_evals_per_s": value, "headline": True}}
        for m in ("tf32", "bf16x3"):
            if m == args.mlp:
                continue
            e2 = ready_engine(m)
            k_steps = min(args.steps, 10)
            ms_m, _, _ = timed(lambda: e2.af_round(want_logits=False), k_steps, 3, eng=e2)
            e2.policy.check()
            modes_info[m] = {"ms_per_step": ms_m / k_steps, "AF_evals_per_s": world * cands_step * k_steps / (ms_m / 1e3),
                             "steps": k_steps}
            del e2
        modes_info["note"] = ("fp16 / tf32: 11-bit significand operands, fp32 accumulate (logits within 1e-3 of the fp32 "
                              "reference, north_star's bar); bf16x3: split-bf16 3-pass, fp32-class (7e-6) -- the default "
                              "mlp_mode of the drop-in NeuralAF / MetaBO / BatchRecorder classes")

    ppo_info = None
    upd_info = None
    train_info = None
    if not args.no_ppo:
        ppo_info = ppo_rollout_rate(dev, world, rank)
        if rank == 0:
            upd_info = ppo_update_rate(dev)
        train_info = [ppo_train_rate(dev, world, rank, None, 4, "C3 reference batch: train_metabo_branin.py (1200 steps = 40 "
                                     "episodes over all ranks, minibatch 60, 4 epochs)"),
                      ppo_train_rate(dev, world, rank, 1024, 1, "C3 at 1024 episodes per GPU (weak scaling of the batch, "
                                     "minibatch 60, 1 epoch per iteration to bound the bench time)")]
        if world > 1:       # data-parallel weak scaling proper: the global minibatch grows with the ranks, every rank keeps 60 transitions per step
            train_info.append(ppo_train_rate(dev, world, rank, 1024, 1, "C3 at 1024 episodes per GPU, minibatch 60 PER GPU (global "
                                             "minibatch %d: weak scaling of batch and minibatch), 1 epoch per iteration" % (60 * world),
                                             minibatch=60 * world))

    cpu_b = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        reps, cpu_b = cpu_arm(4, n, reps=3)        # 1 warm-up + 2 timed repetitions of 4 episode-rounds per core
        cpu_b["value"] = sum(c for c, _ in reps[1:]) / sum(w for _, w in reps[1:])

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "AF evals/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 GP + %s MLP" % args.mlp if args.gp == "fp64" else "f32 GP + %s MLP" % args.mlp,
                "data": "synthetic (U[0,1] observations of Hartmann-3 variants; %s)" % wdesc,
                "config": workload_config(args), "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "AF evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e / args.steps},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu_b, "modes": modes_info,
                "ppo_rollout": ppo_info, "ppo_update": upd_info, "ppo_train": train_info}
        eng.policy.check()            # raises if a tensor-core kernel timed out or the fp16 mode clamped an activation
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
