"""GPU tests of the drop-in layer (NeuralAF, MetaBO env, VecMetaBO, BatchRecorder, PPO, evaluate)
against the oracle and the vectors minted from the reference."""
import json
import os
import pickle as pkl

import numpy as np
import pytest
import torch

from oracle import metabo_oracle as O

pytestmark = pytest.mark.gpu

BRA_SPEC = {
    "env_id": "MetaBO-BRA-v0", "D": 2, "f_type": "BRA-var", "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1},
    "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": 30, "n_init_samples": 0,
    "pass_X_to_pi": False, "kernel_lengthscale": [0.235, 0.578], "kernel_variance": 2.0, "noise_variance": 8.9e-16,
    "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 1000, "N_LS": 1000, "k": 5,
    "reward_transformation": "none"}


def _policy(golden_dir, wfile, deterministic=True, mode="bf16x3", M=966, F=6):
    from metabo_b200 import gym_compat as gym
    from metabo_b200.policies.policies import NeuralAF
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, wfile))
    opts = dict(pw.options, mlp_mode=mode)
    obs = gym.spaces.Box(low=-1e5, high=1e5, shape=(M, F), dtype=np.float32)
    pi = NeuralAF(obs, gym.spaces.Discrete(M), deterministic=deterministic, options=opts)
    pi.load_state_dict({k: v.clone() for k, v in pw.sd.items()})
    return pi, pw


@pytest.mark.parametrize("wfile,F,M", [("weights_bra_1635.npz", 6, 966), ("weights_gps_1161.npz", 6, 2000)])
def test_neural_af_api_matches_reference_semantics(golden_dir, wfile, F, M):
    pi, pw = _policy(golden_dir, wfile, M=M, F=F)
    rng = np.random.RandomState(0)
    states = rng.randn(3, M, F).astype(np.float32)
    states[:, :, -2] = 7.0
    states[:, :, -1] = 30.0
    lo, vo = O.policy_forward(pw, states)
    with torch.no_grad():
        lg, v = pi.forward(torch.from_numpy(states))          # CPU tensors in, kernel path, CPU tensors out
    assert lg.device.type == "cpu" and lg.shape == (3, M) and v.shape == (3,)
    scale = float(lo.abs().max())
    assert float((lg - lo).abs().max()) <= 2e-5 * scale + 1e-5
    np.testing.assert_allclose(v.numpy(), vo.numpy(), rtol=1e-5, atol=1e-5)
    af = pi.af(states[0])
    assert af.shape == (M,) and np.allclose(af, lg[0].numpy())
    a, val = pi.act(torch.from_numpy(states[0]))
    assert a.dim() == 0 and val.dim() == 0
    assert float(lo[0, int(a)]) >= float(lo[0].max()) - 2e-5 * scale - 1e-5
    acts = torch.tensor([1.0, 5.0, 9.0])                      # float actions, as ppo.py:132 passes them
    with torch.no_grad():
        pi.set_requires_grad(False)
        v2, lp, en = pi.predict_vals_logps_ents(torch.from_numpy(states), acts)
    lpo, eno = O.categorical_logp_entropy(lo.double(), acts.numpy())
    np.testing.assert_allclose(lp.numpy(), lpo.numpy(), rtol=1e-4, atol=2e-4 * scale)
    np.testing.assert_allclose(en.numpy(), eno.numpy(), rtol=1e-3, atol=1e-3)
    # autograd path (PPO update) on the GPU agrees with the kernel path
    pi.set_requires_grad(True)
    pi_cuda = pi.to("cuda")
    lg2, _ = pi_cuda.forward(torch.from_numpy(states).cuda())
    assert lg2.requires_grad
    assert float((lg2.detach().cpu() - lo).abs().max()) <= 1e-4 * scale + 1e-5


@pytest.mark.parametrize("mode,tol", [("fp16", 2e-3), ("tf32", 2e-3)])
def test_neural_af_forward_fast_modes_on_materialised_states(golden_dir, mode, tol):
    """NeuralAF.forward on materialised states (the dense feature source of the tensor-core kernel) in the two
    throughput modes: logits within 2e-3 of the fp32 oracle on these synthetic U[0,1] states (the 1e-3 bar is stated
    and tested on the reference's own states in test_af_gpu.py), argmax the oracle's unless near-tied."""
    pi, pw = _policy(golden_dir, "weights_hm3_1988.npz", mode=mode, M=1733, F=7)
    rng = np.random.RandomState(3)
    states = rng.rand(4, 1733, 7).astype(np.float32)
    states[:, :, -2] = 11.0
    states[:, :, -1] = 30.0
    lo, _ = O.policy_forward(pw, states)
    with torch.no_grad():
        lg, _ = pi.forward(torch.from_numpy(states))
    scale = float(lo.abs().max())
    assert float((lg - lo).abs().max()) <= tol * scale
    top = lo.max(dim=1).values
    picked = lo.gather(1, lg.argmax(dim=1, keepdim=True))[:, 0]
    assert bool((picked >= top - 2 * tol * scale).all())
    pi.kernel_handles()[0].check()


def test_single_env_deterministic_branin_episode(golden_dir):
    """evaluate_metabo_branin.py settings, seed 100: the first state is a pure function of the
    weights (empty GP) and must equal the reference's; the free-running episode must reach the
    optimum like the reference did (its final simple regret: 3.8e-6)."""
    from metabo_b200.environment.metabo_gym import MetaBO
    ep = np.load(os.path.join(golden_dir, "episode_bra.npz"))
    pi, pw = _policy(golden_dir, "weights_bra_1635.npz")
    env = MetaBO(**BRA_SPEC)
    env.set_af_functions(pi.af)
    env.seed(100)
    state = env.reset()
    assert state.shape == (966, 6) and state.dtype == np.float32
    assert env.observation_space.contains(state)
    ref_state = ep["t00_state"]
    # multistart part of xi_t is identical; the 5 local maxima may permute/tie (flat AF on the empty GP)
    assert np.array_equal(state[5:], ref_state[5:])
    assert abs(env.y_max - float(ep["y_max"])) < 1e-12        # same objective drawn from RandomState(100)
    rewards = []
    for t in range(30):
        action, value = pi.act(torch.from_numpy(state))
        assert env.action_space.contains(int(action))
        state, r, done, info = env.step(action.numpy())
        rewards.append(r)
        assert env.X.shape == (t + 1, 2) and env.Y.shape == (t + 1, 1)
    assert done and env.is_terminal()
    assert np.all(np.diff(rewards) <= 1e-12)                  # simple regret is non-increasing
    assert rewards[-1] < 1e-3, rewards[-1]
    m, s = env.eval_gp(env.X)                                 # GP interpolates its data (sigma ~ 1e-4 floor)
    assert np.max(np.abs(m - env.Y[:, 0])) < 1e-3 and np.max(s) < 1e-3


def test_vec_env_lanes_are_batch_invariant(golden_dir):
    """Episodes are independent units: lane b of a B=4 batch is bitwise the episode a B=1 batch
    with the same seed plays (this is what makes sharding over GPUs collective-free)."""
    from metabo_b200.environment.vec_env import VecMetaBO
    pi, _ = _policy(golden_dir, "weights_bra_1635.npz")
    spec = dict(BRA_SPEC, T=6)
    seeds = [100, 101, 102, 103]

    def run(lane_seeds):
        env = VecMetaBO(len(lane_seeds), lane_seeds, **spec)
        env.set_policy(pi)
        env.reset()
        rs, acts = [], []
        for _ in range(spec["T"]):
            acts.append(env.cur["action"].clone())
            r, done = env.step()
            rs.append(r.clone())
        torch.cuda.synchronize()
        return torch.stack(rs, 1).cpu().numpy(), torch.stack(acts, 1).cpu().numpy(), env.engine.X.cpu().numpy()

    r4, a4, x4 = run(seeds)
    for b, s in enumerate(seeds):
        r1, a1, x1 = run([s])
        assert np.array_equal(r4[b], r1[0]) and np.array_equal(a4[b], a1[0]) and np.array_equal(x4[b], x1[0])


def test_batchrecorder_and_ppo_iteration(golden_dir, tmp_path):
    from metabo_b200 import gym_compat as gym
    from metabo_b200.policies.policies import NeuralAF
    from metabo_b200.ppo.ppo import PPO
    spec = dict(BRA_SPEC, T=5, N_MS=100, N_LS=50, k=3, reward_transformation="neg_log10",
                f_opts={"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50})
    gym.register(id="MetaBO-TEST-v0", entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=5, kwargs=spec)
    params = {"batch_size": 40, "max_steps": 80, "minibatch_size": 10, "n_epochs": 2, "lr": 1e-3, "epsilon": 0.15,
              "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98, "loss_type": "GAElam",
              "normalize_advs": True, "n_workers": 4, "env_id": "MetaBO-TEST-v0", "seed": 0, "env_seeds": [0, 1, 2, 3],
              "policy_options": {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2,
                                 "T_idx": -1, "arch_spec_value": [200] * 4}}

    def policy_fn(osp, asp, det):
        return NeuralAF(osp, asp, det, options=params["policy_options"])

    logpath = str(tmp_path / "log")
    ppo = PPO(policy_fn=policy_fn, params=params, logpath=logpath, save_interval=1)
    w0 = ppo.pi.policy_net.fc[0].weight.detach().clone()
    ppo.train()
    br = ppo.batch_recorder
    assert len(br) == 40 and br.is_full()
    mem = br.memory
    assert len(mem) == 40 and [m.new for m in mem[:6]] == [1, 0, 0, 0, 0, 1]     # worker-major, episode-major
    assert mem[0].state.shape == (103, 6) and mem[0].state.dtype == np.float32
    # GAE on the device == reference formula on the host
    for lane in range(2):
        tr = mem[lane * 5:(lane + 1) * 5]
        a_o, r_o = O.gae([t.reward for t in tr], [float(t.value) for t in tr], [t.new for t in tr], 1,
                         float(tr[-1].value), 0.98, 0.98)
        np.testing.assert_allclose([t.adv for t in tr], a_o, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose([t.tdlamret for t in tr], r_o, rtol=1e-5, atol=1e-5)
    stats = br.get_batch_stats()
    assert stats["n_new"] == 8 and stats["avg_ep_len"] == 5
    assert not torch.equal(w0, ppo.pi.policy_net.fc[0].weight.detach())          # the update moved the weights
    for f in ("000_overview.txt", "log", "stats_0", "stats_1", "params_1", "weights_0", "weights_1"):
        assert os.path.exists(os.path.join(logpath, f)), f
    st = pkl.load(open(os.path.join(logpath, "stats_1"), "rb"))
    assert st["n_timesteps"] == 80 and st["n_optsteps"] == 2 * 2 * 4 and "sps" in st["batch_stats"]
    sd = torch.load(os.path.join(logpath, "weights_1"), weights_only=False)
    assert set(sd) == set(ppo.pi.state_dict())
    from metabo_b200.ppo.util import get_best_iter_idx, get_last_iter_idx
    assert get_last_iter_idx(logpath) == 1 and get_best_iter_idx(logpath) in (0, 1)


def test_eval_experiment_with_shipped_checkpoint_layout(golden_dir, tmp_path):
    """evaluate.py flow on a checkpoint directory laid out like metabo/iclr2020/**."""
    from metabo_b200 import gym_compat as gym
    from metabo_b200.eval.evaluate import eval_experiment
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, "weights_hm3_1988.npz"))
    logpath = tmp_path / "MetaBO-HM3-v0"
    os.makedirs(logpath)
    pkl.dump({"policy_options": pw.options}, open(logpath / "params_1988", "wb"))
    torch.save({k: v for k, v in pw.sd.items()}, logpath / "weights_1988")
    pkl.dump({"avg_step_rews": [0.0]}, open(logpath / "stats_1988", "wb"))
    spec = {"env_id": "MetaBO-HM3-v0", "D": 3, "f_type": "HM3-var", "f_opts": {"bound_translation": 0.1, "bound_scaling": 0.1},
            "features": ["posterior_mean", "posterior_std", "timestep", "budget", "x"], "T": 30, "n_init_samples": 0,
            "pass_X_to_pi": False, "kernel_lengthscale": [0.716, 0.298, 0.186], "kernel_variance": 0.83,
            "noise_variance": 1.688e-11, "use_prior_mean_function": False, "local_af_opt": True, "N_MS": 2000,
            "N_LS": 2000, "k": 5, "reward_transformation": "none"}
    gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=30, kwargs=spec)
    res = eval_experiment({"env_id": spec["env_id"], "env_seed_offset": 100, "policy": "MetaBO", "logpath": str(logpath),
                           "load_iter": 1988, "deterministic": True, "policy_specs": {}, "savepath": str(tmp_path / "eval"),
                           "n_workers": 4, "n_episodes": 8, "T": 30})
    r = np.array(res.rewards).reshape(8, 30)
    assert np.all(np.diff(r, axis=1) <= 1e-12)
    assert np.median(r[:, -1]) < 5e-2                         # reference: median simple regret ~1e-3 after 30 steps
    assert os.path.exists(tmp_path / "eval" / "result_metabo_iter_1988")


@pytest.mark.parametrize("mode", ["tf32", "fp16"])
def test_cuda_graph_replay_equals_eager_rollouts(golden_dir, mode):
    """The episode loop captured as CUDA graphs (one per BO step index) must reproduce the eager rollouts bit for
    bit over several consecutive batches (new objectives every episode, stochastic policy: fresh Gumbel noise on
    every replay through the device-resident step counter)."""
    from metabo_b200 import gym_compat as gym
    from metabo_b200.policies.policies import NeuralAF
    from metabo_b200.ppo.batchrecorder import BatchRecorder
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, "weights_bra_1635.npz"))
    spec = dict(BRA_SPEC, T=6, N_MS=400, N_LS=200, f_opts={"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50},
                reward_transformation="neg_log10", env_id="MetaBO-GRAPH-%s-v0" % mode)
    gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=6, kwargs=spec)
    opts = dict(pw.options, mlp_mode=mode)

    def run(use_graphs):
        fn = lambda o, a, d: NeuralAF(o, a, d, options=opts)
        br = BatchRecorder(size=6 * 12, env_id=spec["env_id"], env_seeds=[3, 4, 5], policy_fn=fn, n_workers=3,
                           deterministic=False, store_states=True, mlp_mode=mode, use_graphs=use_graphs)
        pi = fn(br.observation_space, br.action_space, False)
        pi.load_state_dict({k: v.clone() for k, v in pw.sd.items()})
        br.set_worker_weights(pi)
        outs = []
        for _ in range(4):                       # batch 1 eager warm-up, batch 2 capture, batches 3-4 replay
            br.record_batch(0.98, 0.98)
            outs.append({k: v.clone() for k, v in br.buf.items()})
        return outs, br

    eager, _ = run(False)
    graph, br = run(True)
    assert br.env.use_graphs and any("graph" in e for e in br.env._graphs.values())
    for a, b in zip(eager, graph):
        for k in ("actions", "rewards", "values", "states", "advs"):
            assert torch.equal(a[k], b[k]), k
    assert not torch.equal(eager[2]["actions"], eager[3]["actions"])      # different noise / objectives per batch


def test_graph_captured_ppo_update_equals_eager(golden_dir, tmp_path, monkeypatch):
    """PPO.optimize_on_batch with the optimiser step captured in a CUDA graph follows the eager trajectory
    (same minibatch order, same Adam state): weights after two iterations agree to fp32 round-off."""
    import random
    from metabo_b200 import gym_compat as gym
    from metabo_b200.policies.policies import NeuralAF
    from metabo_b200.ppo.ppo import PPO
    spec = dict(BRA_SPEC, T=5, N_MS=100, N_LS=50, k=3, reward_transformation="neg_log10", env_id="MetaBO-UPD-v0",
                f_opts={"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50})
    gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=5, kwargs=spec)

    def run(update_graph, name, **extra):
        params = {"batch_size": 80, "max_steps": 160, "minibatch_size": 10, "n_epochs": 2, "lr": 1e-3, "epsilon": 0.15,
                  "value_coeff": 1.0, "ent_coeff": 0.01, "gamma": 0.98, "lambda": 0.98, "loss_type": "GAElam",
                  "normalize_advs": True, "n_workers": 4, "env_id": spec["env_id"], "seed": 0, "env_seeds": [0, 1, 2, 3],
                  "update_graph": update_graph,
                  "policy_options": {"activations": "relu", "arch_spec": [200] * 4, "use_value_network": True, "t_idx": -2,
                                     "T_idx": -1, "arch_spec_value": [200] * 4}}
        params.update(extra)
        ppo = PPO(policy_fn=lambda o, a, d: NeuralAF(o, a, d, options=params["policy_options"]), params=params,
                  logpath=str(tmp_path / name), save_interval=10, verbose=True)
        ppo.train()
        return torch.cat([p.detach().reshape(-1) for p in ppo.pi.parameters()]), ppo

    w_e, _ = run(False, "eager")
    # same kernels, same order: per-step graphs and the one-graph-per-epoch path (minibatches gathered by torch ops)
    # follow the eager trajectory exactly
    w_s, ppo_s = run(True, "step_graph", epoch_graph=False)
    assert ppo_s._upd is not None and ppo_s.stats["n_optsteps"] == 2 * 2 * 8
    assert float((w_e - w_s).abs().max()) <= 1e-5 * float(w_e.abs().max())
    w_g, ppo = run(True, "epoch_graph", fused_minibatch=False)
    assert ppo._ep is not None and ppo.stats["n_optsteps"] == 2 * 2 * 8
    assert float((w_e - w_g).abs().max()) <= 1e-5 * float(w_e.abs().max())
    # default: minibatch gather + advantage normalisation in mbo_ppo_prepare_minibatch (double sums instead of torch's fp32
    # reductions: last-bit differences in the advantages, which Adam turns into +-lr steps wherever a gradient entry is at
    # noise level -- this freshly initialised policy): logged losses agree, weights within Adam's step bound
    w_f, ppo_f = run(True, "epoch_fused")
    assert ppo_f._ep is not None
    log = lambda p: [float(x.split("=")[1]) for line in p.iter_log_str.splitlines() if "loss_ppo" in line for x in line.split(",")[1:]]
    np.testing.assert_allclose(log(ppo_f), log(ppo), rtol=1e-3)
    assert float((w_f - w_g).abs().max()) <= 2.5 * 1e-3 * 24


def test_graph_replay_values_follow_weight_updates(golden_dir):
    """the rollout graphs capture the value net's device pointers by value; `set_worker_weights` with
    NEW weights before every batch (what PPO.train does once per iteration) must be seen by the replayed graphs for
    at least 3 iterations after the capture -- values of the graph path == eager path, and they change with the weights."""
    from metabo_b200 import gym_compat as gym
    from metabo_b200.policies.policies import NeuralAF
    from metabo_b200.ppo.batchrecorder import BatchRecorder
    pw = O.PolicyWeights.from_npz(os.path.join(golden_dir, "weights_bra_1635.npz"))
    spec = dict(BRA_SPEC, T=4, N_MS=100, N_LS=50, k=3, f_opts={"bound_translation": 0.1, "bound_scaling": 0.1, "M": 50},
                reward_transformation="neg_log10", env_id="MetaBO-VALNET-v0")
    gym.register(id=spec["env_id"], entry_point="metabo.environment.metabo_gym:MetaBO", max_episode_steps=4, kwargs=spec)
    opts = dict(pw.options, mlp_mode="fp16")

    def run(use_graphs):
        fn = lambda o, a, d: NeuralAF(o, a, d, options=opts)
        br = BatchRecorder(size=4 * 6, env_id=spec["env_id"], env_seeds=[3, 4, 5], policy_fn=fn, n_workers=3,
                           deterministic=False, store_states=False, mlp_mode="fp16", use_graphs=use_graphs)
        pi = fn(br.observation_space, br.action_space, False)
        pi.load_state_dict({k: v.clone() for k, v in pw.sd.items()})
        h_val = None
        vals = []
        for it in range(6):
            with torch.no_grad():                      # a different value head (and policy) every iteration
                pi.value_net.fc[4].bias.add_(0.25)
                pi.policy_net.fc[4].bias.add_(0.01)
            br.set_worker_weights(pi)
            if h_val is None:
                h_val = br.env.engine.value
            assert br.env.engine.value is h_val        # ONE handle for the life of the engine
            br.record_batch(0.98, 0.98)
            vals.append(br.buf["values"].clone())
        return vals, br

    eager, _ = run(False)
    graph, br = run(True)
    assert any("graph" in e for e in br.env._graphs.values())
    for it, (a, b) in enumerate(zip(eager, graph)):
        assert torch.equal(a, b), it
    for it in range(1, 6):
        assert float((eager[it] - eager[it - 1]).mean()) == pytest.approx(0.25, abs=1e-3)


def test_gp_family_with_local_af_optimisation(golden_dir):
    """f_type "GP" with local_af_opt=True (evaluate_metabo_gps.py-style continuous domain): the
    extrema come from the uniform grid with ceil(1000^(1/D)) points per dimension (metabo_gym.py:254-258)."""
    from metabo_b200.environment.vec_env import VecMetaBO
    from metabo_b200.environment.functions import FunctionSampler
    from metabo_b200.engine import uniform_grid
    pi, pw = _policy(golden_dir, "weights_gps_1161.npz", mode="fp32", M=2000, F=6)
    D = 3
    spec = {"env_id": "MetaBO-GPLOC-v0", "D": D, "f_type": "GP",
            "f_opts": {"kernel": "Matern52", "lengthscale_low": 0.05, "lengthscale_high": 0.5, "noise_var_low": 0.1,
                       "noise_var_high": 0.1, "signal_var_low": 1.0, "signal_var_high": 1.0, "min_regret": 1e-6},
            "features": ["posterior_mean", "posterior_std", "incumbent", "timestep_perc", "timestep", "budget"],
            "T": 4, "n_init_samples": 0, "pass_X_to_pi": False, "kernel": "Matern52", "kernel_lengthscale": None,
            "kernel_variance": None, "noise_variance": None, "use_prior_mean_function": True, "local_af_opt": True,
            "N_MS": 500, "N_LS": 200, "k": 3, "reward_transformation": "neg_log10"}
    env = VecMetaBO(2, [11, 12], mlp_mode="fp32", **spec)
    env.set_policy(pi)
    env.reset(deterministic=True)
    # y_max / y_min == max / min of the drawn sample over the 10^3 uniform grid
    grid = torch.as_tensor(uniform_grid(D, 10), device=env.dev)
    y = env.fb.raw(grid[None].expand(2, -1, -1))
    assert torch.equal(env.fb.y_max, y.max(dim=1).values) and torch.equal(env.fb.y_min, y.min(dim=1).values)
    for _ in range(4):
        r, done = env.step(deterministic=True)
    torch.cuda.synchronize()
    assert done and bool(torch.isfinite(r).all())


@pytest.mark.parametrize("family,D", [("bra", 2), ("hm3", 3)])
@pytest.mark.parametrize("reward", ["none", "neg_linear", "neg_log10"])
def test_fused_env_step_matches_torch_objectives(family, D, reward):
    """`mbo_env_step` (one launch: objective value, append, incumbent, reward, next state's scalars) against the torch
    expressions it replaces (metabo_b200/objectives.py = objectives.py:26-80 / :165-219, metabo_gym.py:679-704, 723-730):
    the objective value to a few 1e-16 (one rounding per operation in the same order; only the libm calls and the order of two short
    sums can differ), everything derived from it likewise."""
    from metabo_b200 import ops, objectives as obj, _lib as L
    dev = torch.device("cuda:0")
    B, T, t_idx = 333, 30, 7
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, D, generator=g, dtype=torch.float64).to(dev)
    trans = ((torch.rand(B, D, generator=g, dtype=torch.float64) - 0.5) * 2.2).to(dev)      # beyond the clip ranges too
    scale = (0.9 + 0.2 * torch.rand(B, generator=g, dtype=torch.float64)).to(dev)
    fvar, fmax = (obj.bra_var, obj.bra_max) if family == "bra" else (obj.hm3_var, obj.hm3_max)
    y_ref = fvar(x, trans, scale)
    y_max, _ = fmax(trans, scale)
    y_min = y_max - 5.0
    best0 = torch.full((B,), -float("inf"), dtype=torch.float64, device=dev)
    best0[::3] = y_ref[::3] + 0.01                      # lanes whose incumbent stays
    X = torch.zeros(B, T, D, dtype=torch.float64, device=dev); Y = torch.zeros(B, T, dtype=torch.float64, device=dev)
    n = torch.zeros(B, dtype=torch.int32, device=dev)
    best = best0.clone()
    epi = torch.zeros(B, 4, dtype=torch.float32, device=dev)
    rew = torch.empty(B, dtype=torch.float64, device=dev)
    ops.env_step(L.ENV_BRANIN if family == "bra" else L.ENV_HARTMANN3, t_idx, x, trans, scale, y_max.contiguous(), y_min.contiguous(),
                 X, Y, n, best, L.REWARDS[reward], rew, epi=epi, t_next=t_idx + 1, T_budget=T, T_training=None)
    torch.cuda.synchronize()
    y = Y[:, t_idx]
    assert float(((y - y_ref).abs() / y_ref.abs().clamp(min=1.0)).max()) < 1e-15      # values are O(1) sums of O(1) terms
    others = torch.ones(T, dtype=torch.bool, device=dev); others[t_idx] = False
    assert torch.equal(X[:, t_idx], x) and float(X[:, others].abs().max()) == 0.0 and float(Y[:, others].abs().max()) == 0.0
    assert torch.equal(n, torch.full_like(n, t_idx + 1))
    best_ref = torch.maximum(best0, y)
    assert torch.equal(best, best_ref)
    regret = y_max - best_ref
    rew_ref = {"none": regret, "neg_linear": -regret, "neg_log10": -torch.log10(torch.clamp(regret, min=1e-20))}[reward]
    assert float((rew - rew_ref).abs().max()) <= 1e-14 * float(rew_ref.abs().max())
    assert torch.equal(epi[:, 0], best_ref.float())
    assert torch.equal(epi[:, 1], torch.full((B,), float(t_idx + 1), device=dev) / torch.full((B,), float(T), device=dev))
    assert float((epi[:, 2] - (t_idx + 1)).abs().max()) == 0.0 and float((epi[:, 3] - T).abs().max()) == 0.0
