"""BatchRecorder drop-in (metabo/ppo/batchrecorder.py:199-342).

The reference forks `n_workers` EnvRunner processes, each playing `size / n_workers` transitions
sequentially on the CPU and pickling every transition (including its state[M,F]) back to the
learner.  Here every episode of the batch is a lane of one `VecMetaBO` on the GPU:

  * worker i keeps its seed semantics: its `RandomState(env_seeds[i])` is consumed once per episode
    in the same order, so the objectives of "worker i, episode j" are the reference's;
  * rollouts stay device-resident (actions, rewards, values, optionally states), GAE(lambda)
    (batchrecorder.py:117-137) is evaluated on the device;
  * `memory` (list of Transition, worker-major / time-minor like :272) is materialised lazily for
    callers that want the reference's host-side view (evaluate.py:151-152, plot_results.py:45-51).

Restriction: `size / n_workers` must be a multiple of the (fixed) horizon T, which holds for every
shipped config; then every worker batch ends on an episode boundary and the bootstrap term of
:111-113 vanishes (nonterminal = 0).
"""
import os
import random
import time

import numpy as np
import torch

from .. import gym_compat as gym
from ..environment.vec_env import VecMetaBO


class Transition:
    """mutable record with the reference's field order (namedlist in batchrecorder.py:35)"""
    __slots__ = ("state", "action", "reward", "value", "new", "tdlamret", "adv")
    _fields = __slots__

    def __init__(self, state, action, reward, value, new, tdlamret, adv):
        self.state, self.action, self.reward, self.value = state, action, reward, value
        self.new, self.tdlamret, self.adv = new, tdlamret, adv

    def __iter__(self):
        return iter((self.state, self.action, self.reward, self.value, self.new, self.tdlamret, self.adv))


def gae_device(rewards, values, gamma, lam):
    """GAE(lambda) over [B, T] episodes that start with new=1 and end terminal
    (batchrecorder.py:117-137 with next_new = 1 at the batch end)."""
    B, T = rewards.shape
    adv = torch.zeros_like(rewards)
    next_adv = torch.zeros(B, dtype=rewards.dtype, device=rewards.device)
    next_value = torch.zeros(B, dtype=rewards.dtype, device=rewards.device)
    nonterminal = 0.0
    for t in reversed(range(T)):
        delta = -values[:, t] + rewards[:, t] + gamma * nonterminal * next_value
        next_adv = delta + lam * gamma * nonterminal * next_adv
        adv[:, t] = next_adv
        next_value = values[:, t]
        nonterminal = 1.0          # within an episode; the step before a `new` transition is terminal
    return adv, adv + values


class BatchRecorder:
    def __init__(self, size, env_id, env_seeds, policy_fn, n_workers, deterministic=False, store_states=True,
                 gp_precision="fp64", mlp_mode=None, device=None, use_graphs=True):
        self.env_id = env_id
        self.deterministic = bool(deterministic)
        assert size > 0
        self.n_workers = n_workers
        self.size = size
        assert len(env_seeds) == n_workers
        self.env_seeds = list(env_seeds)
        self.worker_batch_sizes = [self.size // self.n_workers] * self.n_workers
        assert self.size - sum(self.worker_batch_sizes) == 0, 'All workers shall get assigned the same batch size!'
        spec = gym.registry.env_specs[env_id]
        kwargs = dict(spec._kwargs)
        T = kwargs["T"] if "T" in kwargs else kwargs["T_min"]
        wbs = self.worker_batch_sizes[0]
        if wbs % T != 0:
            raise NotImplementedError("size / n_workers must be a multiple of the horizon T (lock-step lanes)")
        self.T = T
        self.episodes_per_worker = wbs // T
        self.B = self.n_workers * self.episodes_per_worker
        self.worker_rngs = []
        for s in self.env_seeds:
            r = np.random.RandomState()
            r.seed(int(s))
            self.worker_rngs.append(r)
        lane_rngs = [self.worker_rngs[i] for i in range(self.n_workers) for _ in range(self.episodes_per_worker)]
        # policy (built like EnvRunner does, batchrecorder.py:48-49), weights arrive via set_worker_weights
        dummy_obs = gym.spaces.Box(low=-100000.0, high=100000.0, shape=(1, 1), dtype=np.float32)
        self.env = None
        self._kwargs = kwargs
        self._lane_rngs = lane_rngs
        self._opts = dict(gp_precision=gp_precision, device=device)
        self._mlp_mode = mlp_mode
        self.store_states = store_states
        self.use_graphs = use_graphs
        self._build_env(policy_fn)
        self.sample_seed = int(self.env_seeds[-1])     # the reference's workers all inherit env_seeds[-1] (Appendix D)
        self.clear()
        del dummy_obs

    def _build_env(self, policy_fn):
        # observation/action spaces need the env's cardinality -> build the env first
        probe_mode = self._mlp_mode or "bf16x3"
        self.env = VecMetaBO(self.B, None, lane_rngs=self._lane_rngs, mlp_mode=probe_mode, **self._opts, **self._kwargs)
        self.observation_space = gym.spaces.Box(low=-100000.0, high=100000.0,
                                                shape=(self.env.cardinality_xi_t, self.env.n_features), dtype=np.float32)
        self.action_space = gym.spaces.Discrete(self.env.cardinality_xi_t)
        self.pi = policy_fn(self.observation_space, self.action_space, self.deterministic)
        self.pi.set_requires_grad(False)
        from ..policies.policies import NeuralAF
        # a foreign policy (TAF, ...: no kernel form) plays its episodes one at a time through the single-episode MetaBO
        # env and the reference's host-side AF gridding, like one EnvRunner per worker (batchrecorder.py:38-113)
        self.host_policy = not isinstance(self.pi, NeuralAF) and not hasattr(self.pi, "engine_spec")
        self._host_envs = None
        if self.host_policy:
            return
        if hasattr(self.pi, "engine_spec"):
            self.deterministic = True          # closed-form baselines act by argmax (policies.py:191-199)
        self.env.set_policy(self.pi)
        self.env.use_graphs = bool(self.use_graphs) and os.environ.get("MBO_NO_GRAPHS") is None

    def clear(self):
        self.cur_size = self.size
        self.worker_sizes = [0 for _ in range(self.n_workers)]
        self._memory = None
        self.buf = None
        self.reward_sum = 0
        self.n_new = 0
        self.worker_n_news = [0 for _ in range(self.n_workers)]
        self.initial_rewards = []
        self.terminal_rewards = []

    def overview_dict(self):
        return {"size": self.size, "n_workers": self.n_workers, "worker_batch_sizes": self.worker_batch_sizes,
                "env_seeds": self.env_seeds, "deterministic": self.deterministic}

    def set_worker_weights(self, pi):
        now = time.time()
        if not self.host_policy:
            self.pi.load_state_dict(pi.state_dict())
            self.env.sync_policy()
        return time.time() - now

    def _record_batch_host(self, gamma, lam):
        """foreign policies: sequential episodes per worker on the host-driven MetaBO env (GP still on the kernels)"""
        from ..environment.metabo_gym import MetaBO
        now = time.time()
        self.clear()
        B, T, dev = self.B, self.T, self.env.dev
        if self._host_envs is None:
            self._host_envs = []
            for seed in self.env_seeds:
                e = MetaBO(**self._kwargs)
                e.set_af_functions(self.pi.af)
                e.seed(int(seed))
                self._host_envs.append(e)
        actions = torch.zeros((B, T), dtype=torch.int32)
        rewards = torch.zeros((B, T), dtype=torch.float64)
        values = torch.zeros((B, T), dtype=torch.float32)
        states = torch.zeros((B, T, self.env.cardinality_xi_t, self.env.n_features), dtype=torch.float32) if self.store_states else None
        lane = 0
        for w, env in enumerate(self._host_envs):
            for _ in range(self.episodes_per_worker):
                self.pi.reset()
                state = env.reset()
                for t in range(T):
                    st = torch.from_numpy(state.astype(np.float32))
                    with torch.no_grad():
                        a, v = self.pi.act(st, env.X, env.gp) if env.pass_X_to_pi else self.pi.act(st)   # :165-170
                    if states is not None:
                        states[lane, t] = st
                    actions[lane, t], values[lane, t] = int(a), float(v)
                    state, r, done, _ = env.step(int(a))
                    rewards[lane, t] = r
                assert done
                lane += 1
        adv, tdlamret = gae_device(rewards, values.double(), gamma, lam)
        self.buf = dict(states=None if states is None else states.to(dev), actions=actions.to(dev), rewards=rewards.to(dev),
                        values=values.to(dev), advs=adv.float().to(dev), tdlamrets=tdlamret.float().to(dev))
        r_host = rewards.numpy()
        self.reward_sum = float(r_host.sum())
        self.n_new = B
        self.worker_sizes = list(self.worker_batch_sizes)
        self.worker_n_news = [self.episodes_per_worker] * self.n_workers
        self.initial_rewards = list(r_host[:, 0])
        self.terminal_rewards = list(r_host[:, -1])
        return time.time() - now

    def record_batch(self, gamma, lam):
        if self.host_policy:
            return self._record_batch_host(gamma, lam)
        now = time.time()
        self.clear()
        env, B, T = self.env, self.B, self.T
        dev = env.dev
        env.sample_seed = self.sample_seed
        env.want_states = self.store_states
        actions = torch.zeros((B, T), dtype=torch.int32, device=dev)
        rewards = torch.zeros((B, T), dtype=torch.float64, device=dev)
        values = torch.zeros((B, T), dtype=torch.float32, device=dev)
        states = None
        if self.store_states:
            states = torch.zeros((B, T, env.cardinality_xi_t, env.n_features), dtype=torch.float32, device=dev)
        env.reset(deterministic=self.deterministic)
        for t in range(T):
            actions[:, t] = env.cur["action"]
            values[:, t] = env.values()
            if states is not None:
                states[:, t] = env.current_state()
            r, done = env.step(deterministic=self.deterministic)
            rewards[:, t] = r
        assert done
        adv, tdlamret = gae_device(rewards, values.double(), gamma, lam)
        self.buf = dict(states=states, actions=actions, rewards=rewards, values=values, advs=adv.float(),
                        tdlamrets=tdlamret.float())
        torch.cuda.synchronize(dev)
        env.engine.policy.check()
        # host-side statistics (batchrecorder.py:304-316)
        r_host = rewards.cpu().numpy()
        self.reward_sum = float(r_host.sum())
        self.n_new = B
        self.worker_sizes = list(self.worker_batch_sizes)
        self.worker_n_news = [self.episodes_per_worker] * self.n_workers
        self.initial_rewards = list(r_host[:, 0])
        self.terminal_rewards = list(r_host[:, -1])
        return time.time() - now

    # -- reference-style host view ---------------------------------------------------------------
    @property
    def memory(self):
        if self._memory is None:
            if self.buf is None:
                return []
            b = {k: (v.cpu().numpy() if v is not None else None) for k, v in self.buf.items()}
            mem = []
            for lane in range(self.B):              # lanes are worker-major, episodes of a worker in order
                for t in range(self.T):
                    st = b["states"][lane, t] if b["states"] is not None else None
                    mem.append(Transition(st, np.array(int(b["actions"][lane, t])), float(b["rewards"][lane, t]),
                                          np.array(b["values"][lane, t]), int(t == 0),
                                          float(b["tdlamrets"][lane, t]), float(b["advs"][lane, t])))
            self._memory = mem
        return self._memory

    def _order(self, shuffle):
        idx = list(range(len(self)))
        if shuffle:
            random.shuffle(idx)                     # main-process `random`, no re-seeding (batchrecorder.py:322-324)
        return idx

    def _minibatch_bounds(self, minibatch_size):
        n, pos, out = len(self), 0, []
        while pos < n:
            cur = n - pos if pos + 2 * minibatch_size > n else minibatch_size   # :326-330
            out.append((pos, pos + cur))
            pos += cur
        return out

    def iterate(self, minibatch_size, shuffle):
        assert self.is_full()
        idx = self._order(shuffle)
        mem = self.memory
        for a, b in self._minibatch_bounds(minibatch_size):
            yield [mem[i] for i in idx[a:b]]

    def iterate_device(self, minibatch_size, shuffle):
        """Same minibatch partition as `iterate`, as device tensors (no host round trip)."""
        assert self.is_full()
        idx = torch.as_tensor(self._order(shuffle), dtype=torch.int64, device=self.env.dev)
        flat = {k: (v.reshape((self.B * self.T,) + tuple(v.shape[2:])) if v is not None else None)
                for k, v in self.buf.items()}
        for a, b in self._minibatch_bounds(minibatch_size):
            sel = idx[a:b]
            yield {k: (v.index_select(0, sel) if v is not None else None) for k, v in flat.items()}

    def get_batch_stats(self):
        assert self.is_full()
        bs = dict()
        bs["size"] = len(self)
        bs["worker_sizes"] = self.worker_sizes
        bs["avg_step_reward"] = self.reward_sum / len(self)
        bs["avg_initial_reward"] = np.mean(self.initial_rewards)
        bs["avg_terminal_reward"] = np.mean(self.terminal_rewards)
        bs["avg_ep_reward"] = self.reward_sum / self.n_new
        bs["avg_ep_len"] = len(self) / self.n_new
        bs["n_new"] = self.n_new
        bs["worker_n_news"] = self.worker_n_news
        return bs

    def cleanup(self):
        pass

    def is_empty(self):
        return len(self) == 0

    def is_full(self):
        return len(self) == self.cur_size

    def __len__(self):
        return 0 if self.buf is None else self.B * self.T
