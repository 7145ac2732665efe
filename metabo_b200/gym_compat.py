"""The slice of the gym 0.12 API the MetaBO scripts use (spaces, register/make, registry), so the
drop-in runs where gym is not installed.  If a real `gym` is importable it is used instead.

Reference usage: train_metabo_branin.py:90-96 (register), batchrecorder.py:42 (gym.make),
evaluate.py:41-43,94-99 (env.spec.id, env.spec._kwargs, registry.env_specs)."""
import importlib

import numpy as np

try:  # pragma: no cover - not installed in the build image
    import gym as _gym
    from gym import spaces  # noqa: F401
    from gym.envs.registration import register, registry  # noqa: F401
    make = _gym.make
    Env = _gym.Env
    HAVE_GYM = True
except Exception:
    HAVE_GYM = False

    class Env:
        spec = None

        @property
        def unwrapped(self):
            return self

        def close(self):
            pass

    class _Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    class _Discrete:
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()

        def contains(self, x):
            x = int(np.asarray(x))
            return 0 <= x < self.n

    class _Spaces:
        Box = _Box
        Discrete = _Discrete

    spaces = _Spaces()

    class EnvSpec:
        def __init__(self, id, entry_point, max_episode_steps=None, reward_threshold=None, kwargs=None):
            self.id, self.entry_point = id, entry_point
            self.max_episode_steps, self.reward_threshold = max_episode_steps, reward_threshold
            self._kwargs = {} if kwargs is None else kwargs

    class _Registry:
        def __init__(self):
            self.env_specs = {}

    registry = _Registry()

    def register(id, entry_point=None, max_episode_steps=None, reward_threshold=None, kwargs=None, **_):
        registry.env_specs[id] = EnvSpec(id, entry_point, max_episode_steps, reward_threshold, kwargs)

    def make(env_id):
        spec = registry.env_specs[env_id]
        ep = spec.entry_point
        # the reference registers "metabo.environment.metabo_gym:MetaBO"; route it to the drop-in
        if isinstance(ep, str):
            mod, cls = ep.split(":")
            if mod.startswith("metabo."):
                mod = "metabo_b200." + mod[len("metabo."):]
            ep = getattr(importlib.import_module(mod), cls)
        env = ep(**spec._kwargs)
        env.spec = spec
        return env
