class EnvSpec:
    def __init__(self, id, entry_point, max_episode_steps=None, reward_threshold=None, kwargs=None):
        self.id = id
        self.entry_point = entry_point
        self.max_episode_steps = max_episode_steps
        self.reward_threshold = reward_threshold
        self._kwargs = {} if kwargs is None else kwargs


class _Registry:
    def __init__(self):
        self.env_specs = {}


registry = _Registry()


def register(id, entry_point=None, max_episode_steps=None, reward_threshold=None, kwargs=None, **_):
    registry.env_specs[id] = EnvSpec(id, entry_point, max_episode_steps, reward_threshold, kwargs)
