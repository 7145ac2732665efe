"""ORACLE shim: slice of gym 0.12.5 used by the reference (Env, spaces, make, registry)."""
import importlib
from . import spaces  # noqa: F401
from .envs import registration as _reg
from .envs.registration import register, registry  # noqa: F401


class Env:
    spec = None

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


def make(env_id):
    spec = registry.env_specs[env_id]
    mod_name, cls_name = spec.entry_point.split(":")
    cls = getattr(importlib.import_module(mod_name), cls_name)
    env = cls(**spec._kwargs)
    env.spec = spec
    return env
