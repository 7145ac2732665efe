from .vec_env import VecMetaBO  # noqa: F401
