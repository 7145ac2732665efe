from . import gp_regression  # noqa: F401
