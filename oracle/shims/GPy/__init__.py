"""ORACLE shim: the slice of the GPy API that /root/reference touches
(metabo_gym.py:556-593,612,740).  Numerics live in oracle/gp_oracle.py."""
from . import kern, core, models  # noqa: F401
