"""ORACLE (test infrastructure): import the reference's own `metabo` package, unmodified,
from /root/reference in the authoring container.

/root/reference does not exist on the GPU box; callers must check `reference_available()`
and skip.  Used only by oracle/make_golden.py and by `-m "not gpu"` tests that validate the
oracle restatements against the reference classes.

Compat patches applied (numpy 2 removed these; the reference was written for numpy 1.16):
  np.int           metabo/environment/metabo_gym.py:75
  np.asscalar      metabo/environment/metabo_gym.py:688
  ndarray.ptp      metabo/environment/util.py:37,42  (rebinds the two scale_* helpers with
                   np.ptp versions before metabo_gym imports them)
"""
import os
import sys

REFERENCE_ROOT = "/root/reference"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "metabo"))


def load_reference():
    """Returns the imported reference `metabo` package (with shims + numpy patches)."""
    if not reference_available():
        raise RuntimeError("reference not mounted at " + REFERENCE_ROOT)
    import numpy as np
    for p in (_REPO, _SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "asscalar"):
        np.asscalar = lambda a: np.asarray(a).item()
    import metabo.environment.util as util

    def scale_from_unit_square_to_domain(X, domain):
        return X * np.ptp(domain, axis=1) + domain[:, 0]

    def scale_from_domain_to_unit_square(X, domain):
        return (X - domain[:, 0]) / np.ptp(domain, axis=1)

    util.scale_from_unit_square_to_domain = scale_from_unit_square_to_domain
    util.scale_from_domain_to_unit_square = scale_from_domain_to_unit_square
    import metabo
    import metabo.environment.metabo_gym  # noqa: F401
    import metabo.policies.policies  # noqa: F401
    return metabo


def shipped(*parts):
    return os.path.join(REFERENCE_ROOT, "metabo", "iclr2020", *parts)
