"""ORACLE shim for sobol_seq 0.1.2 -> oracle/sobol.py restatement."""
from oracle.sobol import i4_sobol_generate  # noqa: F401
