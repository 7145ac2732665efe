"""ORACLE shim: import-only stand-in for matplotlib."""
