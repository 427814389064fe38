"""Minimal stand-in for `linear_operator` (TEST INFRASTRUCTURE ONLY; see oracle/shim/README.md)."""
from . import operators  # noqa: F401
