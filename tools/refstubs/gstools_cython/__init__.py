"""Offline stand-in for the external ``gstools-cython`` package (test infrastructure)."""
from . import field, krige, variogram  # noqa: F401

__version__ = "0.0.0+oracle"
