"""Locate and import the UNMODIFIED reference ``gstools`` for tests (test infrastructure).

Search order for the package tree:
  1. ``/root/reference/src``      -- the build container (read-only mount)
  2. ``<repo>/baseline/_ref``     -- the git-ignored copy that travels to the GPU box
                                    (made by ``__graft_entry__.build()``)

The stand-in packages in ``tools/refstubs`` (gstools_cython bound to the CPU
oracle, emcee, hankel, meshio, pyevtk) go on ``sys.path`` first.  GPU-marked tests
never require the reference at run time: they fall back to ``tests/golden``.
"""

from __future__ import annotations

import importlib
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(REPO, "tools", "refstubs")
_CANDIDATES = ["/root/reference/src", os.path.join(REPO, "baseline", "_ref")]


def reference_root():
    for cand in _CANDIDATES:
        if os.path.isfile(os.path.join(cand, "gstools", "__init__.py")):
            return cand
    return None


def have_reference() -> bool:
    return reference_root() is not None


def import_gstools():
    """Import and return the reference ``gstools`` (or raise ImportError)."""
    root = reference_root()
    if root is None:
        raise ImportError("reference gstools not found (no /root/reference, no baseline/_ref)")
    for p in (root, STUBS, REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    return importlib.import_module("gstools")
