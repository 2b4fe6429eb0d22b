"""Plug the B200 summator into an UNMODIFIED ``gstools`` as a third backend.

The reference chooses its native backend per call inside two module-level wrappers,
``gstools.field.generator._summate`` / ``_summate_incompr`` (generator.py:42-64), from the
flags in ``gstools.config`` (config.py:8-17).  Both wrappers are looked up as module globals
at call time (generator.py:266, :554), so rebinding them takes effect for existing
``RandMeth`` / ``SRF`` / ``CondSRF`` objects immediately.  :func:`enable`

  * adds ``gstools.config.USE_GSTOOLS_B200`` next to ``USE_GSTOOLS_CORE`` and rebinds the two
    wrappers to versions that consult it first and otherwise fall through to the originals
    (Cython / Rust) -- zero edits to the reference;
  * optionally (``lazy_grid=True``) wraps ``Field.pre_pos`` (field/base.py:254-297) so that, for
    ``mesh_type="structured"`` calls of an SRF/CondSRF driven by RandMeth/IncomprRandMeth/Fourier, the
    mesh is NOT expanded on the host: ``pre_pos`` returns a zero-stride placeholder of the right
    shape ``(dim, n)`` that carries ``(axes, isometrisation matrix)``, and the rebound wrapper
    recognises it and runs the separable structured kernel.  Everything between ``pre_pos`` and
    the wrapper (``RandMeth.__call__``: dtype coercion, sqrt(var/N) scaling, nugget,
    generator.py:261-270) runs unchanged.
"""

from __future__ import annotations

import threading
import weakref

import numpy as np

from . import backend

__all__ = ["enable", "disable", "is_enabled", "LazyGridPos"]

_STATE = {"enabled": False}
_LOCK = threading.Lock()
_LAZY = {}  # address of the placeholder buffer -> (axes, matrix)


class LazyGridPos(np.ndarray):
    """Zero-stride ``(dim, n)`` placeholder for an unexpanded structured mesh.

    Reading it yields zeros; only its shape, its buffer address (registry key) and the attached
    ``axes`` / ``matrix`` matter.  It is handed from the wrapped ``Field.pre_pos`` to the rebound
    ``_summate`` wrappers and never reaches any other consumer.
    """

    def __new__(cls, axes, matrix):
        axes = tuple(np.ascontiguousarray(a, dtype=np.double).reshape(-1) for a in axes)
        n = int(np.prod([a.shape[0] for a in axes])) if axes else 0
        holder = np.zeros(1, dtype=np.double)
        obj = np.lib.stride_tricks.as_strided(holder, shape=(len(axes), n), strides=(0, 0),
                                              writeable=False).view(cls)
        obj.axes = axes
        obj.matrix = None if matrix is None else np.ascontiguousarray(matrix, dtype=np.double)
        obj._holder = holder
        key = holder.__array_interface__["data"][0]
        _LAZY[key] = (obj.axes, obj.matrix)
        weakref.finalize(holder, _LAZY.pop, key, None)
        return obj

    def __array_finalize__(self, obj):
        if obj is not None:
            self.axes = getattr(obj, "axes", None)
            self.matrix = getattr(obj, "matrix", None)
            self._holder = getattr(obj, "_holder", None)


def _lookup_lazy(pos):
    """(axes, matrix) if ``pos`` is (a base-class view of) a :class:`LazyGridPos`."""
    if not _LAZY or not isinstance(pos, np.ndarray) or pos.ndim != 2 or pos.strides != (0, 0):
        return None
    return _LAZY.get(pos.__array_interface__["data"][0])


def is_enabled() -> bool:
    return _STATE["enabled"]


def enable(lazy_grid: bool = True):
    """Route ``RandMeth`` / ``IncomprRandMeth`` summation of ``gstools`` to the B200 backend."""
    import gstools  # the user's (unmodified) installation
    from gstools import config
    from gstools.field import base as fbase
    from gstools.field import generator as gen
    from gstools.tools.geometric import matrix_isometrize

    with _LOCK:
        if not _STATE["enabled"]:
            _STATE.update(orig_summate=gen._summate, orig_summate_incompr=gen._summate_incompr,
                          orig_summate_fourier=gen._summate_fourier,
                          orig_pre_pos=fbase.Field.pre_pos, gen=gen, fbase=fbase, config=config)
        orig_s, orig_si = _STATE["orig_summate"], _STATE["orig_summate_incompr"]
        orig_sf = _STATE["orig_summate_fourier"]
        orig_pre_pos = _STATE["orig_pre_pos"]
        config.USE_GSTOOLS_B200 = True
        config._GSTOOLS_B200_AVAIL = True

        def _summate(cov_samples, z_1, z_2, pos, num_threads=None):
            """A wrapper function for calling the randomization algorithms (B200 first)."""
            if getattr(config, "USE_GSTOOLS_B200", False):
                lazy = _lookup_lazy(pos)
                if lazy is not None:
                    return backend.summate_structured(cov_samples, z_1, z_2, lazy[0],
                                                      lazy[1]).reshape(-1)
                return backend.summate(cov_samples, z_1, z_2, pos, num_threads)
            return orig_s(cov_samples, z_1, z_2, _materialise(pos), num_threads)

        def _summate_incompr(cov_samples, z_1, z_2, pos, num_threads=None):
            """A wrapper function for calling the incompr. randomization algorithms (B200 first)."""
            if getattr(config, "USE_GSTOOLS_B200", False):
                lazy = _lookup_lazy(pos)
                if lazy is not None:
                    out = backend.summate_incompr_structured(cov_samples, z_1, z_2, lazy[0], lazy[1])
                    return out.reshape(out.shape[0], -1)
                return backend.summate_incompr(cov_samples, z_1, z_2, pos, num_threads)
            return orig_si(cov_samples, z_1, z_2, _materialise(pos), num_threads)

        def _summate_fourier(spectrum_factor, modes, z_1, z_2, pos, num_threads=None):
            """A wrapper function for calling the Fourier algorithms (B200 first)."""
            if getattr(config, "USE_GSTOOLS_B200", False):
                lazy = _lookup_lazy(pos)
                if lazy is not None:
                    return backend.summate_fourier_structured(spectrum_factor, modes, z_1, z_2,
                                                              lazy[0], lazy[1]).reshape(-1)
                return backend.summate_fourier(spectrum_factor, modes, z_1, z_2, pos, num_threads)
            return orig_sf(spectrum_factor, modes, z_1, z_2, _materialise(pos), num_threads)

        def _materialise(pos):
            lazy = _lookup_lazy(pos)
            if lazy is None:
                return pos
            grid = gen.generate_grid(lazy[0])
            return grid if lazy[1] is None else np.dot(lazy[1], grid)

        gen._summate = _summate
        gen._summate_incompr = _summate_incompr
        gen._summate_fourier = _summate_fourier

        if lazy_grid:
            exact = (gen.RandMeth, gen.IncomprRandMeth, gen.Fourier)

            def pre_pos(self, pos=None, mesh_type="unstructured", info=False):
                generator = getattr(self, "_generator", None)
                model = getattr(self, "model", None)
                lazy_ok = (
                    getattr(config, "USE_GSTOOLS_B200", False)
                    and type(generator) in exact
                    and model is not None
                    and not model.latlon
                    and model.dim >= 2
                )
                if not lazy_ok:
                    return orig_pre_pos(self, pos, mesh_type, info)
                info_ret = {"deleted": False}
                if pos is None:
                    if self.pos is None:
                        raise ValueError("Field: no position tuple 'pos' present")
                else:
                    info_ret = self.set_pos(pos, mesh_type, info=True)
                if self.mesh_type == "unstructured" or model.field_dim != model.dim \
                        or getattr(generator, "zero_var", False):
                    out = orig_pre_pos(self, None, self.mesh_type, False)
                    return out + info * (info_ret,)
                matrix = matrix_isometrize(model.dim, model.angles, model.anis)
                lazy = LazyGridPos(self.pos, matrix)
                return (lazy, self.field_shape) + info * (info_ret,)

            pre_pos.__doc__ = orig_pre_pos.__doc__
            fbase.Field.pre_pos = pre_pos
        else:
            fbase.Field.pre_pos = orig_pre_pos
        _STATE["enabled"] = True
    return gstools


def disable():
    """Restore the reference's own wrappers and ``Field.pre_pos``."""
    with _LOCK:
        if not _STATE["enabled"]:
            return
        gen, fbase, config = _STATE["gen"], _STATE["fbase"], _STATE["config"]
        gen._summate = _STATE["orig_summate"]
        gen._summate_incompr = _STATE["orig_summate_incompr"]
        gen._summate_fourier = _STATE["orig_summate_fourier"]
        fbase.Field.pre_pos = _STATE["orig_pre_pos"]
        config.USE_GSTOOLS_B200 = False
        _STATE["enabled"] = False
