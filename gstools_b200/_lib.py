"""ctypes binding of the C ABI declared in ``include/gsb200.h``.

The product path fails loudly when the CUDA library is missing or when no GPU is
present -- there is no CPU fallback (and nothing under ``oracle/`` is ever imported here).
"""

from __future__ import annotations

import ctypes
import os

from . import _build

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)
_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int

MEM_HOST = 0
MEM_DEVICE = 1

EPI_MAX_ADD = 4
EPI_MAX_COMP = 3


class Epilogue(ctypes.Structure):
    """``gsb_epilogue`` of include/gsb200.h: v = scale*sum, then + add[0][c], + add[1][c], ..."""

    _fields_ = [("scale", ctypes.c_double), ("n_add", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("add", (ctypes.c_double * EPI_MAX_COMP) * EPI_MAX_ADD)]


_epi_p = ctypes.POINTER(Epilogue)


class PointEpilogue(ctypes.Structure):
    """``gsb_point_epilogue`` of include/gsb200.h: v = gain[i]*v; v = offset[i] + v; then + add[0], ..."""

    _fields_ = [("gain", ctypes.c_void_p), ("offset", ctypes.c_void_p), ("n_add", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("add", ctypes.c_double * EPI_MAX_ADD)]


_pepi_p = ctypes.POINTER(PointEpilogue)

COV_TYPES = {"Gaussian": 1, "Exponential": 2, "Stable": 3, "Rational": 4, "Cubic": 5, "Linear": 6,
             "Circular": 7, "Spherical": 8}


PDF_KINDS = {"Exponential": 1, "Matern": 2, "Gaussian": 3}


class CovModelSpec(ctypes.Structure):
    """``gsb_cov_model`` of include/gsb200.h."""

    _fields_ = [("type", ctypes.c_int32), ("exact", ctypes.c_int32), ("var", ctypes.c_double),
                ("len_rescaled", ctypes.c_double), ("sill", ctypes.c_double), ("param", ctypes.c_double)]


_cov_p = ctypes.POINTER(CovModelSpec)

# every symbol include/gsb200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gsb_version": (_int, []),
    "gsb_last_error": (ctypes.c_char_p, []),
    "gsb_device_count": (_int, [ctypes.POINTER(_int)]),
    "gsb_summate": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _i64, _i64, _vp, _int, _int, _vp]),
    "gsb_summate_incompr": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _i64, _i64, _vp, _i64, _int,
                                   _int, _vp]),
    "gsb_summate_structured": (_int, [_vp, _vp, _vp, _vp, _c_int64_p, _vp, _int, _i64, _i64, _vp,
                                      _int, _int, _vp]),
    "gsb_summate_incompr_structured": (_int, [_vp, _vp, _vp, _vp, _c_int64_p, _vp, _int, _i64,
                                              _i64, _vp, _int, _int, _vp]),
    "gsb_summate_ex": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _i64, _i64, _vp, _epi_p, _int, _int,
                              _vp]),
    "gsb_summate_incompr_ex": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _i64, _i64, _vp, _i64, _epi_p,
                                      _int, _int, _vp]),
    "gsb_summate_structured_ex": (_int, [_vp, _vp, _vp, _vp, _c_int64_p, _vp, _int, _i64, _i64, _vp,
                                         _epi_p, _int, _int, _vp]),
    "gsb_summate_incompr_structured_ex": (_int, [_vp, _vp, _vp, _vp, _c_int64_p, _vp, _int, _i64,
                                                 _i64, _vp, _epi_p, _int, _int, _vp]),
    "gsb_summate_pp": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _i64, _i64, _vp, _epi_p, _pepi_p, _int, _int,
                              _vp]),
    "gsb_summate_structured_pp": (_int, [_vp, _vp, _vp, _vp, _c_int64_p, _vp, _int, _i64, _i64, _vp,
                                         _epi_p, _pepi_p, _int, _int, _vp]),
    "gsb_cond_scaling": (_int, [_vp, _i64, ctypes.c_double, ctypes.c_double, _vp, _vp, _int, _vp]),
    "gsb_summate_fourier": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _i64, _i64, _vp, _int, _int,
                                   _vp]),
    "gsb_summate_fourier_structured": (_int, [_vp, _vp, _vp, _vp, _vp, _c_int64_p, _vp, _int, _i64,
                                              _vp, _int, _int, _vp]),
    "gsb_calc_field_krige_and_variance": (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _int, _int,
                                                 _vp]),
    "gsb_calc_field_krige": (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _vp, _int, _int, _vp]),
    "gsb_krige_evaluate": (_int, [_cov_p, _vp, _vp, _i64, _vp, _i64, _int, _vp, _i64, _i64, _int, _vp, _i64,
                                  _vp, _vp, _int, _int, _vp]),
    "gsb_krige_evaluate_structured": (_int, [_cov_p, _vp, _vp, _i64, _vp, _i64, _int, _vp, _c_int64_p, _vp,
                                             _int, _vp, _i64, _vp, _vp, _int, _int, _vp]),
    "gsb_sample_radii_mcmc": (_int, [_int, _int, ctypes.c_double, ctypes.c_double, _vp, _int, _vp, _int, _vp,
                                     _int, _int, _int, _vp]),
    "gsb_sample_radii_mcmc_cb": (_int, [_vp, _vp, _vp, _int, _vp, _int, _vp, _int, _int, _int, _vp]),
    "gsb_sample_modes_batch": (_int, [_int, _int, ctypes.c_double, ctypes.c_double, _vp, _i64, _i64, _int, _int, _i64,
                                      ctypes.c_double, ctypes.c_double, _int, _vp, _vp, _vp, _vp, _vp]),
    "gsb_scale_shift": (_int, [_vp, _i64, ctypes.c_double, ctypes.c_double, _int, _vp]),
    "gsb_plan_create": (_int, [ctypes.POINTER(_int), _int, ctypes.POINTER(_vp)]),
    "gsb_plan_destroy": (_int, [_vp]),
    "gsb_plan_info": (_int, [_vp, ctypes.POINTER(_int), ctypes.POINTER(_int), ctypes.POINTER(_int)]),
    "gsb_plan_share": (_int, [_i64, _int, _int, _c_int64_p, _c_int64_p]),
    "gsb_plan_summate": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _i64, _i64, _vp, _i64, _int, _epi_p, _vp, _int,
                                _int, _vp]),
    "gsb_plan_summate_structured": (_int, [_vp, _vp, _vp, _vp, _vp, _c_int64_p, _vp, _int, _i64, _i64, _vp, _int,
                                           _epi_p, _vp, _int, _int, _vp]),
    "gsb_plan_krige_evaluate": (_int, [_vp, _cov_p, _vp, _vp, _i64, _vp, _i64, _int, _vp, _i64, _i64, _vp, _c_int64_p, _vp,
                                       _int, _vp, _i64, _vp, _vp, _int, _int, _vp]),
    "gsb_summate_structured_slab": (_int, [_vp, _vp, _vp, _vp, _c_int64_p, _vp, _int, _i64, _i64, _i64, _i64, _vp, _int,
                                           _epi_p, _vp, _int, _int, _vp]),
    "gsb_ipc_export": (_int, [_vp, _int, _vp, _c_int64_p]),
    "gsb_ipc_open": (_int, [_vp, _int, ctypes.POINTER(_vp)]),
    "gsb_ipc_close": (_int, [_vp, _int]),
    "gsb_set_option": (_int, [ctypes.c_char_p, _i64]),
    "gsb_release_memory": (_int, [_int]),
    "gsb_get_counter": (_i64, [ctypes.c_char_p]),
    "gsb_streamk_plan": (_int, [_i64, _i64, _i64, _i64, _int, _int, _c_int64_p, ctypes.POINTER(ctypes.c_int32),
                                ctypes.POINTER(_int)]),
    "gsb_kernel_times": (_int, [_c_double_p, _c_int64_p]),
    "gsb_measure_fp64_peak": (_int, [_int, _int, ctypes.c_double, _c_double_p]),
}

LN_PDF_FN = ctypes.CFUNCTYPE(_int, _c_double_p, _int, _c_double_p, _vp)

_lib = None


class GSB200Error(RuntimeError):
    """Device-side failure reported by the C ABI (CUDA error, no device, ...)."""


def lib_path() -> str:
    return _build.LIB


def load():
    """Load ``libgsb200.so`` (never builds implicitly on a machine without nvcc)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m gstools_b200._build` "
            "(or __graft_entry__.build()).  gstools_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().gsb_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = last_error()
    if rc == 1:
        raise ValueError(f"{what}: {msg}")
    raise GSB200Error(f"{what}: {msg}")


def device_count() -> int:
    c = _int(0)
    check(load().gsb_device_count(ctypes.byref(c)), "gsb_device_count")
    return int(c.value)


def set_option(name: str, value: int):
    check(load().gsb_set_option(name.encode(), int(value)), "gsb_set_option")


def release_memory(device=None, pinned=True):
    """Return the scratch memory the library keeps cached in the device's memory pool to the driver and
    (``pinned=True``) the page-locked host blocks of result arrays that have been freed: they otherwise stay
    in torch's pinned-memory cache for reuse by the next call (result arrays still alive stay pinned)."""
    if device is None:
        from . import backend

        device = backend.get_device()
    check(load().gsb_release_memory(int(device)), "gsb_release_memory")
    if pinned:
        try:
            import torch

            torch._C._host_emptyCache()
        except Exception:  # noqa: BLE001  (older torch: nothing to trim)
            pass


def get_counter(name: str) -> int:
    return int(load().gsb_get_counter(name.encode()))


def kernel_times():
    """(total_ms, n_launches) of the timed dominant kernels since the last call."""
    ms = ctypes.c_double(0.0)
    n = ctypes.c_int64(0)
    check(load().gsb_kernel_times(ctypes.byref(ms), ctypes.byref(n)), "gsb_kernel_times")
    return float(ms.value), int(n.value)


def plan_share(n: int, parts: int, part: int):
    """``[lo, hi)`` of part ``part`` of ``parts`` over ``n`` units, as the multi-GPU plan cuts them (host only)."""
    lo, hi = ctypes.c_int64(0), ctypes.c_int64(0)
    check(load().gsb_plan_share(int(n), int(parts), int(part), ctypes.byref(lo), ctypes.byref(hi)), "gsb_plan_share")
    return int(lo.value), int(hi.value)


def streamk_plan(tile_begin, tile_end, ly, lc, n_stages, max_grid):
    """Share boundaries ``[(tile, stage), ...]`` (grid + 1 entries) of the stream-K contraction (host only)."""
    tiles = (ctypes.c_int64 * (max_grid + 1))()
    stages = (ctypes.c_int32 * (max_grid + 1))()
    grid = _int(0)
    check(load().gsb_streamk_plan(int(tile_begin), int(tile_end), int(ly), int(lc), int(n_stages), int(max_grid),
                                  tiles, stages, ctypes.byref(grid)), "gsb_streamk_plan")
    return [(int(tiles[c]), int(stages[c])) for c in range(grid.value + 1)]


def measure_fp64_peak(device: int = 0, kind: int = 0, seconds: float = 0.3) -> float:
    """Sustained FP64 FMA/s of the DFMA pipe (kind 0) or the DMMA tensor path (kind 1)."""
    out = ctypes.c_double(0.0)
    check(load().gsb_measure_fp64_peak(int(device), int(kind), float(seconds), ctypes.byref(out)),
          "gsb_measure_fp64_peak")
    return float(out.value)
