"""CPU oracle for the randomisation-method summators -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may
import this package.  ``gstools_b200`` never does.

Two restatements of the same algorithm (see ``summate_oracle.c`` for the
reference citations):

* :func:`summate` / :func:`summate_incompr` -- the C/OpenMP library
  ``liboracle.so`` built by ``oracle/Makefile`` (reference loop order, libm).
* :func:`summate_np` / :func:`summate_incompr_np` -- a numpy restatement used to
  cross-check the C build on small cases.

Parity pin: ``tests/test_oracle_golden.py`` (golden values asserted by the
reference's own tests, mode arrays produced by the reference's ``RandMeth``).
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile ``liboracle.so`` with gcc (``make -C oracle``)."""
    srcs = [os.path.join(_HERE, f) for f in ("summate_oracle.c", "krige_oracle.c", "Makefile")]
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs)
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        for name in ("oracle_summate", "oracle_summate_incompr"):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [dp, dp, dp, dp, ctypes.c_int, ctypes.c_int64,
                           ctypes.c_int64, dp, ctypes.c_int]
        lib.oracle_summate_fourier.restype = ctypes.c_int
        lib.oracle_summate_fourier.argtypes = [dp, dp, dp, dp, dp, ctypes.c_int, ctypes.c_int64,
                                               ctypes.c_int64, dp, ctypes.c_int]
        lib.oracle_max_threads.restype = ctypes.c_int
        i64 = ctypes.c_int64
        lib.oracle_calc_field_krige_and_variance.restype = ctypes.c_int
        lib.oracle_calc_field_krige_and_variance.argtypes = [dp, dp, i64, dp, i64, i64, dp, dp,
                                                             ctypes.c_int]
        lib.oracle_calc_field_krige.restype = ctypes.c_int
        lib.oracle_calc_field_krige.argtypes = [dp, dp, i64, dp, i64, i64, dp, ctypes.c_int]
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(_load().oracle_max_threads())


def _prep(cov_samples, z_1, z_2, pos):
    cov = np.ascontiguousarray(cov_samples, dtype=np.float64)
    z1 = np.ascontiguousarray(z_1, dtype=np.float64)
    z2 = np.ascontiguousarray(z_2, dtype=np.float64)
    p = np.ascontiguousarray(pos, dtype=np.float64)
    if cov.ndim != 2 or p.ndim != 2 or cov.shape[0] != p.shape[0]:
        raise ValueError("oracle: cov_samples (d,N) and pos (d,n) must share d")
    if z1.shape != (cov.shape[1],) or z2.shape != (cov.shape[1],):
        raise ValueError("oracle: z_1, z_2 must have shape (N,)")
    return cov, z1, z2, p


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def summate(cov_samples, z_1, z_2, pos, num_threads=None):
    """C/OpenMP oracle of ``summate`` (reference call site generator.py:42-48)."""
    cov, z1, z2, p = _prep(cov_samples, z_1, z_2, pos)
    out = np.zeros(p.shape[1], dtype=np.float64)
    rc = _load().oracle_summate(_ptr(cov), _ptr(z1), _ptr(z2), _ptr(p), cov.shape[0],
                                cov.shape[1], p.shape[1], _ptr(out),
                                int(num_threads or 0))
    if rc:
        raise RuntimeError(f"oracle_summate failed: {rc}")
    return out


def summate_incompr(cov_samples, z_1, z_2, pos, num_threads=None):
    """C/OpenMP oracle of ``summate_incompr`` (generator.py:51-64)."""
    cov, z1, z2, p = _prep(cov_samples, z_1, z_2, pos)
    out = np.zeros((cov.shape[0], p.shape[1]), dtype=np.float64)
    rc = _load().oracle_summate_incompr(_ptr(cov), _ptr(z1), _ptr(z2), _ptr(p),
                                        cov.shape[0], cov.shape[1], p.shape[1],
                                        _ptr(out), int(num_threads or 0))
    if rc:
        raise RuntimeError(f"oracle_summate_incompr failed: {rc}")
    return out


def summate_fourier(spectrum_factor, modes, z_1, z_2, pos, num_threads=None):
    """C/OpenMP oracle of ``summate_fourier`` (generator.py:67-75)."""
    cov, z1, z2, p = _prep(modes, z_1, z_2, pos)
    sf = np.ascontiguousarray(spectrum_factor, dtype=np.float64)
    if sf.shape != z1.shape:
        raise ValueError("oracle: spectrum_factor must have shape (N,)")
    out = np.zeros(p.shape[1], dtype=np.float64)
    rc = _load().oracle_summate_fourier(_ptr(sf), _ptr(cov), _ptr(z1), _ptr(z2), _ptr(p),
                                        cov.shape[0], cov.shape[1], p.shape[1], _ptr(out),
                                        int(num_threads or 0))
    if rc:
        raise RuntimeError(f"oracle_summate_fourier failed: {rc}")
    return out


def summate_np(cov_samples, z_1, z_2, pos, num_threads=None):
    """numpy restatement (small cases): phase matrix then two mat-vecs."""
    cov, z1, z2, p = _prep(cov_samples, z_1, z_2, pos)
    out = np.zeros(p.shape[1])
    step = max(1, 2_000_000 // max(1, cov.shape[1]))
    for a in range(0, p.shape[1], step):
        phase = p[:, a:a + step].T @ cov            # (n, N)
        out[a:a + step] = np.cos(phase) @ z1 + np.sin(phase) @ z2
    return out


def summate_incompr_np(cov_samples, z_1, z_2, pos, num_threads=None):
    """numpy restatement of the projector form (generator.py:479-495)."""
    cov, z1, z2, p = _prep(cov_samples, z_1, z_2, pos)
    d = cov.shape[0]
    k2 = np.sum(cov * cov, axis=0)
    e1 = np.zeros((d, 1))
    e1[0] = 1.0
    proj = e1 - cov * cov[0] / k2                   # (d, N)
    out = np.zeros((d, p.shape[1]))
    step = max(1, 2_000_000 // max(1, cov.shape[1]))
    for a in range(0, p.shape[1], step):
        phase = p[:, a:a + step].T @ cov
        amp = np.cos(phase) * z1 + np.sin(phase) * z2   # (n, N)
        out[:, a:a + step] = proj @ amp.T
    return out


def apply_epilogue(raw, scale, adds=()):
    """numpy restatement of the caller epilogue the reference applies to the summed modes:
    ``scale * raw`` (generator.py:269-270, 561-567), then one array pass per additive term
    (nugget, ``field += mean``, ``field += trend``: normalizer/tools.py:99-103).  Each entry of
    ``adds`` is a scalar or one value per component of a vector field ``(d, ...)``."""
    out = np.float64(scale) * np.asarray(raw, dtype=np.float64)
    for a in adds:
        a = np.asarray(a, dtype=np.float64).reshape(-1)
        if a.size > 1:
            a = a.reshape((a.size,) + (1,) * (out.ndim - 1))
        out = out + a
    return out


def apply_point_epilogue(field, gain=None, offset=None, adds=()):
    """numpy restatement of the per-point step of a conditioned field, one array pass per operation in
    the reference's order: ``var_scale * rawfield``, ``rawkrige + ...``, ``+ nugget`` (cond_srf.py:145-150),
    then the constant mean / trend of ``post_field`` (normalizer/tools.py:99-103)."""
    out = np.asarray(field, dtype=np.float64)
    if gain is not None:
        out = np.asarray(gain, dtype=np.float64).reshape(out.shape) * out
    if offset is not None:
        out = np.asarray(offset, dtype=np.float64).reshape(out.shape) + out
    for a in adds:
        out = out + np.float64(a)
    return out


def cond_scaling_np(error, sill, var):
    """``(krige_var, var_scale)``: the clamp of krige/base.py:296-298 and CondSRF.get_scaling without
    nugget (cond_srf.py:175-177)."""
    krige_var = np.maximum(sill - np.asarray(error, dtype=np.float64), 0)
    return krige_var, np.sqrt(krige_var / var)


def _prep_krige(krig_mat, krig_vecs, cond):
    mat = np.ascontiguousarray(krig_mat, dtype=np.float64)
    kv = np.ascontiguousarray(krig_vecs, dtype=np.float64)
    c = np.ascontiguousarray(cond, dtype=np.float64)
    if mat.ndim != 2 or mat.shape[0] != mat.shape[1] or kv.ndim != 2 or kv.shape[0] != mat.shape[0] \
            or c.shape != (mat.shape[0],):
        raise ValueError("oracle: krig_mat (K,K), krig_vecs (K,n), cond (K,)")
    return mat, kv, c


def calc_field_krige_and_variance(krig_mat, krig_vecs, cond, num_threads=None):
    """C/OpenMP oracle of ``calc_field_krige_and_variance`` (krige/base.py:42-61, 307-317)."""
    mat, kv, c = _prep_krige(krig_mat, krig_vecs, cond)
    n = kv.shape[1]
    field, error = np.zeros(n), np.zeros(n)
    rc = _load().oracle_calc_field_krige_and_variance(_ptr(mat), _ptr(kv), max(n, 1), _ptr(c),
                                                      mat.shape[0], n, _ptr(field), _ptr(error),
                                                      int(num_threads or 0))
    if rc:
        raise RuntimeError(f"oracle_calc_field_krige_and_variance failed: {rc}")
    return field, error


def calc_field_krige(krig_mat, krig_vecs, cond, num_threads=None):
    """C/OpenMP oracle of ``calc_field_krige`` (krige/base.py:42-49)."""
    mat, kv, c = _prep_krige(krig_mat, krig_vecs, cond)
    n = kv.shape[1]
    field = np.zeros(n)
    rc = _load().oracle_calc_field_krige(_ptr(mat), _ptr(kv), max(n, 1), _ptr(c), mat.shape[0], n,
                                         _ptr(field), int(num_threads or 0))
    if rc:
        raise RuntimeError(f"oracle_calc_field_krige failed: {rc}")
    return field


def calc_field_krige_and_variance_np(krig_mat, krig_vecs, cond, num_threads=None):
    """numpy restatement (BLAS order) used to cross-check the C build."""
    mat, kv, c = _prep_krige(krig_mat, krig_vecs, cond)
    mk = mat @ kv
    return c @ mk, np.einsum("ij,ij->j", kv, mk)


# ---------------------------------------------------------------------------------------------
# right-hand sides of the kriging system (numpy restatement of Krige._get_krige_vecs)
# ---------------------------------------------------------------------------------------------
def cov_cor_np(kind, h, param=0.0):
    """Normalised correlation ``cor(h)`` of the models with a device implementation, restated from
    src/gstools/covmodel/models.py (Gaussian :139-141, Exponential :213-215, Stable :343-345,
    Rational :610-612, Cubic :654-657, Linear :687-689, Circular :728-738, Spherical :774-777)."""
    h = np.asarray(np.abs(h), dtype=np.float64)
    if kind == "Gaussian":
        return np.exp(-(h**2))
    if kind == "Exponential":
        return np.exp(-h)
    if kind == "Stable":
        return np.exp(-np.power(h, param))
    if kind == "Rational":
        return np.power(1 + h**2 / param, -param)
    if kind == "Cubic":
        h = np.minimum(h, 1.0)
        return 1.0 - 7 * h**2 + 8.75 * h**3 - 3.5 * h**5 + 0.75 * h**7
    if kind == "Linear":
        return np.maximum(1 - h, 0.0)
    if kind == "Circular":
        res = np.zeros_like(h)
        lo = h < 1.0
        hl = h[lo]
        res[lo] = 2 / np.pi * (np.arccos(hl) - hl * np.sqrt(1 - hl**2))
        return res
    if kind == "Spherical":
        h = np.minimum(h, 1.0)
        return 1.0 - 1.5 * h + 0.5 * h**3
    raise ValueError(kind)


def krige_vecs_np(kind, var, len_rescaled, sill, cond_pos, pos, unbiased=True, tail_rows=None,
                  param=0.0, exact=False):
    """``Krige._get_krige_vecs`` (krige/base.py:359-388) for one chunk: covariance rows from the
    pairwise distances of the isometrised positions (base.py:430-450, scipy ``cdist``), the row of
    ones of an unbiased system, then the drift rows."""
    cond_pos, pos = np.asarray(cond_pos, dtype=np.float64), np.asarray(pos, dtype=np.float64)
    diff = cond_pos[:, :, None] - pos[:, None, :]
    r = np.sqrt(np.sum(diff * diff, axis=0))
    cov = var * cov_cor_np(kind, r / len_rescaled, param)            # covmodel/tools.py:65-76
    if exact:                                                       # cov_nugget, covmodel/base.py:313-320
        cov[np.isclose(r, 0)] = sill
    rows = [cov]
    if unbiased:
        rows.append(np.ones((1, pos.shape[1])))
    if tail_rows is not None and np.size(tail_rows):
        rows.append(np.asarray(tail_rows, dtype=np.float64).reshape(-1, pos.shape[1]))
    return np.concatenate(rows, axis=0)


def krige_evaluate(spec, krig_mat, cond, cond_pos, pos, unbiased=True, tail_rows=None, num_threads=None):
    """Right-hand sides by :func:`krige_vecs_np`, then the C oracle of the native evaluation."""
    kv = krige_vecs_np(spec["kind"], spec["var"], spec["len_rescaled"], spec.get("sill", spec["var"]),
                       cond_pos, pos, unbiased, tail_rows, spec.get("param", 0.0), spec.get("exact", False))
    return calc_field_krige_and_variance(krig_mat, kv, cond, num_threads)
