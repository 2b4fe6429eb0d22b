"""The five BASELINE.json configs as concrete inputs of the hot path (shared by bench.py and tests).

Mode sets (cov_samples, z_1, z_2) for configs 2-5 were drawn by the UNMODIFIED reference
(tests/golden/make_config_modes.py -> tests/golden/config_modes.npz); config 1 comes from
tests/golden/config1_gaussian2d_100x100.npz.  Positions are synthetic as described in
SURVEY.md section 8(d).  Nothing here touches /root/reference at run time.
"""

from __future__ import annotations

import json
import os

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(REPO, "tests", "golden")

_modes_cache = None


def _modes():
    global _modes_cache
    if _modes_cache is None:
        _modes_cache = np.load(os.path.join(GOLDEN, "config_modes.npz"))
    return _modes_cache


def mode_set(tag):
    m = _modes()
    return m[tag + "_cov"], m[tag + "_z1"], m[tag + "_z2"]


def synth_mode_set(dim, n_modes, seed, len_scale=10.0):
    """Synthetic stand-in with the same shape/statistics class as RandMeth's modes."""
    rs = np.random.RandomState(seed)
    z1, z2 = rs.normal(size=n_modes), rs.normal(size=n_modes)
    v = rs.normal(size=(dim, n_modes))
    v /= np.linalg.norm(v, axis=0)
    rad = np.minimum(np.abs(rs.standard_cauchy(n_modes)), 100.0) / len_scale
    return rad * v, z1, z2


def config1():
    """SRF Gaussian 2D var=1 len_scale=10, structured 100x100, mode_no=1000, seed=20170519."""
    d = np.load(os.path.join(GOLDEN, "config1_gaussian2d_100x100.npz"))
    meta = json.loads(str(d["meta"]))
    return dict(name="C1 Gaussian 2D 100x100 structured N=1000", kind="scalar", dim=2, var=1.0,
                cov=d["cov_samples"], z1=d["z_1"], z2=d["z_2"], axes=[d["axis0"], d["axis1"]],
                matrix=None, pos=d["pos"], raw=d["raw"], field=d["field"], meta=meta)


def config2(edge=512):
    """SRF Exponential 3D structured edge^3 mesh, mode_no=1000 (BASELINE metric config)."""
    cov, z1, z2 = mode_set("c2")
    axes = [np.arange(float(edge))] * 3
    return dict(name=f"C2 Exponential 3D {edge}^3 structured N=1000", kind="scalar", dim=3, var=1.0,
                cov=cov, z1=z1, z2=z2, axes=axes, matrix=None)


def config3(n=20_000_000):
    """SRF Matern(nu=1) 2D unstructured, n random points in [0,1000)^2, mode_no=10000."""
    cov, z1, z2 = mode_set("c3")
    pos = np.random.RandomState(20170519).uniform(0.0, 1000.0, (2, n))
    return dict(name=f"C3 Matern 2D unstructured n={n} N=10000", kind="scalar", dim=2, var=1.0,
                cov=cov, z1=z1, z2=z2, pos=pos)


def config4(edge=256):
    """Incompressible vector field 3D (summate_incompr) on an edge^3 structured mesh, N=1000."""
    cov, z1, z2 = mode_set("c4")
    axes = [np.arange(float(edge))] * 3
    return dict(name=f"C4 incompressible Gaussian 3D {edge}^3 structured N=1000", kind="incompr",
                dim=3, var=1.0, cov=cov, z1=z1, z2=z2, axes=axes, matrix=None)


def config5(edge=128, n_real=256):
    """CondSRF ensemble: n_real realisations on an edge^3 mesh (summation part of the path).

    The first 8 mode sets are the reference's own draws for seeds MasterRNG(20170519)();
    the remaining ones are synthetic stand-ins of the same shape (documented in DESIGN.md).
    """
    sets = [mode_set(f"c5_{i}") for i in range(min(8, n_real))]
    sets += [synth_mode_set(3, 1000, 1000 + i) for i in range(max(0, n_real - 8))]
    cov = np.stack([s[0] for s in sets])
    z1 = np.stack([s[1] for s in sets])
    z2 = np.stack([s[2] for s in sets])
    axes = [np.arange(float(edge))] * 3
    return dict(name=f"C5 ensemble {n_real} x {edge}^3 structured N=1000", kind="scalar", dim=3,
                var=1.0, cov=cov, z1=z1, z2=z2, axes=axes, matrix=None)


def grid_points(axes, matrix=None, index=None):
    """Flat (dim, m) positions of mesh nodes (all, or the flat C-order indices in ``index``)."""
    lens = [len(a) for a in axes]
    if index is None:
        grid = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    else:
        idx = np.unravel_index(np.asarray(index), lens)
        grid = np.stack([np.asarray(a)[i] for a, i in zip(axes, idx)])
    return grid if matrix is None else np.asarray(matrix) @ grid
