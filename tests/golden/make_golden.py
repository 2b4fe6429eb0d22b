#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Run from the repo root (needs /root/reference, which only exists here):

    python tests/golden/make_golden.py

For every case the reference's own generator classes (src/gstools/field/generator.py)
are driven exactly as the reference's tests drive them; the arrays that cross the
``_summate`` / ``_summate_incompr`` boundary (generator.py:42-64) are recorded
together with the returned sum.  The ``asserts`` entries are the literal golden
values the reference's tests assert, with their file:line, so the fixtures are
self-describing on the GPU box where /root/reference does not exist.

The native summator in this run is the CPU oracle (tools/refstubs/gstools_cython);
radius sampling for 3D models goes through tools/refstubs/emcee, whose stream
reproduces the reference's 16-digit 3D goldens (checked below before saving).
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refharness  # noqa: E402

gs = refharness.import_gstools()
from gstools.field import generator as gen  # noqa: E402

RECORD = []
_orig_summate = gen._summate
_orig_summate_incompr = gen._summate_incompr


def _rec(kind, fn):
    def wrapper(cov_samples, z_1, z_2, pos, num_threads=None):
        out = fn(cov_samples, z_1, z_2, pos, num_threads)
        RECORD.append(dict(kind=kind, cov_samples=np.array(cov_samples), z_1=np.array(z_1),
                           z_2=np.array(z_2), pos=np.array(pos), raw=np.array(out)))
        return out
    return wrapper


gen._summate = _rec("scalar", _orig_summate)
gen._summate_incompr = _rec("incompr", _orig_summate_incompr)
_orig_summate_fourier = gen._summate_fourier


def _rec_fourier(spectrum_factor, modes, z_1, z_2, pos, num_threads=None):
    out = _orig_summate_fourier(spectrum_factor, modes, z_1, z_2, pos, num_threads)
    RECORD.append(dict(kind="fourier", cov_samples=np.array(modes), z_1=np.array(z_1), z_2=np.array(z_2),
                       pos=np.array(pos), raw=np.array(out), spectrum_factor=np.array(spectrum_factor)))
    return out


gen._summate_fourier = _rec_fourier


def save(name, field, asserts, extra=None, places_default=7):
    """Save the LAST recorded boundary crossing plus the returned field."""
    rec = RECORD[-1]
    field = np.asarray(field)
    meta = dict(name=name, kind=rec["kind"], asserts=[])
    for idx, val, cite in asserts:
        got = float(field[idx]) if idx is not None else None
        meta["asserts"].append(dict(index=list(idx) if isinstance(idx, tuple) else idx,
                                    value=val, places=places_default, cite=cite))
        if idx is not None:
            assert round(got - val, places_default) == 0, (name, idx, got, val)
    if extra:
        meta.update({k: v for k, v in extra.items() if not isinstance(v, np.ndarray)})
    arrays = dict(cov_samples=rec["cov_samples"], z_1=rec["z_1"], z_2=rec["z_2"],
                  pos=rec["pos"], raw=rec["raw"], field=field)
    if "spectrum_factor" in rec:
        arrays["spectrum_factor"] = rec["spectrum_factor"]
    if extra:
        arrays.update({k: v for k, v in extra.items() if isinstance(v, np.ndarray)})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=json.dumps(meta), **arrays)
    print(f"  wrote {name}.npz  kind={rec['kind']} cov={rec['cov_samples'].shape} "
          f"pos={rec['pos'].shape}")


def main():
    seed = 19031977
    x_t = np.linspace(0.0, 10.0, 10)
    y_t = np.linspace(-5.0, 5.0, 10)
    z_t = np.linspace(-6.0, 8.0, 10)
    x_g = np.linspace(0.0, 10.0, 9)
    y_g = np.linspace(-5.0, 5.0, 16)

    # ---- tests/test_randmeth.py:15-71 -------------------------------------------------
    for dim, pos, vals, cite in [
        (1, (x_t,), (3.19799030, 2.44848295), "tests/test_randmeth.py:35-36"),
        (2, (x_t, y_t), (1.67318010, 2.12310269), "tests/test_randmeth.py:40-41"),
        (3, (x_t, y_t, z_t), (1.3240234883187239, 1.6367244277732766),
         "tests/test_randmeth.py:45-46"),
    ]:
        model = gs.Gaussian(dim=dim, var=1.5, len_scale=3.5)
        rm = gen.RandMeth(model, mode_no=100, seed=seed)
        f = rm(pos)
        save(f"randmeth_{dim}d", f, [((0,), vals[0], cite), ((1,), vals[1], cite)],
             extra=dict(var=1.5, mode_no=100, scale=float(np.sqrt(1.5 / 100))))
    # reset: new seed, then mode_no=800 (test_randmeth.py:58-71)
    model = gs.Gaussian(dim=2, var=1.5, len_scale=3.5)
    rm = gen.RandMeth(model, mode_no=100, seed=seed)
    rm.seed = 74893621
    f = rm((x_t, y_t))
    save("randmeth_2d_reseed", f, [((0,), -1.94278053, "tests/test_randmeth.py:60"),
                                   ((1,), -1.12401651, "tests/test_randmeth.py:61")],
         extra=dict(var=1.5, mode_no=100, scale=float(np.sqrt(1.5 / 100))))
    rm.mode_no = 800  # seed stays 74893621, as in the reference test
    f = rm((x_t, y_t))
    save("randmeth_2d_modes800", f, [((0,), -3.20809251, "tests/test_randmeth.py:70"),
                                     ((1,), -2.62032778, "tests/test_randmeth.py:71")],
         extra=dict(var=1.5, mode_no=800, scale=float(np.sqrt(1.5 / 800))))

    # ---- tests/test_incomprrandmeth.py:34-59 -------------------------------------------
    model = gs.Gaussian(dim=2, var=1.5, len_scale=2.5)
    rm = gen.IncomprRandMeth(model, mode_no=100, seed=seed)
    f = rm((x_t, y_t))
    save("incompr_2d", f, [((0, 0), 0.50751115, "tests/test_incomprrandmeth.py:36"),
                           ((0, 1), 1.03291018, "tests/test_incomprrandmeth.py:37"),
                           ((1, 1), -0.22003005, "tests/test_incomprrandmeth.py:38")],
         extra=dict(var=1.5, mode_no=100))
    model = gs.Gaussian(dim=3, var=1.5, len_scale=2.5)
    rm = gen.IncomprRandMeth(model, mode_no=100, seed=seed)
    f = rm((x_t, y_t, z_t))
    save("incompr_3d", f, [((0, 0), 0.7924546333550331, "tests/test_incomprrandmeth.py:42"),
                           ((0, 1), 1.660747056686244, "tests/test_incomprrandmeth.py:43"),
                           ((1, 0), -0.28049855754819514, "tests/test_incomprrandmeth.py:44")],
         extra=dict(var=1.5, mode_no=100))
    model = gs.Gaussian(dim=2, var=1.5, len_scale=2.5)
    srf = gs.SRF(model, mean=(0.5, 0), generator="VectorField", seed=198412031)
    srf.structured((x_g, y_g))
    assert round(np.mean(srf.field[0]) - 1.3025621393180298, 7) == 0
    assert round(np.mean(srf.field[1]) - -0.04729596839446052, 7) == 0
    save("incompr_2d_vector_mean_struct", srf.field, [],
         extra=dict(var=1.5, mode_no=1000, axis0=x_g, axis1=y_g,
                    mean0=1.3025621393180298, mean1=-0.04729596839446052,
                    cite="tests/test_incomprrandmeth.py:50-59"))

    # ---- tests/test_srf.py:259-275 ------------------------------------------------------
    model = gs.Gaussian(dim=2, var=0.5, len_scale=1.0)
    srf = gs.SRF(model, mean=0.3, mode_no=100, generator="IncomprRandMeth", mean_velocity=0.5)
    rng = np.random.RandomState(123018)  # tests/test_srf.py:36-38
    xs_t = rng.uniform(0.0, 10, 100)
    ys_t = rng.uniform(0.0, 10, 100)
    f = srf((xs_t, ys_t), seed=476356)
    save("srf_incompr_unstruct", f, [((0, 0), 1.23693272, "tests/test_srf.py:269"),
                                     ((0, 1), 0.89242284, "tests/test_srf.py:270")],
         extra=dict(var=0.5, mode_no=100))
    xg = np.linspace(0.0, 12.0, 48)  # tests/test_srf.py:28-29 grids
    yg = np.linspace(0.0, 10.0, 46)
    f = srf((xg, yg), seed=4734654, mesh_type="structured")
    save("srf_incompr_struct", f, [((0, 0, 0), 1.07812013, "tests/test_srf.py:274"),
                                   ((0, 1, 0), 1.06180674, "tests/test_srf.py:275")],
         extra=dict(var=0.5, mode_no=100, axis0=xg, axis1=yg))

    # ---- config 1: README example / tests/test_pgs.py:33-75 ------------------------------
    model = gs.Gaussian(dim=2, var=1, len_scale=10)
    srf = gs.SRF(model, seed=20170519)
    ax = np.arange(100.0)
    f = srf.structured([ax, ax])
    save("config1_gaussian2d_100x100", f, [],
         extra=dict(var=1.0, mode_no=1000, scale=float(np.sqrt(1.0 / 1000)), axis0=ax, axis1=ax,
                    cite="README.md spatial random field example; BASELINE.json configs[0]"))

    # ---- structured 3D with rotation + anisotropy on non-uniform axes (side channel) ----
    model = gs.Exponential(dim=3, var=2.0, len_scale=[12.0, 5.0, 3.0], angles=[0.4, -0.3, 0.7])
    srf = gs.SRF(model, seed=20170519, mode_no=256)
    a0 = np.sort(np.random.RandomState(1).uniform(0, 40, 13))
    a1 = np.linspace(-7.0, 9.0, 17)
    a2 = np.sort(np.random.RandomState(2).uniform(-5, 25, 21))
    f = srf.structured([a0, a1, a2])
    from gstools.tools.geometric import matrix_isometrize
    mat = matrix_isometrize(model.dim, model.angles, model.anis)
    save("srf_exp3d_rot_anis_struct", f, [],
         extra=dict(var=2.0, mode_no=256, scale=float(np.sqrt(2.0 / 256)), axis0=a0, axis1=a1,
                    axis2=a2, matrix=np.array(mat),
                    cite="src/gstools/field/base.py:283-297; covmodel/base.py:572-582"))

    # ---- conditioned field (tests/test_condition.py style), 2D small ---------------------
    cond_pos = [np.array([0.3, 1.9, 1.1, 3.3, 4.7])]
    cond_val = np.array([0.47, 0.56, 0.74, 1.47, 1.74])
    model = gs.Gaussian(dim=1, var=0.5, len_scale=2)
    krige = gs.krige.Ordinary(model, cond_pos, cond_val)
    csrf = gs.CondSRF(krige, mode_no=100)
    gx = np.linspace(0.0, 15.0, 151)
    f = csrf((gx,), seed=20170519)
    save("condsrf_1d", f, [], extra=dict(var=0.5, mode_no=100, cond_pos=cond_pos[0],
                                          cond_val=cond_val, gridx=gx,
                                          cite="examples/06_conditioned_fields/00_condition_ensemble.py"))
    # ---- Fourier generator (next row f3): tests/test_fouriergen.py:13-57 -------------------
    fseed = 19900408
    L = [80, 30, 91]
    fx, fy, fz = np.linspace(0, L[0], 11), np.linspace(0, L[1], 31), np.linspace(0, L[2], 13)
    mode_no = [12, 6, 14]
    for dim, model, pos, idx, val, cite in [
        (1, gs.Gaussian(dim=1, var=0.5, len_scale=10.0), (fx,), (0,), 0.6236929351309081,
         "tests/test_fouriergen.py:48"),
        (2, gs.Gaussian(dim=2, var=2.0, len_scale=30.0), (fx, fy), (0, 0), -0.1431996611581266,
         "tests/test_fouriergen.py:52"),
        (3, gs.Gaussian(dim=3, var=2.1, len_scale=21.0), (fx, fy, fz), (0, 0, 0), -1.0433325279452803,
         "tests/test_fouriergen.py:56"),
    ]:
        srf = gs.SRF(model, generator="Fourier", mode_no=mode_no[:dim], period=L[:dim], seed=fseed)
        f = srf(pos, mesh_type="structured")
        extra = dict(var=float(model.var), mode_no=int(np.prod(mode_no[:dim])))
        for t, a in enumerate(pos):
            extra[f"axis{t}"] = np.asarray(a, dtype=float)
        save(f"fourier_{dim}d", f, [(idx, val, cite)], extra=extra)
    print("done:", len(RECORD), "boundary crossings recorded")


if __name__ == "__main__":
    main()
