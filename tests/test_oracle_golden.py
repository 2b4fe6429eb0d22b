"""The CPU oracle against the reference's golden values (the parity pin).  CPU only.

Fixtures in tests/golden/*.npz were recorded from the UNMODIFIED reference generators
(tests/golden/make_golden.py); every ``asserts`` entry is a literal value asserted by the
reference's own tests (file:line stored in the fixture).
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden, synth_modes


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_recorded_boundary(name, oracle_mod):
    meta, d = load_golden(name)
    if meta["kind"] == "fourier":
        got = oracle_mod.summate_fourier(d["spectrum_factor"], d["cov_samples"], d["z_1"], d["z_2"], d["pos"])
    else:
        fn = oracle_mod.summate if meta["kind"] == "scalar" else oracle_mod.summate_incompr
        got = fn(d["cov_samples"], d["z_1"], d["z_2"], d["pos"])
    # same library, same libm: bit-for-bit
    assert np.array_equal(got, d["raw"])


@pytest.mark.parametrize("name", ["randmeth_1d", "randmeth_2d", "randmeth_3d", "randmeth_2d_reseed",
                                  "randmeth_2d_modes800", "incompr_2d", "incompr_3d"])
def test_oracle_matches_reference_test_literals(name, oracle_mod):
    """sqrt(var/N) * oracle == the literals in tests/test_randmeth.py / test_incomprrandmeth.py."""
    meta, d = load_golden(name)
    fn = oracle_mod.summate if meta["kind"] == "scalar" else oracle_mod.summate_incompr
    n_modes = d["cov_samples"].shape[1]
    field = np.sqrt(meta["var"] / n_modes) * fn(d["cov_samples"], d["z_1"], d["z_2"], d["pos"])
    if meta["kind"] == "incompr":
        # IncomprRandMeth.__call__ adds mean_u * e1 with mean_u = 1 (generator.py:561-567)
        field[0] += 1.0
    assert meta["asserts"], "fixture carries no reference literals"
    for a in meta["asserts"]:
        got = field[tuple(a["index"])]
        assert round(got - a["value"], a["places"]) == 0, (a["cite"], got, a["value"])


@pytest.mark.parametrize("name", ["fourier_1d", "fourier_2d", "fourier_3d"])
def test_oracle_fourier_matches_reference_test_literals(name, oracle_mod):
    """tests/test_fouriergen.py:46-57: the Fourier generator returns the raw sum (generator.py:693-694)."""
    meta, d = load_golden(name)
    raw = oracle_mod.summate_fourier(d["spectrum_factor"], d["cov_samples"], d["z_1"], d["z_2"], d["pos"])
    field = raw.reshape(d["field"].shape)
    for a in meta["asserts"]:
        assert round(field[tuple(a["index"])] - a["value"], a["places"]) == 0, a["cite"]
    # and equals the modified-weights form used by the GPU kernels
    assert np.allclose(raw, oracle_mod.summate(d["cov_samples"], d["spectrum_factor"] * d["z_1"],
                                               d["spectrum_factor"] * d["z_2"], d["pos"]), rtol=0, atol=1e-12)


def test_oracle_3d_literals_are_16_digit(oracle_mod):
    """tests/test_randmeth.py:45-46 holds 16-digit goldens: reproduce them to 1e-15."""
    meta, d = load_golden("randmeth_3d")
    field = np.sqrt(1.5 / 100) * oracle_mod.summate(d["cov_samples"], d["z_1"], d["z_2"], d["pos"])
    assert abs(field[0] - 1.3240234883187239) < 1e-15
    assert abs(field[1] - 1.6367244277732766) < 1e-15


def test_vector_mean_fixture(oracle_mod):
    """tests/test_incomprrandmeth.py:50-59 (structured 9x16 vector field, default mode_no)."""
    meta, d = load_golden("incompr_2d_vector_mean_struct")
    assert abs(np.mean(d["field"][0]) - meta["mean0"]) < 1e-12
    assert abs(np.mean(d["field"][1]) - meta["mean1"]) < 1e-12


@pytest.mark.parametrize("dim", [1, 2, 3, 4])
def test_c_oracle_vs_numpy_restatement(dim, oracle_mod):
    cov, z1, z2 = synth_modes(dim, 257, seed=dim)
    pos = np.random.RandomState(3).uniform(-50, 300, (dim, 1234))
    a = oracle_mod.summate(cov, z1, z2, pos)
    b = oracle_mod.summate_np(cov, z1, z2, pos)
    assert np.max(np.abs(a - b)) < 1e-11
    if dim >= 2:
        av = oracle_mod.summate_incompr(cov, z1, z2, pos)
        bv = oracle_mod.summate_incompr_np(cov, z1, z2, pos)
        assert np.max(np.abs(av - bv)) < 1e-11
        # first component dominates: the projector is the identity minus a rank-1 term
        assert av.shape == (dim, 1234)


def test_oracle_threads_are_deterministic(oracle_mod):
    cov, z1, z2 = synth_modes(3, 100, seed=9)
    pos = np.random.RandomState(4).uniform(0, 100, (3, 5000))
    a = oracle_mod.summate(cov, z1, z2, pos, num_threads=1)
    b = oracle_mod.summate(cov, z1, z2, pos, num_threads=4)
    assert np.array_equal(a, b)


def test_oracle_edge_cases(oracle_mod):
    cov, z1, z2 = synth_modes(2, 10, seed=1)
    assert oracle_mod.summate(cov, z1, z2, np.zeros((2, 0))).shape == (0,)
    out = oracle_mod.summate(cov[:, :0], z1[:0], z2[:0], np.zeros((2, 5)))
    assert np.array_equal(out, np.zeros(5))
    # x = 0: cos = 1, sin = 0 -> sum of z1
    assert abs(oracle_mod.summate(cov, z1, z2, np.zeros((2, 1)))[0] - z1.sum()) < 1e-13
    with pytest.raises(ValueError):
        oracle_mod.summate(cov, z1[:-1], z2, np.zeros((2, 1)))


def test_live_reference_matches_fixtures(oracle_mod):
    """When the reference is importable (build container, or baseline/_ref on the GPU box) rerun
    one generator live and compare with the committed fixture."""
    import refharness

    if not refharness.have_reference():
        pytest.skip("reference gstools not present")
    gs = refharness.import_gstools()
    from gstools.field.generator import RandMeth

    meta, d = load_golden("randmeth_3d")
    rm = RandMeth(gs.Gaussian(dim=3, var=1.5, len_scale=3.5), mode_no=100, seed=19031977)
    assert np.array_equal(rm._cov_sample, d["cov_samples"])
    assert np.array_equal(rm._z_1, d["z_1"])
    x = np.linspace(0.0, 10.0, 10)
    y = np.linspace(-5.0, 5.0, 10)
    z = np.linspace(-6.0, 8.0, 10)
    assert np.allclose(rm((x, y, z)), d["field"], rtol=0, atol=1e-14)


# ---------------------------------------------------------------------------------------------
# kriging evaluation oracle (row f1): fixtures from the reference's Krige classes
# (tests/golden/make_golden_krige.py) and the known-answer facts of tests/test_krige.py
# ---------------------------------------------------------------------------------------------
def _krige_fixtures():
    import glob
    import os

    from conftest import GOLDEN_DIR

    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "krige", "*.npz")))


@pytest.mark.parametrize("path", _krige_fixtures(), ids=lambda p: p.split("/")[-1][:-4])
def test_krige_oracle_reproduces_recorded_boundary(path, oracle_mod):
    import json

    d = np.load(path)
    meta = json.loads(str(d["meta"]))
    field, error = oracle_mod.calc_field_krige_and_variance(d["krig_mat"], d["krig_vecs"], d["cond"])
    assert np.array_equal(field, d["field"]) and np.array_equal(error, d["error"])
    assert np.array_equal(oracle_mod.calc_field_krige(d["krig_mat"], d["krig_vecs"], d["cond"]), field)
    # numpy restatement (BLAS order): agreement to the rounding scale of the system
    f_np, e_np = oracle_mod.calc_field_krige_and_variance_np(d["krig_mat"], d["krig_vecs"], d["cond"])
    am, akv = np.abs(d["krig_mat"]), np.abs(d["krig_vecs"])
    eps = np.finfo(float).eps
    k = am.shape[0]
    assert np.all(np.abs(f_np - field) <= 8 * k * eps * (np.abs(d["cond"]) @ (am @ akv)) + 1e-300)
    assert np.all(np.abs(e_np - error) <= 8 * k * eps * np.einsum("ij,ij->j", akv, am @ akv) + 1e-300)
    # known answer asserted by the reference (tests/test_krige.py:76-79, 104-107): the kriged field
    # reproduces the conditioning values at the conditioning nodes, places=2
    if not meta["name"].startswith("universal"):
        mean = float(np.mean(meta["cond_val"])) if meta["name"].startswith("simple") else 0.0
        for idx, val in zip(meta["node_index"], meta["cond_val"]):
            assert round(field[idx] + mean - val, 2) == 0, meta["cite"]
        # ... and the error variance vanishes there (sill - error == 0 at exact data, base.py:296-298)
        assert np.all(np.abs(meta["sill"] - error[meta["node_index"]]) < 1e-6 * max(1.0, np.abs(d["krig_mat"]).max()))
