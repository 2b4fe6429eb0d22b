"""Row f1 on the GPU: kriging evaluation (calc_field_krige[_and_variance]) through the C ABI against
the CPU oracle (oracle/krige_oracle.c) and the fixtures recorded from the reference's Krige classes.

Tolerance.  The kernel regroups the reference's sums exactly (triangular quadratic form, w = M^T cond),
so only rounding differs.  Both orders carry an error of a few K*eps*S with S = sum_ij |a_i M_ij b_j|;
the bar is  |delta| <= max(1e-9 * unit, 8*K*eps*S)  with unit = sqrt(var) for the field and var for
the error variance (north_star tolerance where the system is well conditioned, the rounding scale of
the reference's own loop where it is not -- e.g. the ordinary/Gaussian/1-D system with |M| ~ 2.5e6)."""
import glob
import json
import os

import numpy as np
import pytest

import refharness
from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def bounds(mat, kv, cond, unit_f=1.0, unit_e=1.0):
    k = mat.shape[0]
    am, akv = np.abs(mat), np.abs(kv)
    s_f = np.abs(cond) @ (am @ akv)
    s_e = np.einsum("ij,ij->j", akv, am @ akv)
    return (np.maximum(1e-9 * unit_f, 8 * k * EPS * s_f), np.maximum(1e-9 * unit_e, 8 * k * EPS * s_e))


def check(gsb, oracle_mod, mat, kv, cond, unit_f=1.0, unit_e=1.0):
    f, e = gsb.calc_field_krige_and_variance(mat, kv, cond)
    wf, we = oracle_mod.calc_field_krige_and_variance(mat, kv, cond)
    tf, te = bounds(mat, np.asarray(kv), cond, unit_f, unit_e)
    assert f.shape == wf.shape and e.shape == we.shape and f.dtype == np.float64
    assert np.all(np.abs(f - wf) <= tf), float(np.max(np.abs(f - wf) / tf))
    assert np.all(np.abs(e - we) <= te), float(np.max(np.abs(e - we) / te))
    f2 = gsb.calc_field_krige(mat, kv, cond)
    assert np.all(np.abs(f2 - wf) <= tf)
    return f, e


def krige_fixtures():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "krige", "*.npz")))


@pytest.mark.parametrize("name", krige_fixtures())
def test_krige_golden_boundary_arrays(name, gsb, oracle_mod):
    d = np.load(os.path.join(GOLDEN_DIR, "krige", name + ".npz"))
    meta = json.loads(str(d["meta"]))
    f, e = gsb.calc_field_krige_and_variance(d["krig_mat"], d["krig_vecs"], d["cond"])
    tf, te = bounds(d["krig_mat"], d["krig_vecs"], d["cond"], np.sqrt(meta["var"]), meta["var"])
    assert np.all(np.abs(f - d["field"]) <= tf) and np.all(np.abs(e - d["error"]) <= te)
    # the reference's own known-answer check: the kriged field reproduces the conditioning values at
    # the conditioning nodes to 2 places (tests/test_krige.py:76-79, 104-107)
    mean = 0.0
    if name.startswith("simple"):
        mean = float(np.mean(meta["cond_val"]))     # Simple kriging adds the mean back (post-processing)
    if not name.startswith("universal"):
        for idx, val in zip(meta["node_index"], meta["cond_val"]):
            assert round(f[idx] + mean - val, 2) == 0, meta["cite"]


@pytest.mark.parametrize("K", [1, 5, 16, 127, 128, 129, 300, 1001])
@pytest.mark.parametrize("n", [1, 2, 127, 130, 1000, 4097])
def test_krige_random_systems_vs_oracle(K, n, gsb, oracle_mod):
    rs = np.random.RandomState(K * 7919 + n)
    mat = rs.normal(size=(K, K))
    mat = mat + 0.25 * rs.normal(size=(K, K))       # NOT symmetric: the regrouping must not assume it
    kv = rs.uniform(-1, 1, (K, n))
    cond = rs.normal(size=K)
    check(gsb, oracle_mod, mat, kv, cond)


def test_krige_strided_and_device_inputs(gsb, oracle_mod):
    import torch

    rs = np.random.RandomState(11)
    K, n = 200, 3001
    mat, cond = rs.normal(size=(K, K)), rs.normal(size=K)
    big = rs.uniform(-1, 1, (K, 5000))
    for a, b in [(0, n), (3, 3 + n), (128, 128 + 3000), (1, 2)]:
        kv = big[:, a:b]                              # row-strided views, odd offsets, odd widths
        f, e = check(gsb, oracle_mod, mat, kv, cond)
        t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
        fd, ed = gsb.calc_field_krige_and_variance(t(mat), t(big)[:, a:b], t(cond))
        assert fd.is_cuda and np.array_equal(fd.cpu().numpy(), f) and np.array_equal(ed.cpu().numpy(), e)
        assert np.array_equal(gsb.calc_field_krige(t(mat), t(big)[:, a:b], t(cond)).cpu().numpy(),
                              gsb.calc_field_krige(mat, kv, cond))


def test_krige_host_chunking_is_invisible(gsb, oracle_mod):
    rs = np.random.RandomState(5)
    K, n = 130, 20001
    mat, cond, kv = rs.normal(size=(K, K)), rs.normal(size=K), rs.uniform(-1, 1, (K, n))
    ref = gsb.calc_field_krige_and_variance(mat, kv, cond)
    gsb.set_option("krige_host_chunk_mb", 1)         # ~ 900 columns per chunk
    try:
        got = gsb.calc_field_krige_and_variance(mat, kv, cond)
        got_f = gsb.calc_field_krige(mat, kv, cond)
    finally:
        gsb.set_option("krige_host_chunk_mb", 256)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])   # a point's bits do not depend on the chunking
    check(gsb, oracle_mod, mat, kv, cond)
    assert np.max(np.abs(got_f - ref[0])) < 1e-9


def test_krige_edge_cases_and_errors(gsb):
    f, e = gsb.calc_field_krige_and_variance(np.zeros((3, 3)), np.zeros((3, 0)), np.zeros(3))
    assert f.shape == (0,) and e.shape == (0,)
    f, e = gsb.calc_field_krige_and_variance(np.zeros((0, 0)), np.zeros((0, 4)), np.zeros(0))
    assert np.array_equal(f, np.zeros(4)) and np.array_equal(e, np.zeros(4))
    with pytest.raises(ValueError):
        gsb.calc_field_krige_and_variance(np.zeros((3, 2)), np.zeros((3, 4)), np.zeros(3))
    with pytest.raises(ValueError):
        gsb.calc_field_krige_and_variance(np.zeros((3, 3)), np.zeros((2, 4)), np.zeros(3))
    with pytest.raises(ValueError):
        gsb.calc_field_krige(np.zeros((3, 3)), np.zeros((3, 4)), np.zeros(2))
    # NaN / inf in one point's right-hand side stay in that point
    rs = np.random.RandomState(0)
    mat, cond, kv = rs.normal(size=(40, 40)), rs.normal(size=40), rs.normal(size=(40, 300))
    clean = gsb.calc_field_krige_and_variance(mat, kv, cond)
    kv2 = kv.copy()
    kv2[7, 131] = np.nan
    f, e = gsb.calc_field_krige_and_variance(mat, kv2, cond)
    assert np.isnan(f[131]) and np.isnan(e[131])
    keep = np.arange(300) != 131
    assert np.array_equal(f[keep], clean[0][keep]) and np.array_equal(e[keep], clean[1][keep])


# ---------------------------------------------------------------------------------------------
# through the unmodified reference
# ---------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not refharness.have_reference(), reason="reference gstools not present")


@needs_ref
def test_reference_krige_and_condsrf_on_gpu(gsb):
    """tests/test_krige.py:81-107 (ordinary) and tests/test_condition.py:62-100 with the backend on:
    kriged / conditioned fields honour the conditioning values; the kriging evaluation ran on the GPU."""
    gs = refharness.import_gstools()
    data = np.array([[0.3, 1.2, 0.5, 0.47], [1.9, 0.6, 1.0, 0.56], [1.1, 3.2, 1.5, 0.74],
                     [3.3, 4.4, 2.0, 1.47], [4.7, 3.8, 2.5, 1.74]])
    cond_pos, cond_val = (data[:, 0], data[:, 1], data[:, 2]), data[:, 3]
    pos = (np.linspace(0, 5, 51), np.linspace(0, 6, 61), np.linspace(0, 7, 71))
    data_idx = tuple(np.array(data[:, :3] * 10, dtype=int).T)
    gsb.enable()
    try:
        calls = gsb.get_counter("krige_calls")
        for Model in (gs.Gaussian, gs.Exponential, gs.Spherical):
            for dim in (1, 2, 3):
                model = Model(dim=dim, var=5, len_scale=10, anis=[0.9, 0.8], angles=[2, 1, 0.5])
                ordinary = gs.krige.Ordinary(model, cond_pos[:dim], cond_val)
                field, var = ordinary.structured(pos[:dim])
                for i, val in enumerate(cond_val):
                    assert round(field[data_idx[:dim]][i] - val, 2) == 0
                assert np.all(var >= 0)
        assert gsb.get_counter("krige_calls") > calls
        grid = np.linspace(5, 20, 10)
        cpos = tuple(np.concatenate((c, grid)) for c in cond_pos)
        for Model in (gs.Gaussian, gs.Exponential):
            model = Model(dim=3, var=0.5, len_scale=2, anis=[0.1, 1], angles=[0.5, 0, 0])
            krige = gs.krige.Ordinary(model, cond_pos, cond_val)
            crf = gs.CondSRF(krige, seed=19970221)
            f1 = crf.unstructured(cpos)
            f2 = crf.structured(cpos)
            for i, val in enumerate(cond_val):
                assert round(f1[i] - val, 2) == 0 and round(f2[(i, i, i)] - val, 2) == 0
    finally:
        gsb.disable()


# ---------------------------------------------------------------------------------------------
# krige_evaluate: right-hand sides generated on the device (Krige._get_krige_vecs, base.py:359-388)
# ---------------------------------------------------------------------------------------------
def _eval_bounds(mat, kv, cond, unit_f, unit_e):
    """Rounding bound as above plus the effect of a 4-ulp perturbation of every right-hand-side
    entry (device exp / pow / acos vs libm), propagated through the same sums."""
    tf, te = bounds(mat, kv, cond, unit_f, unit_e)
    am, akv = np.abs(mat), np.abs(kv)
    return tf + 8 * EPS * (np.abs(cond) @ (am @ akv)), te + 16 * EPS * np.einsum("ij,ij->j", akv, am @ akv)


@pytest.mark.parametrize("name", krige_fixtures())
def test_krige_evaluate_fixtures(name, gsb, oracle_mod):
    d = np.load(os.path.join(GOLDEN_DIR, "krige", name + ".npz"))
    meta = json.loads(str(d["meta"]))
    spec = gsb.cov_model_spec(**meta["spec"])
    tail = d["tail_rows"] if d["tail_rows"].size else None
    tf, te = _eval_bounds(d["krig_mat"], d["krig_vecs"], d["cond"], np.sqrt(meta["var"]), meta["var"])
    f, e = gsb.krige_evaluate(spec, d["krig_mat"], d["cond"], d["cond_pos_iso"], pos=d["pos_iso"],
                              unbiased=meta["unbiased"], tail_rows=tail)
    assert np.all(np.abs(f - d["field"]) <= tf) and np.all(np.abs(e - d["error"]) <= te)
    axes = [d[f"axis{t}"] for t in range(len(meta["shape"]))]
    fs, es = gsb.krige_evaluate(spec, d["krig_mat"], d["cond"], d["cond_pos_iso"], axes=axes, matrix=d["matrix"],
                                unbiased=meta["unbiased"], tail_rows=tail)
    assert fs.shape == tuple(meta["shape"])
    # the mesh is isometrised on the device: positions differ by rounding from the reference's np.dot
    slack = 1e-12 * (np.abs(d["cond"]) @ (np.abs(d["krig_mat"]) @ np.abs(d["krig_vecs"])))
    assert np.all(np.abs(fs.reshape(-1) - d["field"]) <= tf + slack)
    f_only = gsb.krige_evaluate(spec, d["krig_mat"], d["cond"], d["cond_pos_iso"], pos=d["pos_iso"],
                                unbiased=meta["unbiased"], tail_rows=tail, return_var=False)
    assert np.all(np.abs(f_only - d["field"]) <= tf)
    # the kriging variance the reference forms from it (base.py:296-298)
    assert np.all(np.abs(np.maximum(meta["sill"] - e, 0).reshape(meta["shape"]) - d["krige_var"]) <= te.reshape(meta["shape"]))


@pytest.mark.parametrize("kind,param", [("Gaussian", 0.0), ("Exponential", 0.0), ("Stable", 1.3), ("Rational", 0.8),
                                        ("Cubic", 0.0), ("Linear", 0.0), ("Circular", 0.0), ("Spherical", 0.0)])
@pytest.mark.parametrize("dim,C,n,exact", [(3, 300, 5001, False), (2, 129, 640, True), (1, 17, 130, False)])
def test_krige_evaluate_random_vs_oracle(kind, param, dim, C, n, exact, gsb, oracle_mod):
    rs = np.random.RandomState(C + n)
    cond_pos = rs.uniform(0, 50, (dim, C))
    pos = rs.uniform(0, 50, (dim, n))
    pos[:, :5] = cond_pos[:, :5]                       # evaluation points ON conditioning points (r = 0)
    K = C + 1 + 2                                      # unbiased row + two drift rows
    mat = rs.normal(size=(K, K)) / K
    cond = np.concatenate([rs.normal(size=C), np.zeros(3)])
    tail = rs.normal(size=(2, n))
    spec = dict(kind=kind, var=1.7, len_rescaled=9.0, sill=1.9, param=param, exact=exact)
    kv = oracle_mod.krige_vecs_np(kind, 1.7, 9.0, 1.9, cond_pos, pos, True, tail, param, exact)
    wf, we = oracle_mod.calc_field_krige_and_variance(mat, kv, cond)
    tf, te = _eval_bounds(mat, kv, cond, 1.0, 1.0)
    f, e = gsb.krige_evaluate(spec, mat, cond, cond_pos, pos=pos, unbiased=True, tail_rows=tail)
    assert np.all(np.abs(f - wf) <= tf), float(np.max(np.abs(f - wf) / tf))
    assert np.all(np.abs(e - we) <= te), float(np.max(np.abs(e - we) / te))
    f2 = gsb.krige_evaluate(spec, mat, cond, cond_pos, pos=pos, unbiased=True, tail_rows=tail, return_var=False)
    assert np.all(np.abs(f2 - wf) <= tf)


def test_krige_evaluate_chunking_is_invisible(gsb):
    rs = np.random.RandomState(2)
    C, n = 100, 3000
    cond_pos, pos = rs.uniform(0, 30, (3, C)), rs.uniform(0, 30, (3, n))
    mat, cond = rs.normal(size=(C + 1, C + 1)) / C, np.concatenate([rs.normal(size=C), [0.0]])
    spec = dict(kind="Exponential", var=1.0, len_rescaled=5.0)
    ref = gsb.krige_evaluate(spec, mat, cond, cond_pos, pos=pos)
    gsb.set_option("scratch_mb", 1)                   # a few column tiles per chunk
    try:
        got = gsb.krige_evaluate(spec, mat, cond, cond_pos, pos=pos)
    finally:
        gsb.set_option("scratch_mb", 3072)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    with pytest.raises(ValueError):
        gsb.krige_evaluate(dict(kind="Matern", var=1.0, len_rescaled=5.0), mat, cond, cond_pos, pos=pos)
    with pytest.raises(ValueError):
        gsb.krige_evaluate(spec, mat, cond, cond_pos, pos=pos, unbiased=False)     # a drift row is missing


@needs_ref
@pytest.mark.parametrize("mesh", ["structured", "unstructured"])
def test_fast_krige_and_condsrf_equal_reference_path(gsb, mesh):
    """gs.krige.* and gs.CondSRF through the fast Krige.__call__ (device right-hand sides) against
    the same objects on the reference's own chunk loop (host right-hand sides, CPU oracle)."""
    gs = refharness.import_gstools()
    rs = np.random.RandomState(20170519)
    cond_pos = rs.uniform(0, 31, (3, 60))
    cond_val = rs.normal(size=60)
    model = gs.Exponential(dim=3, var=1.3, len_scale=[10, 6, 4], angles=[0.3, 0.2, 0.1])
    if mesh == "structured":
        pos = [np.arange(32.0), np.arange(30.0), np.arange(17.0)]
    else:
        pos = rs.uniform(0, 31, (3, 4000))
    ref_k = gs.krige.Ordinary(model, cond_pos, cond_val)
    want_f, want_v = ref_k(pos, mesh_type=mesh)
    want_c = gs.CondSRF(gs.krige.Ordinary(model, cond_pos, cond_val), seed=4711, mode_no=128)(pos, mesh_type=mesh)
    gsb.enable()
    try:
        calls = gsb.get_counter("krige_calls")
        got_f, got_v = gs.krige.Ordinary(model, cond_pos, cond_val)(pos, mesh_type=mesh)
        got_c = gs.CondSRF(gs.krige.Ordinary(model, cond_pos, cond_val), seed=4711, mode_no=128)(pos, mesh_type=mesh)
        assert gsb.get_counter("krige_calls") >= calls + 2
    finally:
        gsb.disable()
    assert got_f.shape == want_f.shape
    assert np.max(np.abs(got_f - want_f)) <= 1e-9 * np.sqrt(model.var)
    assert np.max(np.abs(got_v - want_v)) <= 1e-9 * model.var
    assert np.max(np.abs(got_c - want_c)) <= 1e-8 * np.sqrt(model.var)   # sqrt(krige_var) near data amplifies


def test_krige_evaluate_device_tensors(gsb):
    """CUDA tensors in, CUDA tensors out: the same bits as the host-array call (flat and mesh, with a
    drift row, field only)."""
    import torch

    rs = np.random.RandomState(12)
    C, n = 70, 2600
    cond_pos, pos = rs.uniform(0, 25, (3, C)), rs.uniform(0, 25, (3, n))
    K = C + 2
    mat, cond, tail = rs.normal(size=(K, K)) / K, np.concatenate([rs.normal(size=C), [0.0, 0.0]]), rs.normal(size=(1, n))
    spec = dict(kind="Spherical", var=1.4, len_rescaled=9.0, sill=1.5, exact=True)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    hf, he = gsb.krige_evaluate(spec, mat, cond, cond_pos, pos=pos, tail_rows=tail)
    df, de = gsb.krige_evaluate(spec, t(mat), t(cond), t(cond_pos), pos=t(pos), tail_rows=t(tail))
    assert df.is_cuda and np.array_equal(df.cpu().numpy(), hf) and np.array_equal(de.cpu().numpy(), he)
    axes = [np.arange(10.0), np.linspace(0, 20, 13), np.linspace(0, 25, 20)]
    mtx = np.array([[0.9, 0.1, 0.0], [-0.1, 1.1, 0.0], [0.0, 0.2, 2.0]])
    mesh_tail = rs.normal(size=(1, 10 * 13 * 20))
    hf, he = gsb.krige_evaluate(spec, mat, cond, cond_pos, axes=axes, matrix=mtx, tail_rows=mesh_tail)
    df, de = gsb.krige_evaluate(spec, t(mat), t(cond), t(cond_pos), axes=[t(a) for a in axes], matrix=mtx,
                                tail_rows=t(mesh_tail))
    assert tuple(df.shape) == (10, 13, 20)
    assert np.array_equal(df.cpu().numpy(), hf) and np.array_equal(de.cpu().numpy(), he)
    f_only = gsb.krige_evaluate(spec, t(mat), t(cond), t(cond_pos), pos=t(pos), tail_rows=t(tail), return_var=False)
    assert np.array_equal(f_only.cpu().numpy(), gsb.krige_evaluate(spec, mat, cond, cond_pos, pos=pos, tail_rows=tail,
                                                                   return_var=False))


def test_field_only_evaluation_of_a_large_system(gsb, oracle_mod):
    """Krige.__call__(return_var=False) with more conditioning points than the field-only kernel
    could stage in shared memory at once (C*D + K doubles > 200 KB) used to fail; the rows are now staged in
    passes.  7000 conditioning points in 3-D, synthetic (well-scaled) weights."""
    rs = np.random.RandomState(4)
    C, D, n = 7000, 3, 3000
    K = C + 1
    cond_pos = rs.uniform(0, 50, (D, C))
    pos = rs.uniform(0, 50, (D, n))
    mat = rs.normal(size=(K, K)) / K
    cond = rs.normal(size=K)
    spec = dict(kind="Exponential", var=1.3, len_rescaled=7.0)
    f = gsb.krige_evaluate(spec, mat, cond, cond_pos, pos=pos, return_var=False)
    kv = oracle_mod.krige_vecs_np("Exponential", 1.3, 7.0, 1.3, cond_pos, pos)
    want = oracle_mod.calc_field_krige(mat, kv, cond)
    s = np.abs(cond) @ (np.abs(mat) @ np.abs(kv))
    assert np.all(np.abs(f - want) <= np.maximum(1e-9, 8 * K * EPS * s))
