"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden
fixtures.  Tolerance from BASELINE.json north_star: max|delta| <= 1e-9 * sqrt(var) on the scaled
field, i.e. max|delta_raw| <= 1e-9 * sqrt(N) on the raw sums (field = sqrt(var/N) * raw)."""
import ctypes

import numpy as np
import pytest

from conftest import golden_names, load_golden, synth_modes

pytestmark = pytest.mark.gpu

TOL = 1e-9  # relative to sqrt(var), see module docstring


def raw_tol(n_modes):
    return TOL * np.sqrt(max(n_modes, 1))


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) if np.size(a) else 0.0


# ---------------------------------------------------------------------------------------------
# golden fixtures (recorded from the reference's own generators)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_golden_boundary_arrays(name, gsb):
    meta, d = load_golden(name)
    if meta["kind"] == "fourier":
        got = gsb.summate_fourier(d["spectrum_factor"], d["cov_samples"], d["z_1"], d["z_2"], d["pos"])
        assert maxabs(got, d["raw"]) <= TOL * np.sqrt(meta["var"])   # spectrum factor carries the scale
        axes = [d[f"axis{t}"] for t in range(d["cov_samples"].shape[0])]
        for force in (1, 2):
            gsb.set_option("force_path", force)
            try:
                st = gsb.summate_fourier_structured(d["spectrum_factor"], d["cov_samples"], d["z_1"], d["z_2"], axes)
            finally:
                gsb.set_option("force_path", 0)
            assert st.shape == d["field"].shape
            assert maxabs(st, d["field"]) <= TOL * np.sqrt(meta["var"])
            for a in meta["asserts"]:
                assert round(st[tuple(a["index"])] - a["value"], a["places"]) == 0, a["cite"]
        return
    fn = gsb.summate if meta["kind"] == "scalar" else gsb.summate_incompr
    got = fn(d["cov_samples"], d["z_1"], d["z_2"], d["pos"])
    assert got.shape == d["raw"].shape and got.dtype == np.float64
    assert maxabs(got, d["raw"]) <= raw_tol(d["cov_samples"].shape[1])


@pytest.mark.parametrize("name", ["randmeth_1d", "randmeth_2d", "randmeth_3d", "randmeth_2d_reseed",
                                  "randmeth_2d_modes800", "incompr_2d", "incompr_3d"])
def test_reference_test_literals_on_gpu(name, gsb):
    """The literals asserted in the reference's tests (tests/test_randmeth.py:33-71,
    tests/test_incomprrandmeth.py:34-44), same 7-decimal criterion, computed on the GPU."""
    meta, d = load_golden(name)
    fn = gsb.summate if meta["kind"] == "scalar" else gsb.summate_incompr
    n_modes = d["cov_samples"].shape[1]
    field = np.sqrt(meta["var"] / n_modes) * fn(d["cov_samples"], d["z_1"], d["z_2"], d["pos"])
    if meta["kind"] == "incompr":
        field[0] += 1.0  # mean_u * e1, generator.py:561-567
    for a in meta["asserts"]:
        assert round(field[tuple(a["index"])] - a["value"], a["places"]) == 0, a["cite"]


# ---------------------------------------------------------------------------------------------
# random inputs against the oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("n,n_modes", [(1, 1), (3, 7), (64, 128), (65, 129), (1000, 100),
                                       (18945, 260), (75777, 33), (303105, 16)])
def test_direct_scalar_vs_oracle(dim, n, n_modes, gsb, oracle_mod):
    cov, z1, z2 = synth_modes(dim, n_modes, seed=dim * 1000 + n_modes)
    pos = np.random.RandomState(n).uniform(-100, 500, (dim, n))
    got = gsb.summate(cov, z1, z2, pos)
    want = oracle_mod.summate(cov, z1, z2, pos)
    assert maxabs(got, want) <= raw_tol(n_modes)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("n,n_modes", [(1, 1), (10, 100), (257, 130), (20000, 1000), (303105, 12)])
def test_direct_incompr_vs_oracle(dim, n, n_modes, gsb, oracle_mod):
    cov, z1, z2 = synth_modes(dim, n_modes, seed=dim * 77 + n_modes)
    pos = np.random.RandomState(n).uniform(-100, 500, (dim, n))
    got = gsb.summate_incompr(cov, z1, z2, pos)
    want = oracle_mod.summate_incompr(cov, z1, z2, pos)
    assert got.shape == (dim, n)
    assert maxabs(got, want) <= raw_tol(n_modes)


@pytest.mark.parametrize("dim,n_modes,n", [(3, 5000, 200), (2, 1000, 10000), (3, 130, 7), (1, 64, 33)])
def test_mode_split_path_small_point_sets(gsb, oracle_mod, dim, n_modes, n):
    """Small point sets (config 1: 10 000 points) split the mode loop over CTAs; every CTA stores the sums of its
    64-mode tiles and the reduce kernel adds them in the unsplit kernel's order: the SAME bits as without the split and as
    the same points inside a big call."""
    cov, z1, z2 = synth_modes(dim, n_modes, seed=3)
    pos = np.random.RandomState(0).uniform(0, 100, (dim, n))
    before = gsb.get_counter("launches")
    a = gsb.summate(cov, z1, z2, pos)
    split_ran = gsb.get_counter("launches") - before == 3            # records + direct + reduce
    assert split_ran == (n_modes > 64)
    assert maxabs(a, oracle_mod.summate(cov, z1, z2, pos)) <= raw_tol(n_modes)
    gsb.set_option("direct_split", 0)
    try:
        assert np.array_equal(gsb.summate(cov, z1, z2, pos), a)
        if dim in (2, 3):
            unsplit_v = gsb.summate_incompr(cov, z1, z2, pos)
    finally:
        gsb.set_option("direct_split", 1)
    if dim in (2, 3):
        av = gsb.summate_incompr(cov, z1, z2, pos)
        assert np.array_equal(av, unsplit_v)
        assert maxabs(av, oracle_mod.summate_incompr(cov, z1, z2, pos)) <= raw_tol(n_modes)
    # the same points as the head of a call big enough for the 8-points-per-thread configuration
    big = np.concatenate([pos, np.random.RandomState(1).uniform(0, 100, (dim, 700000))], axis=1)
    assert np.array_equal(gsb.summate(cov[:, :256], z1[:256], z2[:256], big)[:n],
                          gsb.summate(cov[:, :256], z1[:256], z2[:256], pos))


def test_edge_cases(gsb, oracle_mod):
    cov, z1, z2 = synth_modes(2, 10, seed=1)
    # no modes: zeros (the reference returns np.zeros + nothing added)
    out = gsb.summate(cov[:, :0], z1[:0], z2[:0], np.ones((2, 5)))
    assert np.array_equal(out, np.zeros(5))
    out = gsb.summate_incompr(cov[:, :0], z1[:0], z2[:0], np.ones((2, 5)))
    assert np.array_equal(out, np.zeros((2, 5)))
    # x = 0 -> sum(z1), up to the documented abs error of the cos polynomial (4.3e-13 per mode,
    # gstools_b200/csrc/sincos_coeffs.cuh: the Chebyshev interpolant is not pinned to 1 at r = 0)
    assert abs(gsb.summate(cov, z1, z2, np.zeros((2, 1)))[0] - z1.sum()) < 5e-13 * np.abs(z1).sum()
    # a zero wave vector gives NaN in the projector, exactly like the reference loop (k2 == 0)
    cov0 = cov.copy()
    cov0[:, 3] = 0.0
    got = gsb.summate_incompr(cov0, z1, z2, np.ones((2, 4)))
    want = oracle_mod.summate_incompr(cov0, z1, z2, np.ones((2, 4)))
    assert np.isnan(got).all() and np.isnan(want).all()
    # scalar sum is unaffected by a zero wave vector
    assert maxabs(gsb.summate(cov0, z1, z2, np.ones((2, 4))),
                  oracle_mod.summate(cov0, z1, z2, np.ones((2, 4)))) <= raw_tol(10)


def test_input_coercion_like_the_reference(gsb, oracle_mod):
    cov, z1, z2 = synth_modes(3, 50, seed=2)
    pos = np.random.RandomState(1).uniform(0, 50, (3, 333))
    want = oracle_mod.summate(cov, z1, z2, pos)
    # Fortran order, strided views, lists, read-only arrays, float32 / int positions
    assert maxabs(gsb.summate(np.asfortranarray(cov), z1, z2, np.asfortranarray(pos)), want) <= raw_tol(50)
    big = np.zeros((3, 1000))
    big[:, 100:433] = pos
    assert maxabs(gsb.summate(cov, z1, z2, big[:, 100:433]), want) <= raw_tol(50)
    assert maxabs(gsb.summate(cov.tolist(), list(z1), list(z2), pos.tolist()), want) <= raw_tol(50)
    ro = pos.copy()
    ro.setflags(write=False)
    assert maxabs(gsb.summate(cov, z1, z2, ro), want) <= raw_tol(50)
    ipos = np.arange(30).reshape(3, 10)
    assert maxabs(gsb.summate(cov, z1, z2, ipos), oracle_mod.summate(cov, z1, z2, ipos.astype(float))) <= raw_tol(50)
    # inputs are never modified
    c0, p0 = cov.copy(), pos.copy()
    gsb.summate(cov, z1, z2, pos)
    assert np.array_equal(cov, c0) and np.array_equal(pos, p0)
    # num_threads is accepted and ignored (reference call passes 5 positional args)
    assert np.array_equal(gsb.summate(cov, z1, z2, pos, 4), gsb.summate(cov, z1, z2, pos, None))


def test_large_phases(gsb, oracle_mod):
    """UTM-like coordinates and heavy-tailed wave numbers: |phase| up to ~1e6 rad.  Both
    implementations carry |phase| * 1e-16 of rounding, far inside the tolerance."""
    cov, z1, z2 = synth_modes(2, 1000, seed=5, len_scale=10.0)
    pos = np.random.RandomState(2).uniform(4.0e5, 4.1e5, (2, 2000))
    got = gsb.summate(cov, z1, z2, pos)
    want = oracle_mod.summate(cov, z1, z2, pos)
    assert maxabs(got, want) <= raw_tol(1000)


def test_bitwise_determinism_and_linearity(gsb):
    cov, z1, z2 = synth_modes(3, 300, seed=8)
    pos = np.random.RandomState(3).uniform(0, 300, (3, 50000))
    a = gsb.summate(cov, z1, z2, pos)
    assert np.array_equal(a, gsb.summate(cov, z1, z2, pos))
    # exact sign symmetry and exact scaling by a power of two
    assert np.array_equal(-a, gsb.summate(cov, -z1, -z2, pos))
    assert np.array_equal(4.0 * a, gsb.summate(cov, 4.0 * z1, 4.0 * z2, pos))
    # additivity in the weights
    rs = np.random.RandomState(4)
    y1, y2 = rs.normal(size=300), rs.normal(size=300)
    b = gsb.summate(cov, y1, y2, pos)
    c = gsb.summate(cov, z1 + y1, z2 + y2, pos)
    assert maxabs(a + b, c) <= raw_tol(300)
    # a point's value does not depend on which other points are in the call
    sub = gsb.summate(cov, z1, z2, pos[:, 1234:1300])
    assert np.array_equal(sub, a[1234:1300])


def test_device_tensor_path_matches_host_path(gsb):
    import torch

    cov, z1, z2 = synth_modes(3, 200, seed=9)
    pos = np.random.RandomState(5).uniform(0, 300, (3, 70001))
    host = gsb.summate(cov, z1, z2, pos)
    dev = torch.device("cuda:0")
    tc, t1, t2, tp = (torch.tensor(a, device=dev) for a in (cov, z1, z2, pos))
    out = gsb.summate(tc, t1, t2, tp)
    assert out.is_cuda and out.dtype == torch.float64 and out.shape == (70001,)
    assert np.array_equal(out.cpu().numpy(), host)
    # mixed: numpy modes + CUDA positions, and a strided point range without a copy
    out2 = gsb.summate(cov, z1, z2, tp[:, 1000:3000])
    assert np.array_equal(out2.cpu().numpy(), host[1000:3000])
    v = gsb.summate_incompr(tc, t1, t2, tp)
    assert np.array_equal(v.cpu().numpy(), gsb.summate_incompr(cov, z1, z2, pos))
    # fused epilogue on device: sqrt(var/N) * s + shift (generator.py:269-270)
    f = out.clone()
    gsb.scale_shift_(f, 0.25, 1.5)
    assert np.array_equal(f.cpu().numpy(), 0.25 * host + 1.5)
    # non-default stream
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out3 = gsb.summate(tc, t1, t2, tp)
    s.synchronize()
    assert np.array_equal(out3.cpu().numpy(), host)


def test_raw_c_abi_call_with_host_pointers(gsb, oracle_mod):
    """Call the exported symbol directly, the way a foreign-language binding would."""
    lib = ctypes.CDLL(gsb._lib.lib_path())
    lib.gsb_summate.restype = ctypes.c_int
    lib.gsb_last_error.restype = ctypes.c_char_p
    cov, z1, z2 = synth_modes(2, 64, seed=4)
    pos = np.ascontiguousarray(np.random.RandomState(6).uniform(0, 100, (2, 999)))
    out = np.empty(999)
    vp = ctypes.c_void_p
    rc = lib.gsb_summate(vp(cov.ctypes.data), vp(z1.ctypes.data), vp(z2.ctypes.data),
                         vp(pos.ctypes.data), ctypes.c_int64(999), ctypes.c_int(2),
                         ctypes.c_int64(64), ctypes.c_int64(999), vp(out.ctypes.data),
                         ctypes.c_int(0), ctypes.c_int(0), vp(None))
    assert rc == 0, lib.gsb_last_error()
    assert maxabs(out, oracle_mod.summate(cov, z1, z2, pos)) <= raw_tol(64)


# ---------------------------------------------------------------------------------------------
# structured meshes: separable kernel and the small-mesh (expand + direct) route
# ---------------------------------------------------------------------------------------------
STRUCT_CASES = [
    (2, (100, 100), 1000),     # config 1 shape
    (2, (3, 5), 10),
    (2, (129, 131), 77),       # odd last axis: scalar stores
    (2, (300, 517), 200),
    (3, (40, 50, 130), 300),
    (3, (7, 9, 257), 64),
    (3, (64, 64, 256), 1000),
    (4, (9, 10, 11, 140), 64),
    (5, (3, 4, 5, 6, 130), 20),
]


@pytest.mark.parametrize("dim,lens,n_modes", STRUCT_CASES)
@pytest.mark.parametrize("path", ["direct", "separable"])
def test_structured_vs_oracle(dim, lens, n_modes, path, gsb, oracle_mod):
    cov, z1, z2 = synth_modes(dim, n_modes, seed=7 + dim)
    rs = np.random.RandomState(5)
    axes = [np.sort(rs.uniform(0, 200, L)) for L in lens]    # non-uniform axes
    mat = rs.normal(size=(dim, dim))                          # rotation + anisotropy + shear
    grid = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    pos = mat @ grid
    want = oracle_mod.summate(cov, z1, z2, pos).reshape(lens)
    gsb.set_option("force_path", 1 if path == "direct" else 2)
    try:
        got = gsb.summate_structured(cov, z1, z2, axes, mat)
        assert got.shape == tuple(lens)
        assert maxabs(got, want) <= raw_tol(n_modes)
        if dim in (2, 3):
            wv = oracle_mod.summate_incompr(cov, z1, z2, pos).reshape((dim,) + tuple(lens))
            gv = gsb.summate_incompr_structured(cov, z1, z2, axes, mat)
            assert gv.shape == (dim,) + tuple(lens)
            assert maxabs(gv, wv) <= raw_tol(n_modes)
    finally:
        gsb.set_option("force_path", 0)


def test_structured_identity_matrix_and_auto_path(gsb, oracle_mod):
    cov, z1, z2 = synth_modes(3, 100, seed=1)
    axes = [np.arange(64.0), np.arange(64.0), np.arange(200.0)]     # 32 x 2 = 64 tiles of 128x128
    grid = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    want = oracle_mod.summate(cov, z1, z2, grid).reshape(64, 64, 200)
    before = gsb.get_counter("separable_calls")
    got = gsb.summate_structured(cov, z1, z2, axes)           # matrix=None -> identity
    assert gsb.get_counter("separable_calls") == before + 1   # big enough for the tiled kernel
    assert maxabs(got, want) <= raw_tol(100)
    got_eye = gsb.summate_structured(cov, z1, z2, axes, np.eye(3))
    assert np.array_equal(got, got_eye)
    # tiny mesh: whichever path the cost model picks (a single stream-K share, or expanded + direct kernel)
    before = gsb.get_counter("direct_calls") + gsb.get_counter("separable_calls")
    small = gsb.summate_structured(cov, z1, z2, [np.arange(4.0), np.arange(5.0), np.arange(6.0)])
    assert gsb.get_counter("direct_calls") + gsb.get_counter("separable_calls") == before + 1
    assert maxabs(small, want[:4, :5, :6]) <= raw_tol(100)


def test_structured_batched_ensemble(gsb, oracle_mod):
    n_batch, dim, lens, n_modes = 5, 3, (20, 30, 130), 100
    sets = [synth_modes(dim, n_modes, seed=50 + b) for b in range(n_batch)]
    cov = np.stack([s[0] for s in sets])
    z1 = np.stack([s[1] for s in sets])
    z2 = np.stack([s[2] for s in sets])
    axes = [np.linspace(0, 50, L) for L in lens]
    grid = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    for force in (1, 2):
        gsb.set_option("force_path", force)
        try:
            got = gsb.summate_structured(cov, z1, z2, axes)
            assert got.shape == (n_batch,) + lens
            for b in range(n_batch):
                want = oracle_mod.summate(cov[b], z1[b], z2[b], grid).reshape(lens)
                assert maxabs(got[b], want) <= raw_tol(n_modes)
                # batching changes nothing beyond rounding (direct kernel: not a bit; the stream-K shares of
                # the separable contraction cut the mode sums of a tile at geometry-dependent stages)
                single = gsb.summate_structured(cov[b], z1[b], z2[b], axes)
                if force == 1:
                    assert np.array_equal(single, got[b])
                assert maxabs(single, got[b]) <= 1e-3 * raw_tol(n_modes)
        finally:
            gsb.set_option("force_path", 0)


def test_structured_device_tensors_and_slab_consistency(gsb):
    import torch

    cov, z1, z2 = synth_modes(3, 128, seed=3)
    axes = [np.arange(64.0), np.arange(96.0), np.arange(256.0)]
    host = gsb.summate_structured(cov, z1, z2, axes)
    dev = torch.device("cuda:0")
    out = gsb.summate_structured(torch.tensor(cov, device=dev), torch.tensor(z1, device=dev),
                                 torch.tensor(z2, device=dev),
                                 [torch.tensor(a, device=dev) for a in axes])
    assert out.is_cuda and tuple(out.shape) == (64, 96, 256)
    assert np.array_equal(out.cpu().numpy(), host)
    # a slab computed alone equals the same rows of the full field up to rounding (multi-GPU sharding unit:
    # the tiles are the same, the stream-K shares cut their mode sums at other stages)
    gsb.set_option("force_path", 2)
    try:
        part = gsb.summate_structured(cov, z1, z2, [axes[0][16:48], axes[1], axes[2]])
        again = gsb.summate_structured(cov, z1, z2, [axes[0][16:48], axes[1], axes[2]])
    finally:
        gsb.set_option("force_path", 0)
    assert np.array_equal(part, again)
    assert maxabs(part, host[16:48]) <= 1e-3 * raw_tol(128)


def test_separable_and_direct_kernels_agree(gsb):
    """Two different algorithms (2 DFMA/pair contraction vs per-pair polynomial sincos) on the
    same mesh."""
    cov, z1, z2 = synth_modes(3, 1000, seed=11)
    axes = [np.arange(96.0), np.arange(128.0), np.arange(384.0)]
    gsb.set_option("force_path", 2)
    sep = gsb.summate_structured(cov, z1, z2, axes)
    gsb.set_option("force_path", 1)
    try:
        direct = gsb.summate_structured(cov, z1, z2, axes)
    finally:
        gsb.set_option("force_path", 0)
    assert maxabs(sep, direct) <= raw_tol(1000)


def test_structured_routes_and_batches_agree(gsb, oracle_mod):
    """Vector fields, batches, host and device routes on a ragged mesh (1961 rows): all within rounding of each
    other and of the oracle; identical calls are bit-identical (fixed split, fixed summation order)."""
    import torch

    dev = torch.device("cuda:0")
    cov, z1, z2 = synth_modes(3, 96, seed=21)
    axes = [np.arange(37.0), np.linspace(0, 50, 53), np.linspace(-3, 40, 150)]
    grid = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    tight = 1e-3 * raw_tol(96)
    gsb.set_option("force_path", 2)
    try:
        ref_s = gsb.summate_structured(cov, z1, z2, axes)
        ref_v = gsb.summate_incompr_structured(cov, z1, z2, axes)
        assert maxabs(ref_s, oracle_mod.summate(cov, z1, z2, grid).reshape(37, 53, 150)) <= raw_tol(96)
        assert maxabs(ref_v, oracle_mod.summate_incompr(cov, z1, z2, grid).reshape(3, 37, 53, 150)) <= raw_tol(96)
        assert np.array_equal(gsb.summate_structured(cov, z1, z2, axes), ref_s)
        sets = [synth_modes(3, 96, seed=30 + b) for b in range(5)]
        bc_, b1, b2 = (np.stack([s_[i] for s_ in sets]) for i in range(3))
        ref_b = gsb.summate_structured(bc_, b1, b2, axes)
        ref_bv = gsb.summate_incompr_structured(bc_, b1, b2, axes)
        assert np.array_equal(gsb.summate_incompr_structured(bc_, b1, b2, axes), ref_bv)
        t = [torch.tensor(a, device=dev) for a in (bc_, b1, b2)]
        out = gsb.summate_incompr_structured(t[0], t[1], t[2], [torch.tensor(a, device=dev) for a in axes])
        assert maxabs(out.cpu().numpy(), ref_bv) <= tight
        for b in range(5):
            assert maxabs(ref_b[b], gsb.summate_structured(*sets[b], axes)) <= tight
            want = oracle_mod.summate_incompr(*sets[b], grid).reshape(3, 37, 53, 150)
            assert maxabs(ref_bv[b], want) <= raw_tol(96)
        # a forced odd grid (other shares) moves nothing beyond rounding
        gsb.set_option("sk_grid", 37)
        assert maxabs(gsb.summate_structured(cov, z1, z2, axes), ref_s) <= tight
    finally:
        gsb.set_option("sk_grid", 0)
        gsb.set_option("force_path", 0)


def test_concurrent_calls_from_python_threads(gsb, oracle_mod):
    """SURVEY 8(b) threading: the reference's native functions are re-entrant and release the GIL;
    so does the C ABI (ctypes drops the GIL, per-device work is serialised by a mutex).  Several
    Python threads calling different entry points at once get the single-threaded bits."""
    import threading

    cov, z1, z2 = synth_modes(3, 300, seed=9)
    pos = np.random.RandomState(4).uniform(0, 80, (3, 60001))
    axes = [np.arange(20.0), np.arange(140.0), np.arange(150.0)]
    rs = np.random.RandomState(8)
    kmat, kcond, kv = rs.normal(size=(90, 90)), rs.normal(size=90), rs.uniform(-1, 1, (90, 30000))
    jobs = {
        "flat": lambda: gsb.summate(cov, z1, z2, pos),
        "vec": lambda: gsb.summate_incompr(cov, z1, z2, pos[:, :20000]),
        "mesh": lambda: gsb.summate_structured(cov, z1, z2, axes),
        "krige": lambda: np.stack(gsb.calc_field_krige_and_variance(kmat, kv, kcond)),
    }
    want = {k: np.array(fn()) for k, fn in jobs.items()}
    got, errors = {}, []

    def work(name, rep):
        try:
            got[(name, rep)] = np.array(jobs[name]())
        except Exception as exc:  # pragma: no cover
            errors.append((name, exc))

    threads = [threading.Thread(target=work, args=(n, r)) for r in range(3) for n in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for (name, rep), val in got.items():
        assert np.array_equal(val, want[name]), name


def test_c_program_calls_the_library_on_the_gpu(gsb, tmp_path):
    """The drop-in boundary from plain C: one summation and one kriging evaluation against closed forms."""
    import subprocess

    from test_abi import _build_c_consumer

    exe = _build_c_consumer(tmp_path, gsb)
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "gpu ok" in out.stdout, out.stderr + out.stdout


def test_pinned_output_budget_on_the_gpu(gsb, monkeypatch):
    """Same accounting test as on CPU, here with real pinned allocations."""
    from test_host_logic import test_pinned_output_budget

    test_pinned_output_budget(gsb, monkeypatch)


def test_release_memory_returns_the_cached_scratch(gsb):
    import torch

    cov, z1, z2 = synth_modes(3, 100, seed=2)
    axes = [np.arange(64.0), np.arange(256.0), np.arange(256.0)]
    ref = gsb.summate_structured(cov, z1, z2, axes)                  # host route: device output + scratch cached
    torch.cuda.synchronize()
    free_cached, _ = torch.cuda.mem_get_info()
    gsb.release_memory()
    free_released, _ = torch.cuda.mem_get_info()
    assert free_released >= free_cached + 30 * 1024 * 1024           # at least the 33 MB field came back
    again = gsb.summate_structured(cov, z1, z2, axes)                # and everything still works
    assert np.array_equal(again, ref)


@pytest.mark.parametrize("scale", [1e3, 1e5, 1e7, 1e9])
def test_large_phases_degrade_like_the_reference(scale, gsb, oracle_mod):
    """|phase| up to ~1e10 rad: the exact range reduction keeps the kernel's own error at the
    polynomial level; what grows is the rounding of the phase itself, |phase| * eps, which the
    reference's libm path carries as well.  Bound: sum_j (|z1_j| + |z2_j|) * (4 eps |phase|_max + 1e-12)."""
    cov, z1, z2 = synth_modes(3, 200, seed=17)
    pos = np.random.RandomState(5).uniform(-scale, scale, (3, 20000))
    got = gsb.summate(cov, z1, z2, pos)
    want = oracle_mod.summate(cov, z1, z2, pos)
    phase_max = float(np.max(np.abs(cov.T @ pos[:, :2000])))
    bound = (np.abs(z1).sum() + np.abs(z2).sum()) * (4 * np.finfo(float).eps * phase_max * 3 + 1e-12)
    assert maxabs(got, want) <= bound, (maxabs(got, want), bound, phase_max)
    gv = gsb.summate_incompr(cov, z1, z2, pos)
    wv = oracle_mod.summate_incompr(cov, z1, z2, pos)
    assert maxabs(gv, wv) <= 2 * bound


@pytest.mark.parametrize("offset", [0.0, 1e4, 1e7])
def test_structured_mesh_far_from_the_origin(offset, gsb, oracle_mod):
    """Meshes with large coordinates (e.g. UTM eastings ~ 1e6..1e7): the per-axis phase tables use
    full-range sincos, so the separable path degrades only by the rounding of the phases, like the
    reference."""
    cov, z1, z2 = synth_modes(3, 150, seed=23)
    axes = [offset + np.arange(24.0), 2 * offset + np.linspace(0, 70, 130), -offset + np.arange(260.0)]
    grid = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    want = oracle_mod.summate(cov, z1, z2, grid).reshape(24, 130, 260)
    phase_max = float(np.max(np.abs(cov.T @ grid[:, ::997])))
    bound = (np.abs(z1).sum() + np.abs(z2).sum()) * (4 * np.finfo(float).eps * phase_max * 3 + 1e-12)
    for force in (1, 2):
        gsb.set_option("force_path", force)
        try:
            got = gsb.summate_structured(cov, z1, z2, axes)
        finally:
            gsb.set_option("force_path", 0)
        assert maxabs(got, want) <= bound, (force, maxabs(got, want), bound)


@pytest.mark.parametrize("shape", [(300, 260, 10), (129, 33, 5), (40, 50, 70), (6, 20, 9, 14)])
def test_thin_meshes_fold_the_last_two_axes(shape, gsb, oracle_mod):
    """Meshes whose last axis is short (layered models: nz = 5..50) would use a fraction of every
    128-wide output tile; the separable path then treats the last two axes as one.  Scalar and
    vector fields, with an isometrisation matrix, against the oracle and against the unfolded path."""
    dim = len(shape)
    cov, z1, z2 = synth_modes(dim, 130, seed=sum(shape))
    axes = [np.sort(np.random.RandomState(t).uniform(-5, 60, s)) for t, s in enumerate(shape)]
    mat = np.random.RandomState(9).normal(size=(dim, dim))
    grid = mat @ np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    want = oracle_mod.summate(cov, z1, z2, grid).reshape(shape)
    gsb.set_option("force_path", 2)
    try:
        out = {}
        for fold in (0, 2):
            gsb.set_option("fold_axes", fold)
            before = gsb.get_counter("folded_calls")
            out[fold] = gsb.summate_structured(cov, z1, z2, axes, mat)
            assert (gsb.get_counter("folded_calls") > before) == (fold == 2)
            assert maxabs(out[fold], want) <= raw_tol(130)
        if dim == 3:
            wantv = oracle_mod.summate_incompr(cov, z1, z2, grid).reshape((3,) + shape)
            gotv = gsb.summate_incompr_structured(cov, z1, z2, axes, mat)
            assert maxabs(gotv, wantv) <= raw_tol(130)
        # automatic choice (cost model: padded-tile contraction + table building, either way)
        gsb.set_option("fold_axes", 1)
        auto = gsb.summate_structured(cov, z1, z2, axes, mat)
        assert maxabs(auto, want) <= raw_tol(130)
        # fused epilogue and batches go through the folded path unchanged
        gsb.set_option("fold_axes", 2)
        covb = np.stack([cov, 0.5 * cov])
        rawb = gsb.summate_structured(covb, np.stack([z1, z2]), np.stack([z2, z1]), axes, mat)
        gotb = gsb.summate_structured(covb, np.stack([z1, z2]), np.stack([z2, z1]), axes, mat,
                                      epilogue=(0.25, [0.0, 1.5]))
        assert maxabs(rawb[0], want) <= raw_tol(130)
        assert np.array_equal(gotb, oracle_mod.apply_epilogue(rawb, 0.25, [0.0, 1.5]))
    finally:
        gsb.set_option("fold_axes", 1)
        gsb.set_option("force_path", 0)


def test_badly_filled_tiles_go_to_the_direct_kernel(gsb, oracle_mod):
    """A 2-D mesh with a very short last axis fills 5 % of every 128-wide tile and cannot be folded:
    the structured entry then expands it on the device and uses the direct kernel."""
    cov, z1, z2 = synth_modes(2, 64, seed=4)
    axes = [np.arange(30000.0) * 0.01, np.arange(6.0)]
    before = (gsb.get_counter("direct_calls"), gsb.get_counter("separable_calls"))
    got = gsb.summate_structured(cov, z1, z2, axes)
    assert gsb.get_counter("direct_calls") == before[0] + 1 and gsb.get_counter("separable_calls") == before[1]
    grid = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    assert maxabs(got.reshape(-1), oracle_mod.summate(cov, z1, z2, grid)) <= raw_tol(64)


def test_direct_tail_split_keeps_the_bits(gsb, oracle_mod):
    """A point set of a few full waves plus a fraction runs as two launches (full waves in the big-CTA configuration,
    the tail in smaller CTAs): same bits as a forced single configuration, one direct call counted."""
    import torch

    sms = torch.cuda.get_device_properties(0).multi_processor_count
    n = 2 * sms * 2 * 1024 + 70001           # two waves of the 8-points-per-thread configuration (2-D) and a bit
    cov, z1, z2 = synth_modes(2, 40, seed=2)
    pos = np.random.RandomState(0).uniform(0, 100, (2, n))
    before = gsb.get_counter("direct_calls"), gsb.get_counter("launches")
    got = gsb.summate(cov, z1, z2, torch.tensor(pos, device="cuda:0")).cpu().numpy()
    assert gsb.get_counter("direct_calls") == before[0] + 1
    assert gsb.get_counter("launches") == before[1] + 3          # mode records + two direct launches
    gsb.set_option("direct_cfg", 2)
    try:
        one = gsb.summate(cov, z1, z2, torch.tensor(pos, device="cuda:0")).cpu().numpy()
    finally:
        gsb.set_option("direct_cfg", -1)
    assert np.array_equal(got, one)
    idx = np.r_[0:1000, n - 1000:n]
    assert np.max(np.abs(got[idx] - oracle_mod.summate(cov, z1, z2, pos[:, idx]))) <= 1e-9 * np.sqrt(40)
    # per-point epilogue arrays follow the tail
    gain = torch.rand(n, device="cuda:0", dtype=torch.float64)
    pe = gsb.make_point_epilogue(gain, None, [0.5])
    fused = gsb.summate(cov, z1, z2, torch.tensor(pos, device="cuda:0"), epilogue=gsb.make_epilogue(2.0, []),
                        point_epilogue=pe).cpu().numpy()
    assert np.array_equal(fused, gain.cpu().numpy() * (2.0 * got) + 0.5)
