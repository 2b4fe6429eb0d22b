"""Host-side logic that needs no GPU: argument handling, the gstools plugin wiring, sharding."""
import numpy as np
import pytest

import refharness
from conftest import load_golden


# ---------------------------------------------------------------------------------------------
# backend argument handling (mirrors the reference's buffer coercion errors)
# ---------------------------------------------------------------------------------------------
def test_shape_validation(gsb):
    cov, z = np.zeros((2, 4)), np.zeros(4)
    with pytest.raises(ValueError):
        gsb.summate(cov, z[:3], z, np.zeros((2, 3)))
    with pytest.raises(ValueError):
        gsb.summate(cov, z, z, np.zeros((3, 3)))       # dim mismatch
    with pytest.raises(ValueError):
        gsb.summate(cov, z, z, np.zeros(3))            # pos not 2-d
    with pytest.raises(ValueError):
        gsb.summate(np.zeros(4), z, z, np.zeros((2, 3)))
    with pytest.raises(TypeError):
        gsb.summate(cov, z, z, np.array([["a", "b"], ["c", "d"]]))
    with pytest.raises(ValueError):
        gsb.summate_structured(cov, z, z, [np.arange(3.0)])            # 1 axis for dim 2
    with pytest.raises(ValueError):
        gsb.summate_structured(cov, z, z, [np.arange(3.0)] * 2, matrix=np.eye(3))


def test_rows_contiguous_avoids_copies(gsb):
    from gstools_b200.backend import _rows_contiguous

    big = np.arange(40.0).reshape(2, 20)
    a, ld = _rows_contiguous(big)
    assert a is big and ld == 20
    view = big[:, 5:12]
    a, ld = _rows_contiguous(view)
    assert a is view and ld == 20                       # row-strided: no copy
    a, ld = _rows_contiguous(big[:, ::2])
    assert a.flags.c_contiguous and ld == 10            # inner stride != 1: copied
    a, ld = _rows_contiguous(np.asfortranarray(big))
    assert a.flags.c_contiguous and ld == 20


def test_device_selection(gsb, monkeypatch):
    from gstools_b200 import backend

    monkeypatch.setattr(backend, "_DEVICE", None)
    monkeypatch.delenv("GSB200_DEVICE", raising=False)
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    assert backend.get_device() == 0
    monkeypatch.setenv("LOCAL_RANK", "3")
    assert backend.get_device() == 3
    monkeypatch.setenv("GSB200_DEVICE", "5")
    assert backend.get_device() == 5
    backend.set_device(1)
    assert backend.get_device() == 1
    monkeypatch.setattr(backend, "_DEVICE", None)


# ---------------------------------------------------------------------------------------------
# sharding arithmetic
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (10, 3), (512, 8), (134217728, 8), (7, 7)])
def test_shard_range_partitions(n, world):
    from gstools_b200.dist import shard_range

    ranges = [shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (a, b), (c, d) in zip(ranges, ranges[1:]):
        assert b == c
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1


# ---------------------------------------------------------------------------------------------
# plugin wiring against the UNMODIFIED reference (CPU: the B200 flag is switched off so the
# calls fall through to the reference's own backend, i.e. the oracle stub)
# ---------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not refharness.have_reference(), reason="reference gstools not present")


@needs_ref
def test_enable_rebinds_and_disable_restores(gsb):
    gs = refharness.import_gstools()
    from gstools import config
    from gstools.field import base as fbase
    from gstools.field import generator as gen

    o1, o2, o3 = gen._summate, gen._summate_incompr, fbase.Field.pre_pos
    o4 = gen._summate_fourier
    gsb.enable()
    try:
        assert gsb.is_enabled() and config.USE_GSTOOLS_B200 is True
        assert gen._summate is not o1 and gen._summate_incompr is not o2
        assert fbase.Field.pre_pos is not o3
        gsb.enable()  # idempotent: originals are remembered once
    finally:
        gsb.disable()
    assert gen._summate is o1 and gen._summate_incompr is o2 and fbase.Field.pre_pos is o3
    assert gen._summate_fourier is o4
    assert config.USE_GSTOOLS_B200 is False and not gsb.is_enabled()


@needs_ref
def test_flag_off_falls_through_to_reference_backend(gsb):
    """With USE_GSTOOLS_B200 False the rebound wrappers must behave exactly like the originals,
    including for the lazy structured placeholder (materialised via generate_grid + matrix)."""
    gs = refharness.import_gstools()
    from gstools import config

    meta, d = load_golden("srf_exp3d_rot_anis_struct")
    model = gs.Exponential(dim=3, var=2.0, len_scale=[12.0, 5.0, 3.0], angles=[0.4, -0.3, 0.7])
    axes = [d["axis0"], d["axis1"], d["axis2"]]
    gsb.enable()
    try:
        config.USE_GSTOOLS_B200 = False
        srf = gs.SRF(model, seed=20170519, mode_no=256)
        field = srf.structured(axes)
        assert np.allclose(field, d["field"], rtol=0, atol=1e-12)
    finally:
        gsb.disable()


@needs_ref
def test_lazy_grid_placeholder_round_trip(gsb):
    gs = refharness.import_gstools()
    from gstools.field.generator import generate_grid
    from gstools_b200.plugin import LazyGridPos, _lookup_lazy

    axes = (np.linspace(0, 1, 4), np.linspace(2, 3, 5), np.arange(6.0))
    mat = np.random.RandomState(0).normal(size=(3, 3))
    lazy = LazyGridPos(axes, mat)
    assert lazy.shape == (3, 4 * 5 * 6) and lazy.strides == (0, 0)
    # RandMeth.__call__ coerces with np.asarray(pos, dtype=double) (generator.py:261)
    coerced = np.asarray(lazy, dtype=np.double)
    assert type(coerced) is np.ndarray
    got = _lookup_lazy(coerced)
    assert got is not None and got[1] is lazy.matrix
    assert all(np.array_equal(a, b) for a, b in zip(got[0], axes))
    # an ordinary array never matches
    assert _lookup_lazy(generate_grid(axes)) is None
    assert _lookup_lazy(np.zeros((3, 5))) is None
    key = lazy._holder.__array_interface__["data"][0]
    del lazy, coerced, got
    import gc

    gc.collect()
    from gstools_b200 import plugin

    assert key not in plugin._LAZY


@needs_ref
def test_lazy_pre_pos_only_for_randmeth_structured(gsb):
    gs = refharness.import_gstools()
    from gstools_b200.plugin import LazyGridPos

    gsb.enable()
    try:
        model = gs.Gaussian(dim=2, var=1, len_scale=3)
        srf = gs.SRF(model, seed=1, mode_no=16)
        iso, shape = srf.pre_pos([np.arange(4.0), np.arange(5.0)], "structured")
        assert isinstance(iso, LazyGridPos) and shape == (4, 5)
        iso, shape = srf.pre_pos([np.arange(4.0), np.arange(4.0)], "unstructured")
        assert not isinstance(iso, LazyGridPos) and iso.shape == (2, 4)
        # Krige objects (no generator) keep the real positions
        krige = gs.krige.Ordinary(model, [np.array([0.0, 1.0]), np.array([0.0, 1.0])],
                                  np.array([1.0, 2.0]))
        iso, shape = krige.pre_pos([np.arange(3.0), np.arange(3.0)], "structured")
        assert not isinstance(iso, LazyGridPos) and iso.shape == (2, 9)
        # lat-lon models are not separable (geometric.py:659-664): stay on the flat path
        ll = gs.Gaussian(latlon=True, var=1, len_scale=500, geo_scale=gs.KM_SCALE)
        srf_ll = gs.SRF(ll, seed=1, mode_no=16)
        iso, shape = srf_ll.pre_pos([np.arange(3.0), np.arange(4.0)], "structured")
        assert not isinstance(iso, LazyGridPos)
    finally:
        gsb.disable()
