"""Row f2 on the GPU: the caller epilogue fused into the kernels' stores.

Bar: BIT-EXACT.  ``fn(..., epilogue=E)`` must equal the reference's numpy passes
(oracle.apply_epilogue: generator.py:269-270, 561-567; normalizer/tools.py:99-103) applied to
the raw sums the same kernel returns without an epilogue."""
import numpy as np
import pytest

import refharness
from conftest import load_golden, synth_modes

pytestmark = pytest.mark.gpu

EPIS = [
    (0.0316227766016838, []),
    (0.0316227766016838, [0.0, 1.25, -0.3]),
    (-1.7e-3, [(0.5, -0.0, 0.25), 0.0, 2.0, (1.0, -1.0, 3.0)]),
]


def _adds(adds, ncomp):
    return [a[:ncomp] if isinstance(a, tuple) and ncomp > 1 else (a[0] if isinstance(a, tuple) else a)
            for a in adds]


@pytest.mark.parametrize("scale,adds", EPIS)
@pytest.mark.parametrize("dim,n,n_modes", [(2, 1000, 100), (3, 303105, 16), (1, 77, 5), (3, 200, 5000)])
def test_direct_kernel_epilogue_bits(scale, adds, dim, n, n_modes, gsb, oracle_mod):
    """All three direct configurations plus the mode-split / reduce path (200 points, 5000 modes)."""
    cov, z1, z2 = synth_modes(dim, n_modes, seed=n)
    pos = np.random.RandomState(n).uniform(-100, 500, (dim, n))
    raw = gsb.summate(cov, z1, z2, pos)
    got = gsb.summate(cov, z1, z2, pos, epilogue=gsb.make_epilogue(scale, _adds(adds, 1)))
    assert np.array_equal(got, oracle_mod.apply_epilogue(raw, scale, _adds(adds, 1)))
    if dim in (2, 3):
        raw = gsb.summate_incompr(cov, z1, z2, pos)
        got = gsb.summate_incompr(cov, z1, z2, pos, epilogue=(scale, _adds(adds, dim)))
        assert np.array_equal(got, oracle_mod.apply_epilogue(raw, scale, _adds(adds, dim)))


@pytest.mark.parametrize("scale,adds", EPIS[1:])
@pytest.mark.parametrize("shape,force", [((24, 130, 260), 2), ((300, 520), 2), ((70, 3, 200), 2), ((9, 16), 1)])
def test_structured_kernel_epilogue_bits(scale, adds, shape, force, gsb, oracle_mod):
    """The contraction with and without slow axes (3-D / 2-D), partial tiles, and a mesh expanded on the device
    and sent through the direct kernel."""
    dim = len(shape)
    cov, z1, z2 = synth_modes(dim, 130, seed=sum(shape))
    axes = [np.linspace(-3.0, 40.0, s) for s in shape]
    mat = np.random.RandomState(1).normal(size=(dim, dim))
    gsb.set_option("force_path", force)
    try:
        raw = gsb.summate_structured(cov, z1, z2, axes, mat)
        got = gsb.summate_structured(cov, z1, z2, axes, mat, epilogue=(scale, _adds(adds, 1)))
        assert np.array_equal(got, oracle_mod.apply_epilogue(raw, scale, _adds(adds, 1)))
        rawv = gsb.summate_incompr_structured(cov, z1, z2, axes, mat)
        gotv = gsb.summate_incompr_structured(cov, z1, z2, axes, mat, epilogue=(scale, _adds(adds, dim)))
        assert np.array_equal(gotv, oracle_mod.apply_epilogue(rawv, scale, _adds(adds, dim)))
        # batched mode sets share one epilogue
        covb = np.stack([cov, 0.5 * cov])
        rawb = gsb.summate_structured(covb, np.stack([z1, z2]), np.stack([z2, z1]), axes, mat)
        gotb = gsb.summate_structured(covb, np.stack([z1, z2]), np.stack([z2, z1]), axes, mat,
                                      epilogue=(scale, _adds(adds, 1)))
        assert np.array_equal(gotb, oracle_mod.apply_epilogue(rawb, scale, _adds(adds, 1)))
    finally:
        gsb.set_option("force_path", 0)


def test_device_tensor_epilogue(gsb, oracle_mod):
    import torch

    cov, z1, z2 = synth_modes(3, 64, seed=5)
    axes = [np.arange(20.0), np.arange(140.0), np.arange(150.0)]
    t = lambda a: torch.as_tensor(a, device="cuda")
    raw = gsb.summate_structured(t(cov), t(z1), t(z2), [t(a) for a in axes])
    got = gsb.summate_structured(t(cov), t(z1), t(z2), [t(a) for a in axes], epilogue=(0.125, [0.0, 3.0]))
    assert got.is_cuda
    assert np.array_equal(got.cpu().numpy(), oracle_mod.apply_epilogue(raw.cpu().numpy(), 0.125, [0.0, 3.0]))


def test_epilogue_argument_errors(gsb):
    cov, z1, z2 = synth_modes(2, 8, seed=1)
    with pytest.raises(ValueError):
        gsb.make_epilogue(1.0, [0.0] * 5)
    with pytest.raises(ValueError):
        gsb.make_epilogue(1.0, [(1.0, 2.0, 3.0, 4.0)])
    e = gsb.make_epilogue(1.0, [0.0])
    e.n_add = 9
    with pytest.raises(ValueError):
        gsb.summate(cov, z1, z2, np.zeros((2, 3)), epilogue=e)


# ---------------------------------------------------------------------------------------------
# through the unmodified reference: gs.SRF(...)(...) with the fused call
# ---------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not refharness.have_reference(), reason="reference gstools not present")


@needs_ref
@pytest.mark.parametrize("mesh", ["structured", "unstructured"])
@pytest.mark.parametrize("skw", [dict(mean=1.25, trend=-0.3),
                                 dict(generator="VectorField", mean=(0.5, 0.0, 1.0), mean_velocity=-1.3)])
def test_fused_srf_equals_unfused_backend_bits(gsb, mesh, skw):
    """fused=True and fused=False differ only in WHERE the affine map runs: same bits."""
    gs = refharness.import_gstools()
    model = gs.Exponential(dim=3, var=2.0, len_scale=[12.0, 5.0, 3.0], angles=[0.4, -0.3, 0.7])
    if mesh == "structured":
        pos = [np.arange(20.0), np.linspace(0, 70, 140), np.arange(150.0)]
    else:
        pos = np.random.RandomState(3).uniform(0, 100, (3, 5000))
    out = {}
    for fused in (False, True):
        gsb.enable(fused=fused)
        try:
            before = gsb.get_counter("launches")
            out[fused] = gs.SRF(model, seed=20170519, mode_no=200, **skw)(pos, mesh_type=mesh)
            assert gsb.get_counter("launches") > before
        finally:
            gsb.disable()
    assert out[True].shape == out[False].shape
    assert np.array_equal(out[True], out[False])


@needs_ref
def test_config1_through_fused_srf(gsb):
    """BASELINE.json configs[0] through the fused call, against the golden field of the reference."""
    gs = refharness.import_gstools()
    meta, d = load_golden("config1_gaussian2d_100x100")
    gsb.enable()
    try:
        srf = gs.SRF(gs.Gaussian(dim=2, var=1, len_scale=10), seed=20170519)
        field = srf.structured([np.arange(100.0), np.arange(100.0)])
        shifted = gs.SRF(gs.Gaussian(dim=2, var=1, len_scale=10), seed=20170519, mean=3.0)
        f3 = shifted.structured([np.arange(100.0), np.arange(100.0)])
    finally:
        gsb.disable()
    assert np.max(np.abs(field - d["field"])) <= 1e-9
    assert abs(field[0, 0] - -0.221389860323504) < 1e-9 and abs(field[50, 50] - 1.135479083967714) < 1e-9
    assert np.array_equal(f3, field + 0.0 + 3.0)
