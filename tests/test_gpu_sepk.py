"""The second-generation separable contraction (gsb_sepk.cuh: no pre-generated A operand, stream-K split) on the
GPU against the CPU oracle: scalar and vector fields, partial row / column tiles, folded tile axes, tiles split
over several CTAs (forced small and odd grids), batches, the host route's pieces, fused epilogues."""
import numpy as np
import pytest

from conftest import synth_modes

pytestmark = pytest.mark.gpu


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


def raw_tol(n_modes, var=1.0):
    return 1e-9 * np.sqrt(var) / np.sqrt(var / n_modes)


@pytest.fixture()
def sk(gsb):
    gsb.set_option("force_path", 2)
    yield gsb
    gsb.set_option("force_path", 0)
    gsb.set_option("sk_grid", 0)
    gsb.set_option("sk_table_mb", 256)


SK_CASES = [
    (2, (100, 100), 1000),        # config 1 shape: no prefix axes, nothing rescaled
    (2, (3, 5), 10),
    (2, (129, 131), 77),          # odd last axis: scalar stores; partial rows and columns
    (2, (300, 517), 200),
    (3, (40, 50, 130), 300),      # tile axis = both row axes folded (2000 rows)
    (3, (7, 9, 257), 64),
    (3, (64, 64, 256), 1000),
    (3, (6, 200, 136), 50),       # tile axis 200: partial row tile of 72 rows
    (4, (9, 10, 11, 140), 64),
    (5, (3, 4, 5, 6, 130), 20),
]


@pytest.mark.parametrize("dim,lens,n_modes", SK_CASES)
@pytest.mark.parametrize("grid,table_mb", [(0, 256), (5, 256), (148, 1)])
def test_sk_structured_vs_oracle(dim, lens, n_modes, grid, table_mb, sk, oracle_mod):
    """grid = 5 splits every tile chain over few CTAs, 148 (with tiny meshes) splits single tiles over many;
    table_mb = 1 keeps the tile-axis table small, i.e. forces prefix axes (the rescaling variant)."""
    cov, z1, z2 = synth_modes(dim, n_modes, seed=7 + dim)
    rs = np.random.RandomState(5)
    axes = [np.sort(rs.uniform(0, 200, L)) for L in lens]
    mat = rs.normal(size=(dim, dim))
    pos = mat @ np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    want = oracle_mod.summate(cov, z1, z2, pos).reshape(lens)
    sk.set_option("sk_grid", grid)
    sk.set_option("sk_table_mb", table_mb)
    before = sk.get_counter("sk_calls")
    got = sk.summate_structured(cov, z1, z2, axes, mat)
    assert sk.get_counter("sk_calls") == before + 1
    assert got.shape == tuple(lens)
    assert maxabs(got, want) <= raw_tol(n_modes)
    again = sk.summate_structured(cov, z1, z2, axes, mat)
    assert np.array_equal(got, again), "the split is deterministic"
    if dim in (2, 3):
        wv = oracle_mod.summate_incompr(cov, z1, z2, pos).reshape((dim,) + tuple(lens))
        gv = sk.summate_incompr_structured(cov, z1, z2, axes, mat)
        assert gv.shape == (dim,) + tuple(lens)
        assert maxabs(gv, wv) <= raw_tol(n_modes)


def test_sk_single_128cubed_field_and_batch(sk, oracle_mod):
    """One 128^3 field is 128 tiles on 148 SMs: every CTA owns ~0.86 tile, almost every tile is finished from two
    partial accumulations.  Device route, host route and a batch agree with the oracle on a node sample."""
    import torch

    cov, z1, z2 = synth_modes(3, 1000, seed=11)
    axes = [np.arange(128.0)] * 3
    rs = np.random.RandomState(1)
    idx = np.unique(np.concatenate([rs.randint(0, 128 ** 3, 6000), [0, 128 ** 3 - 1]]))
    sub = np.unravel_index(idx, (128,) * 3)
    pos = np.stack([axes[t][sub[t]] for t in range(3)])
    want = oracle_mod.summate(cov, z1, z2, pos)
    host = sk.summate_structured(cov, z1, z2, axes)
    assert maxabs(host[sub], want) <= raw_tol(1000)
    dev = sk.summate_structured(*(torch.tensor(a, device="cuda:0") for a in (cov, z1, z2)),
                                [torch.tensor(a, device="cuda:0") for a in axes])
    assert np.array_equal(dev.cpu().numpy(), host)
    # batch of 5 mode sets: the shares cross field boundaries
    covb = np.stack([cov * s for s in (1.0, 0.5, -1.0, 2.0, 0.25)])
    z1b, z2b = np.stack([z1] * 5), np.stack([z2] * 5)
    got = sk.summate_structured(covb, z1b, z2b, axes)
    # (not the bits of the single-field call: the stream-K shares cut the mode sums at other stages)
    assert maxabs(got[0], host) <= 1e-3 * raw_tol(1000)
    for b in (1, 4):
        assert maxabs(got[b][sub], oracle_mod.summate(covb[b], z1, z2, pos)) <= raw_tol(1000)
    # fused epilogue: the bits of the numpy passes applied to the raw sums of the same kernel
    fused = sk.summate_structured(cov, z1, z2, axes, epilogue=(0.03, [0.0, 1.5]))
    assert np.array_equal(fused, oracle_mod.apply_epilogue(host, 0.03, [0.0, 1.5]))


@pytest.mark.parametrize("shape", [(200, 200, 200), (100, 100, 100), (300, 260, 10), (512, 16, 300), (1, 150, 140),
                                   (70, 1, 1, 200), (260, 3, 5, 129)])
def test_sk_odd_and_thin_meshes(shape, sk, oracle_mod):
    dim = len(shape)
    cov, z1, z2 = synth_modes(dim, 130, seed=sum(shape))
    axes = [np.sort(np.random.RandomState(t).uniform(-5, 60, s)) for t, s in enumerate(shape)]
    mat = np.random.RandomState(9).normal(size=(dim, dim))
    n = int(np.prod(shape))
    idx = np.unique(np.concatenate([np.random.RandomState(2).randint(0, n, 20000), [0, n - 1]]))
    sub = np.unravel_index(idx, shape)
    pos = mat @ np.stack([axes[t][sub[t]] for t in range(dim)])
    got = sk.summate_structured(cov, z1, z2, axes, mat)
    assert maxabs(got[sub], oracle_mod.summate(cov, z1, z2, pos)) <= raw_tol(130)
    if dim == 3:
        gv = sk.summate_incompr_structured(cov, z1, z2, axes, mat)
        assert maxabs(gv[(slice(None),) + sub], oracle_mod.summate_incompr(cov, z1, z2, pos)) <= raw_tol(130)
    # host route (pieces + overlapped D2H) against the device route: the same tiles; the host route may pick another
    # tile axis (it wants pieces that leave as contiguous copies) and other stream-K shares -> equal up to rounding
    import torch

    dev = sk.summate_structured(*(torch.tensor(a, device="cuda:0") for a in (cov, z1, z2)),
                                [torch.tensor(a, device="cuda:0") for a in axes], mat)
    assert maxabs(dev.cpu().numpy(), got) <= 1e-3 * raw_tol(130)


PACK_CASES = [
    (3, (5, 200, 136), 96),        # 200 rows per slow index: tiles span two slow indices
    (3, (9, 72, 130), 64),         # 72 rows: up to three slow indices per tile
    (3, (7, 50, 300), 40),         # 50 rows (padded to 56): four segments per tile
    (3, (4, 300, 140), 72),
    (3, (23, 100, 100), 130),      # the 100^3 family
    (4, (3, 5, 100, 129), 48),     # two outer slow axes
    (3, (33, 129, 64), 50),        # one row more than a tile
]


@pytest.mark.parametrize("dim,lens,n_modes", PACK_CASES)
@pytest.mark.parametrize("grid", [0, 7])
def test_sk_row_packing_vs_oracle(dim, lens, n_modes, grid, sk, oracle_mod):
    """Row tiles packed across the slow indices (sk_pack = 2: whenever the layout allows it): scalar and vector
    fields, device and host route (several pieces), batches, fused epilogue -- against the CPU oracle."""
    import torch

    cov, z1, z2 = synth_modes(dim, n_modes, seed=3 + dim)
    rs = np.random.RandomState(11)
    axes = [np.sort(rs.uniform(0, 150, L)) for L in lens]
    mat = rs.normal(size=(dim, dim))
    pos = mat @ np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    want = oracle_mod.summate(cov, z1, z2, pos).reshape(lens)
    sk.set_option("sk_grid", grid)
    sk.set_option("sk_pack", 2)
    try:
        before = sk.get_counter("packed_calls")
        dev = sk.summate_structured(*(torch.tensor(a, device="cuda:0") for a in (cov, z1, z2)),
                                    [torch.tensor(a, device="cuda:0") for a in axes], mat)
        assert sk.get_counter("packed_calls") == before + 1, "the packed kernel did not run"
        assert maxabs(dev.cpu().numpy(), want) <= raw_tol(n_modes)
        for pieces in (0, 5):
            sk.set_option("host_pieces", pieces)
            host = sk.summate_structured(cov, z1, z2, axes, mat)
            assert maxabs(host, want) <= raw_tol(n_modes)
        sk.set_option("host_pieces", 3)
        if dim == 3:
            wv = oracle_mod.summate_incompr(cov, z1, z2, pos).reshape((3,) + tuple(lens))
            assert maxabs(sk.summate_incompr_structured(cov, z1, z2, axes, mat), wv) <= raw_tol(n_modes)
        covb = np.stack([cov, 0.5 * cov, -cov])
        z1b, z2b = np.stack([z1, z1, z2]), np.stack([z2, z1, z1])
        got = sk.summate_structured(covb, z1b, z2b, axes, mat)
        assert maxabs(got[0], want) <= raw_tol(n_modes)
        assert maxabs(got[2], oracle_mod.summate(-cov, z2, z1, pos).reshape(lens)) <= raw_tol(n_modes)
        fused = sk.summate_structured(cov, z1, z2, axes, mat, epilogue=(0.25, [1.0]))
        assert maxabs(fused, 0.25 * want + 1.0) <= raw_tol(n_modes)
        # packed against unpacked: the same sums in another tiling
        sk.set_option("sk_pack", 1)
        plain = sk.summate_structured(cov, z1, z2, axes, mat)
        assert maxabs(plain, host) <= 1e-3 * raw_tol(n_modes)
    finally:
        sk.set_option("sk_pack", 0)
        sk.set_option("host_pieces", 0)


def test_sk_row_packing_with_inner_slow_axis(sk, oracle_mod):
    """Tile axis = axis 0 (100 rows), the 7 entries of axis 1 are an INNER slow axis: packed tiles span slow indices whose
    rows interleave in the output (row = iy * n_in + inner).  A small table cap rules the folded layout out."""
    import torch

    lens, n_modes = (100, 7, 140), 64
    cov, z1, z2 = synth_modes(3, n_modes, seed=12)
    rs = np.random.RandomState(3)
    axes = [np.sort(rs.uniform(0, 90, L)) for L in lens]
    mat = rs.normal(size=(3, 3))
    pos = mat @ np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    want = oracle_mod.summate(cov, z1, z2, pos).reshape(lens)
    sk.set_option("sk_pack", 2)
    sk.set_option("sk_table_mb", 1)
    try:
        before = sk.get_counter("packed_calls")
        dev = sk.summate_structured(*(torch.tensor(a, device="cuda:0") for a in (cov, z1, z2)),
                                    [torch.tensor(a, device="cuda:0") for a in axes], mat)
        assert sk.get_counter("packed_calls") == before + 1
        assert maxabs(dev.cpu().numpy(), want) <= raw_tol(n_modes)
        sk.set_option("host_pieces", 3)
        host = sk.summate_incompr_structured(cov, z1, z2, axes, mat)
        wv = oracle_mod.summate_incompr(cov, z1, z2, pos).reshape((3,) + lens)
        assert maxabs(host, wv) <= raw_tol(n_modes)
    finally:
        sk.set_option("sk_pack", 0)
        sk.set_option("host_pieces", 0)
