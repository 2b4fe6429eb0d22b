"""The five BASELINE.json configs at FULL size on the GPU, checked through the oracle on samples
and through size-independent properties (sign symmetry, slab/batch invariance, agreement of the
two kernels)."""
import numpy as np
import pytest

import bench_configs as bc

pytestmark = pytest.mark.gpu
TOL = 1e-9


def sample_check(field_flat, cfg, oracle_mod, n_sample, seed, vec=False):
    """max |GPU - oracle| * sqrt(var/N) on a random sample of mesh nodes (+ the 8 corners)."""
    n = int(np.prod([len(a) for a in cfg["axes"]]))
    rs = np.random.RandomState(seed)
    idx = np.unique(np.concatenate([rs.randint(0, n, n_sample), [0, n - 1]]))
    pos = bc.grid_points(cfg["axes"], cfg.get("matrix"), idx)
    n_modes = cfg["cov"].shape[-1]
    scale = np.sqrt(cfg["var"] / n_modes)
    if vec:
        want = oracle_mod.summate_incompr(cfg["cov"], cfg["z1"], cfg["z2"], pos)
        return np.max(np.abs(field_flat[:, idx] - want)) * scale
    want = oracle_mod.summate(cfg["cov"], cfg["z1"], cfg["z2"], pos)
    return np.max(np.abs(field_flat[idx] - want)) * scale


def test_config1_full(gsb, oracle_mod):
    cfg = bc.config1()
    scale = np.sqrt(1.0 / 1000)
    for force in (1, 2):
        gsb.set_option("force_path", force)
        try:
            got = gsb.summate_structured(cfg["cov"], cfg["z1"], cfg["z2"], cfg["axes"])
        finally:
            gsb.set_option("force_path", 0)
        assert np.max(np.abs(got.reshape(-1) - cfg["raw"])) * scale <= TOL
        assert np.max(np.abs(scale * got - cfg["field"])) <= TOL
    flat = gsb.summate(cfg["cov"], cfg["z1"], cfg["z2"], cfg["pos"])
    assert np.max(np.abs(flat - cfg["raw"])) * scale <= TOL


def test_config2_full_512cubed(gsb, oracle_mod):
    import torch

    cfg = bc.config2(512)
    dev = torch.device("cuda:0")
    tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
    axes = [torch.tensor(a, device=dev) for a in cfg["axes"]]
    out = gsb.summate_structured(tc, t1, t2, axes)
    assert tuple(out.shape) == (512, 512, 512)
    flat = out.reshape(-1)
    # oracle on 20000 random nodes + corners
    n = 512 ** 3
    rs = np.random.RandomState(1)
    idx = np.unique(np.concatenate([rs.randint(0, n, 20000), [0, n - 1, 511, 512 * 511]]))
    got = flat[torch.tensor(idx, device=dev)].cpu().numpy()
    want = oracle_mod.summate(cfg["cov"], cfg["z1"], cfg["z2"], bc.grid_points(cfg["axes"], None, idx))
    assert np.max(np.abs(got - want)) * np.sqrt(1.0 / 1000) <= TOL
    # exact sign symmetry at full size: u(-z1, -z2) == -u(z1, z2) bit for bit
    neg = gsb.summate_structured(tc, -t1, -t2, axes)
    assert torch.equal(neg, -out)
    del neg
    # direct kernel on 4 whole x-planes agrees with the separable kernel
    for ix in (0, 137, 300, 511):
        yy, zz = np.meshgrid(cfg["axes"][1], cfg["axes"][2], indexing="ij")
        pos = np.stack([np.full(yy.size, cfg["axes"][0][ix]), yy.reshape(-1), zz.reshape(-1)])
        plane = gsb.summate(tc, t1, t2, torch.tensor(pos, device=dev)).reshape(512, 512)
        assert float((plane - out[ix]).abs().max()) * np.sqrt(1.0 / 1000) <= TOL
    # host-buffer route (8 pieces with overlapped D2H: other stream-K shares than the single device launch)
    host = gsb.summate_structured(cfg["cov"], cfg["z1"], cfg["z2"], cfg["axes"])
    assert np.max(np.abs(host[::37] - out[::37].cpu().numpy())) * np.sqrt(1.0 / 1000) <= 1e-3 * TOL
    assert np.array_equal(host[::37], gsb.summate_structured(cfg["cov"], cfg["z1"], cfg["z2"], cfg["axes"])[::37])


def test_config3_full_20m_points(gsb, oracle_mod):
    import torch

    cfg = bc.config3(20_000_000)
    dev = torch.device("cuda:0")
    tpos = torch.tensor(cfg["pos"], device=dev)
    out = gsb.summate(cfg["cov"], cfg["z1"], cfg["z2"], tpos)
    idx = np.random.RandomState(2).randint(0, 20_000_000, 8000)
    want = oracle_mod.summate(cfg["cov"], cfg["z1"], cfg["z2"], cfg["pos"][:, idx])
    got = out[torch.tensor(idx, device=dev)].cpu().numpy()
    assert np.max(np.abs(got - want)) * np.sqrt(1.0 / 10000) <= TOL
    # value of a point is independent of the batch it is evaluated in (sharding unit)
    sub = gsb.summate(cfg["cov"], cfg["z1"], cfg["z2"], tpos[:, 5_000_000:5_100_000])
    assert torch.equal(sub, out[5_000_000:5_100_000])
    # host-buffer route over pipelined chunks: same bits on a 6M-point prefix
    host = gsb.summate(cfg["cov"][:, :500], cfg["z1"][:500], cfg["z2"][:500], cfg["pos"][:, :6_000_000])
    devp = gsb.summate(cfg["cov"][:, :500], cfg["z1"][:500], cfg["z2"][:500], tpos[:, :6_000_000])
    assert np.array_equal(host, devp.cpu().numpy())


def test_config4_full_incompressible_256cubed(gsb, oracle_mod):
    import torch

    cfg = bc.config4(256)
    dev = torch.device("cuda:0")
    tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
    axes = [torch.tensor(a, device=dev) for a in cfg["axes"]]
    out = gsb.summate_incompr_structured(tc, t1, t2, axes)
    assert tuple(out.shape) == (3, 256, 256, 256)
    flat = out.reshape(3, -1).cpu().numpy()
    assert sample_check(flat, cfg, oracle_mod, 8000, seed=3, vec=True) <= TOL
    # incompressibility: the spectral projector makes every mode divergence free, so the
    # analytic divergence sum_t k_t p_t(k) vanishes identically
    k = cfg["cov"]
    proj = np.eye(3)[:, 0][:, None] - k * k[0] / np.sum(k * k, axis=0)
    assert np.max(np.abs(np.sum(k * proj, axis=0))) < 1e-15
    # flat direct kernel on a sub-block agrees
    sub_axes = [cfg["axes"][0][:8], cfg["axes"][1][100:140], cfg["axes"][2]]
    pos = bc.grid_points(sub_axes)
    direct = gsb.summate_incompr(cfg["cov"], cfg["z1"], cfg["z2"], pos).reshape(3, 8, 40, 256)
    assert np.max(np.abs(direct - out[:, :8, 100:140].cpu().numpy())) * np.sqrt(1.0 / 1000) <= TOL


def test_config5_full_ensemble_256x128cubed(gsb, oracle_mod):
    import torch

    cfg = bc.config5(128, 256)
    dev = torch.device("cuda:0")
    tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
    axes = [torch.tensor(a, device=dev) for a in cfg["axes"]]
    out = gsb.summate_structured(tc, t1, t2, axes)          # (256, 128, 128, 128): 4.3 GB
    assert tuple(out.shape) == (256, 128, 128, 128)
    n = 128 ** 3
    rs = np.random.RandomState(4)
    idx = rs.randint(0, n, 1500)
    pos = bc.grid_points(cfg["axes"], None, idx)
    tidx = torch.tensor(idx, device=dev)
    for b in (0, 1, 7, 8, 100, 255):   # the reference's own draws (0..7) and synthetic ones
        want = oracle_mod.summate(cfg["cov"][b], cfg["z1"][b], cfg["z2"][b], pos)
        got = out[b].reshape(-1)[tidx].cpu().numpy()
        assert np.max(np.abs(got - want)) * np.sqrt(1.0 / 1000) <= TOL
    # one realisation computed alone == the same realisation inside the batch up to rounding (seed sharding)
    single = gsb.summate_structured(tc[37], t1[37], t2[37], axes)
    assert float((single - out[37]).abs().max()) * np.sqrt(1.0 / 1000) <= 1e-3 * TOL


def test_beyond_the_configs_1024cubed_index_arithmetic(gsb, oracle_mod):
    """Maximum-size case: 1024^3 nodes (2^30 points, 8.6 GB of output, element offsets beyond 2^31 for
    the vector field).  Device resident; sampled against the oracle, plus the x-plane symmetry
    u(axis0 reversed) == u reversed, which touches every row chunk."""
    import torch

    cfg = bc.config2(512)
    dev = torch.device("cuda:0")
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~25 GB of device memory")
    tc, t1, t2 = (torch.tensor(cfg[k][..., :64], device=dev) for k in ("cov", "z1", "z2"))
    ax = [np.arange(1024.0), np.arange(1024.0) * 0.5, np.arange(1024.0) * 0.25]
    axes = [torch.tensor(a, device=dev) for a in ax]
    out = gsb.summate_structured(tc, t1, t2, axes)
    assert tuple(out.shape) == (1024, 1024, 1024)
    n = 1024 ** 3
    rs = np.random.RandomState(5)
    idx = np.unique(np.concatenate([rs.randint(0, n, 20000), [0, n - 1, n // 2, 2 ** 29 + 12345]]))
    got = out.reshape(-1)[torch.tensor(idx, device=dev)].cpu().numpy()
    want = oracle_mod.summate(cfg["cov"][:, :64], cfg["z1"][:64], cfg["z2"][:64], bc.grid_points(ax, None, idx))
    assert np.max(np.abs(got - want)) * np.sqrt(1.0 / 64) <= TOL
    rev = gsb.summate_structured(tc, t1, t2, [axes[0].flip(0), axes[1], axes[2]])
    assert float((rev - out.flip(0)).abs().max()) * np.sqrt(1.0 / 64) <= 1e-3 * TOL
    del rev, out
    # vector field on 1024 x 1024 x 768: 3 x 0.8e9 elements, component offsets beyond 2^31
    axes[2] = axes[2][:768]
    vec = gsb.summate_incompr_structured(tc, t1, t2, axes)
    assert tuple(vec.shape) == (3, 1024, 1024, 768)
    nv = 1024 * 1024 * 768
    idx = np.unique(np.concatenate([rs.randint(0, nv, 5000), [0, nv - 1]]))
    gotv = vec.reshape(3, -1)[:, torch.tensor(idx, device=dev)].cpu().numpy()
    wantv = oracle_mod.summate_incompr(cfg["cov"][:, :64], cfg["z1"][:64], cfg["z2"][:64],
                                       bc.grid_points([ax[0], ax[1], ax[2][:768]], None, idx))
    assert np.max(np.abs(gotv - wantv)) * np.sqrt(1.0 / 64) <= TOL


def test_config2_faces_rows_and_full_grid(gsb, oracle_mod):
    """SURVEY.md 8(d) "Parity check" for C2: all six faces and a random 1 % of the rows of the 512^3 field
    against the CPU oracle, and the FULL grid once against the direct kernel (independent arithmetic:
    per-point polynomial sincos instead of per-axis tables)."""
    import torch

    cfg = bc.config2(512)
    edge = 512
    dev = torch.device("cuda:0")
    tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
    axes = [torch.tensor(a, device=dev) for a in cfg["axes"]]
    out = gsb.summate_structured(tc, t1, t2, axes)
    scale = np.sqrt(1.0 / 1000)
    ax = cfg["axes"]
    worst = 0.0

    def oracle_at(i0, i1, i2):
        pos = np.stack([ax[0][i0], ax[1][i1], ax[2][i2]]).astype(np.float64)
        return oracle_mod.summate(cfg["cov"], cfg["z1"], cfg["z2"], pos)

    # six faces, 512 x 512 nodes each
    a, b = np.meshgrid(np.arange(edge), np.arange(edge), indexing="ij")
    a, b = a.reshape(-1), b.reshape(-1)
    for axis in range(3):
        for side in (0, edge - 1):
            ijk = [a, b]
            ijk.insert(axis, np.full(a.size, side))
            got = out[tuple(torch.tensor(v, device=dev) for v in ijk)].cpu().numpy()
            d = float(np.max(np.abs(got - oracle_at(*ijk)))) * scale
            assert d <= TOL, (axis, side, d)
            worst = max(worst, d)
    # a random 1 % of the 262 144 rows (a row = all 512 nodes along the last axis)
    rs = np.random.RandomState(2)
    rows = np.unique(rs.randint(0, edge * edge, edge * edge // 100))
    i0, i1 = np.repeat(rows // edge, edge), np.repeat(rows % edge, edge)
    i2 = np.tile(np.arange(edge), rows.size)
    got = out.reshape(edge * edge, edge)[torch.tensor(rows, device=dev)].reshape(-1).cpu().numpy()
    d = float(np.max(np.abs(got - oracle_at(i0, i1, i2)))) * scale
    assert d <= TOL, d
    worst = max(worst, d)
    # the full grid against the direct kernel, slab by slab (134 217 728 nodes)
    worst_direct = 0.0
    yy, zz = torch.meshgrid(axes[1], axes[2], indexing="ij")
    yy, zz = yy.reshape(-1), zz.reshape(-1)
    planes = 32
    for x0 in range(0, edge, planes):
        xs = axes[0][x0:x0 + planes]
        pos = torch.stack([xs.repeat_interleave(edge * edge), yy.repeat(planes), zz.repeat(planes)])
        direct = gsb.summate(tc, t1, t2, pos).reshape(planes, edge, edge)
        worst_direct = max(worst_direct, float((direct - out[x0:x0 + planes]).abs().max()) * scale)
        del pos, direct
    assert worst_direct <= TOL, worst_direct
    print(f"C2 512^3: faces + 1% rows vs oracle max|d|/sqrt(var) = {worst:.2e}; "
          f"full grid vs direct kernel = {worst_direct:.2e}")
