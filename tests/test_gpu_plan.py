"""Single-process multi-GPU plan (gsb_plan_*): slabs along axis 0, point ranges and batch
entries fanned out over the plan's devices, host route (every device copies into its slice of ONE host array) and
device route (peers store into the home device's tensor).  A device may be listed more than once, so the share logic
is covered on a one-GPU box; the tests marked `two_gpus` need a second device."""
import numpy as np
import pytest

from conftest import synth_modes

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


def tight(n_modes):
    # other stream-K shares / tile boundaries than the single-device call: equal up to rounding
    return 1e-12 * np.sqrt(n_modes)


DEVICE_LISTS = [[0], [0, 0, 0]] + ([[0, 1]] if _ngpu() >= 2 else [])


@pytest.fixture(params=DEVICE_LISTS, ids=lambda d: "dev" + "".join(map(str, d)))
def plan(request, gsb):
    p = gsb.Plan(request.param)
    yield p
    p.close()
    gsb.set_option("force_path", 0)


def test_plan_info(plan, gsb):
    assert len(plan) == len(plan.devices) >= 1
    shares = [plan.share(10, g) for g in range(len(plan))]
    assert shares[0][0] == 0 and shares[-1][1] == 10
    assert all(a[1] == b[0] for a, b in zip(shares, shares[1:]))


@pytest.mark.parametrize("lens,n_modes,force", [((67, 64, 256), 200, 2), ((5, 40, 130), 64, 2), ((2, 9, 31), 50, 0),
                                                ((33, 300), 100, 2), ((41, 17, 19), 30, 1)])
def test_plan_structured_host_slabs(plan, gsb, lens, n_modes, force):
    dim = len(lens)
    cov, z1, z2 = synth_modes(dim, n_modes, seed=3)
    rs = np.random.RandomState(1)
    axes = [np.sort(rs.uniform(0, 100, L)) for L in lens]
    mat = rs.normal(size=(dim, dim))
    gsb.set_option("force_path", force)
    want = gsb.summate_structured(cov, z1, z2, axes, mat)
    got = plan.summate_structured(cov, z1, z2, axes, mat)
    assert got.shape == tuple(lens)
    assert maxabs(got, want) <= tight(n_modes)
    # vector field: every component of a slab lands in its own strided block of the (dim, n) result
    if dim in (2, 3):
        want = gsb.summate_incompr_structured(cov, z1, z2, axes, mat)
        got = plan.summate_incompr_structured(cov, z1, z2, axes, mat)
        assert got.shape == (dim,) + tuple(lens)
        assert maxabs(got, want) <= tight(n_modes)


def test_plan_structured_batch_shares(plan, gsb):
    """n_batch >= number of devices: whole fields per device (ensemble seeds)."""
    lens, n_modes, nb = (6, 70, 140), 96, 5
    sets = [synth_modes(3, n_modes, seed=20 + b) for b in range(nb)]
    cov = np.stack([s[0] for s in sets])
    z1 = np.stack([s[1] for s in sets])
    z2 = np.stack([s[2] for s in sets])
    axes = [np.arange(float(L)) for L in lens]
    gsb.set_option("force_path", 2)
    got = plan.summate_structured(cov, z1, z2, axes)
    assert got.shape == (nb,) + lens
    for b in range(nb):
        want = gsb.summate_structured(*sets[b], axes)
        assert maxabs(got[b], want) <= tight(n_modes)


def test_plan_flat_points_bit_equal(plan, gsb):
    """Direct kernel: a point's bits do not depend on what else is in the call, so the shares reproduce the single call."""
    cov, z1, z2 = synth_modes(3, 300, seed=9)
    pos = np.random.RandomState(2).uniform(-50, 50, (3, 10007))
    assert np.array_equal(plan.summate(cov, z1, z2, pos), gsb.summate(cov, z1, z2, pos))
    assert np.array_equal(plan.summate_incompr(cov, z1, z2, pos), gsb.summate_incompr(cov, z1, z2, pos))
    epi = gsb.make_epilogue(0.37, [0.5])
    assert np.array_equal(plan.summate(cov, z1, z2, pos, epilogue=epi), gsb.summate(cov, z1, z2, pos, epilogue=epi))
    empty = plan.summate(cov, z1, z2, pos[:, :0])
    assert empty.shape == (0,)


def test_plan_point_epilogue(plan, gsb):
    """Conditioned fields: gain / offset replicated on every device, each share reads its own part of them."""
    import torch

    lens, n_modes = (19, 24, 136), 80
    n = int(np.prod(lens))
    cov, z1, z2 = synth_modes(3, n_modes, seed=4)
    axes = [np.arange(float(L)) for L in lens]
    rs = np.random.RandomState(8)
    gain, offset = rs.uniform(0, 1, n), rs.normal(size=n)
    epi = gsb.make_epilogue(np.sqrt(1.0 / n_modes), [0.0])
    raw = gsb.summate_structured(cov, z1, z2, axes, epilogue=epi).reshape(-1)
    gsb.set_option("force_path", 2)
    pe = plan.make_point_epilogue(gain, offset, [0.0, 1.5])
    got = plan.summate_structured(cov, z1, z2, axes, epilogue=epi, point_epilogue=pe).reshape(-1)
    one = gsb.make_point_epilogue(torch.tensor(gain, device="cuda:0"), torch.tensor(offset, device="cuda:0"), [0.0, 1.5])
    want = gsb.summate_structured(cov, z1, z2, axes, epilogue=epi, point_epilogue=one).reshape(-1)
    assert maxabs(got, want) <= 1e-12
    assert maxabs(got, offset + gain * raw + 0.0 + 1.5) <= 1e-12
    # flat points with the same arrays
    pos = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    flat = plan.summate(cov, z1, z2, pos, epilogue=epi, point_epilogue=pe)
    assert maxabs(flat, want) <= 1e-10          # direct kernel against the separable one: 1e-9 * sqrt(var) class


def test_plan_device_route_stores_into_home_tensor(plan, gsb):
    """CUDA tensors in, ONE CUDA tensor out on the inputs' device; the other devices store into it through peer memory."""
    import torch

    if len(set(plan.devices)) > 1 and not plan.peer_access:
        pytest.skip("no peer access between the devices of this box")
    home = torch.device("cuda", plan.devices[-1])
    lens, n_modes = (37, 48, 256), 120
    cov, z1, z2 = synth_modes(3, n_modes, seed=6)
    axes = [np.arange(float(L)) for L in lens]
    t = [torch.tensor(a, device=home) for a in (cov, z1, z2)]
    tax = [torch.tensor(a, device=home) for a in axes]
    gsb.set_option("force_path", 2)
    with torch.cuda.device(home):
        got = plan.summate_structured(*t, tax)
        got_vec = plan.summate_incompr_structured(*t, tax)
        torch.cuda.synchronize()
    assert got.device == home and tuple(got.shape) == lens
    want = gsb.summate_structured(cov, z1, z2, axes)
    assert maxabs(got.cpu().numpy(), want) <= tight(n_modes)
    assert maxabs(got_vec.cpu().numpy(), gsb.summate_incompr_structured(cov, z1, z2, axes)) <= tight(n_modes)
    pos = torch.tensor(np.random.RandomState(3).uniform(0, 30, (3, 5001)), device=home)
    with torch.cuda.device(home):
        flat = plan.summate(*t, pos)
        torch.cuda.synchronize()
    assert np.array_equal(flat.cpu().numpy(), gsb.summate(cov, z1, z2, pos.cpu().numpy()))


def test_use_devices_routes_big_calls(gsb):
    """The process-wide plan takes over host-array calls above the pair threshold (here: every call)."""
    devices = [0, 1] if _ngpu() >= 2 else [0, 0]
    plan = gsb.Plan(devices)
    from gstools_b200 import backend

    old = dict(backend._PLAN)
    backend._PLAN.update(plan=plan, min_pairs=0.0)
    try:
        cov, z1, z2 = synth_modes(3, 64, seed=1)
        axes = [np.arange(float(L)) for L in (10, 20, 130)]
        got = gsb.summate_structured(cov, z1, z2, axes)
        backend._PLAN.update(plan=None)
        want = gsb.summate_structured(cov, z1, z2, axes)
        assert maxabs(got, want) <= tight(64)
    finally:
        backend._PLAN.update(old)
        plan.close()


def test_srf_slabs_over_all_gpus(gsb):
    """gs.SRF(...)(structured mesh) from ONE process on every GPU of the box (enable(devices="all")): the same field
    as on one GPU.  On a one-GPU box the plan lists that GPU twice."""
    import torch

    import refharness

    if not refharness.have_reference():
        pytest.skip("reference gstools not present")
    gs = refharness.import_gstools()
    gsb.enable()
    try:
        _srf_on_all(gs, gsb, torch)
    finally:
        gsb.disable()


def _srf_on_all(gs, gsb, torch):
    model = gs.Gaussian(dim=3, var=2.0, len_scale=8.0)
    axes = [np.arange(96.0), np.arange(64.0), np.arange(256.0)]
    srf = gs.SRF(model, seed=20170519, mode_no=256)
    one = np.array(srf(axes, mesh_type="structured"))
    plan = gsb.use_devices("all" if torch.cuda.device_count() > 1 else [0, 0], min_pairs=0)
    try:
        assert plan is not None and len(plan) == max(2, torch.cuda.device_count())
        before = [gsb.get_counter("sk_calls")]
        many = np.array(srf(axes, mesh_type="structured"))
        assert gsb.get_counter("sk_calls") - before[0] == len(plan)
    finally:
        gsb.use_devices(None, min_pairs=4e9)
    assert maxabs(one, many) <= 1e-12 * np.sqrt(2.0)


def test_plan_batch_shares_of_vector_fields(plan, gsb):
    """Ensembles of vector fields: whole (batch entry, component) blocks per device."""
    lens, n_modes, nb = (5, 33, 130), 48, 4
    sets = [synth_modes(3, n_modes, seed=40 + b) for b in range(nb)]
    cov, z1, z2 = (np.stack([s[k] for s in sets]) for k in range(3))
    axes = [np.arange(float(L)) for L in lens]
    got = plan.summate_incompr_structured(cov, z1, z2, axes)
    assert got.shape == (nb, 3) + lens
    for b in range(nb):
        assert maxabs(got[b], gsb.summate_incompr_structured(*sets[b], axes)) <= tight(n_modes)


def test_ensemble_over_the_plan_equals_one_device(gsb):
    """gstools_b200.ensemble with the process-wide plan: the seeds are dealt out to the devices, the per-point epilogue
    of the conditioned field is replicated on each; same fields as on one device."""
    import torch

    import refharness

    if not refharness.have_reference():
        pytest.skip("reference gstools not present")
    gs = refharness.import_gstools()
    gsb.enable()
    try:
        rs = np.random.RandomState(5)
        model = gs.Exponential(dim=3, var=0.8, len_scale=5.0)
        krige = gs.krige.Ordinary(model, rs.uniform(0, 20, (3, 25)), rs.normal(size=25))
        crf = gs.CondSRF(krige, mode_no=128)
        axes = [np.arange(20.0), np.arange(24.0), np.arange(130.0)]
        seeds = [11, 12, 13, 14, 15]
        one = gsb.ensemble(crf, seeds, axes, mesh_type="structured")
        plan = gsb.use_devices("all" if torch.cuda.device_count() > 1 else [0, 0], min_pairs=0)
        try:
            before = gsb.get_counter("sk_calls")
            many = gsb.ensemble(crf, seeds, axes, mesh_type="structured")
            assert gsb.get_counter("sk_calls") - before == min(len(plan), len(seeds))
        finally:
            gsb.use_devices(None, min_pairs=4e9)
        assert maxabs(one, many) <= 1e-12
        single = np.array(crf(axes, seed=13, mesh_type="structured", store=False))
        assert maxabs(many[2], single) <= 1e-9 * np.sqrt(0.8)
    finally:
        gsb.disable()


def test_closed_plan_is_refused(gsb):
    p = gsb.Plan([0])
    p.close()
    p.close()
    with pytest.raises(ValueError, match="closed"):
        p.summate_structured(*synth_modes(2, 8, seed=1), [np.arange(4.0), np.arange(5.0)])


def test_plan_krige_evaluate_equals_one_device(plan, gsb):
    """Row f1 over the plan: slabs of a mesh / ranges of a flat point set per device, the kriging system replicated,
    drift rows following their points.  A point's kriging sums do not depend on how the points are cut: same bits."""
    import torch

    rs = np.random.RandomState(7)
    cpos = rs.uniform(0, 30, (3, 40))
    kmat, kcond = rs.normal(size=(42, 42)) / 40, np.concatenate([rs.normal(size=40), [0.0, 0.0]])
    spec = dict(kind="Exponential", var=1.0, len_rescaled=6.0)
    axes = [np.arange(33.0), np.arange(16.0), np.arange(24.0)]
    n = 33 * 16 * 24
    drift = rs.normal(size=(1, n))
    want_f, want_e = gsb.krige_evaluate(spec, kmat, kcond, cpos, axes=axes, tail_rows=drift)
    got_f, got_e = plan.krige_evaluate(spec, kmat, kcond, cpos, axes=axes, tail_rows=drift)
    assert got_f.shape == (33, 16, 24)
    assert np.array_equal(got_f, want_f) and np.array_equal(got_e, want_e)
    only_f = plan.krige_evaluate(spec, kmat, kcond, cpos, axes=axes, tail_rows=drift, return_var=False)
    assert maxabs(only_f, want_f) <= 1e-12
    pos = rs.uniform(0, 30, (3, 5003))
    dr = rs.normal(size=(1, 5003))
    wf, we = gsb.krige_evaluate(spec, kmat, kcond, cpos, pos=pos, tail_rows=dr)
    gf, ge = plan.krige_evaluate(spec, kmat, kcond, cpos, pos=pos, tail_rows=dr)
    assert np.array_equal(gf, wf) and np.array_equal(ge, we)
    # device route: tensors on one device of the plan, the others store their share into them
    if len(set(plan.devices)) == 1 or plan.peer_access:
        home = torch.device("cuda", plan.devices[-1])
        up = lambda a: torch.tensor(a, device=home)            # noqa: E731
        with torch.cuda.device(home):
            df, de = plan.krige_evaluate(spec, up(kmat), up(kcond), up(cpos), axes=[up(a) for a in axes],
                                         tail_rows=up(drift))
            pf, pe = plan.krige_evaluate(spec, up(kmat), up(kcond), up(cpos), pos=up(pos), tail_rows=up(dr))
            torch.cuda.synchronize()
        assert df.device == home
        assert np.array_equal(df.cpu().numpy(), want_f) and np.array_equal(de.cpu().numpy(), want_e)
        assert np.array_equal(pf.cpu().numpy(), wf) and np.array_equal(pe.cpu().numpy(), we)
