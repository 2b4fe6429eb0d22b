"""Multi-GPU path on real devices (NCCL): skipped unless the box has >= 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    import gstools_b200 as gsb
    from conftest import synth_modes
    from gstools_b200 import dist as gdist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        cov, z1, z2 = synth_modes(3, 200, seed=5)
        tc, t1, t2 = (torch.tensor(a, device=dev) for a in (cov, z1, z2))
        axes = [torch.arange(float(n), device=dev, dtype=torch.float64) for n in (67, 64, 256)]
        gsb.set_option("force_path", 2)
        slab, (lo, hi) = gdist.summate_structured_sharded(tc, t1, t2, axes)
        assert slab.device == dev and tuple(slab.shape) == (hi - lo, 64, 256)
        full = gdist.gather_field(slab, 67)                      # NCCL all-gather
        only0 = gdist.gather_field(slab, 67, dst=0)              # NCCL gather to rank 0
        single = gsb.summate_structured(tc, t1, t2, axes)        # whole mesh on this GPU
        # slabs vs the whole mesh: the same tiles, other stream-K shares -> equal up to rounding (1e-12 * sqrt(var))
        tight = 1e-12 * np.sqrt(200.0)
        assert float((full - single).abs().max()) <= tight
        assert (only0 is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(only0, full)
        # sum + gather onto rank 0 as one operation: NCCL pieces on a side stream / stores into rank 0's memory
        for mode, kw in (("nccl", dict(pieces=1)), ("nccl", dict(pieces=4)), ("p2p", {})):
            got = gdist.summate_structured_gathered(tc, t1, t2, axes, dst=0, mode=mode, **kw)
            assert (got is None) == (rank != 0)
            if rank == 0:
                torch.cuda.synchronize()
                assert tuple(got.shape) == (67, 64, 256)
                assert float((got - single).abs().max()) <= tight, mode
            gv = gdist.summate_structured_gathered(tc, t1, t2, axes, dst=world - 1, mode=mode, incompr=True, **kw)
            if rank == world - 1:
                torch.cuda.synchronize()
                vec = gsb.summate_incompr_structured(tc, t1, t2, axes)
                assert tuple(gv.shape) == (3, 67, 64, 256)
                assert float((gv - vec).abs().max()) <= tight, mode
        everywhere = gdist.summate_structured_gathered(tc, t1, t2, axes, dst=None, mode="nccl", pieces=3)
        assert float((everywhere - single).abs().max()) <= tight
        dist.barrier()
        gdist.close_peer_fields()
        # flat points, host arrays in: every rank uses its own device (LOCAL_RANK)
        gsb.set_option("force_path", 0)
        pos = np.random.RandomState(1).uniform(0, 100, (3, 100001))
        local, (lo, hi) = gdist.summate_sharded(cov, z1, z2, pos)
        whole = gsb.summate(cov, z1, z2, pos)
        assert np.array_equal(local, whole[lo:hi])
        # kriging evaluation (row f1): slabs along axis 0, the system replicated on every device
        rs = np.random.RandomState(7)
        cpos = rs.uniform(0, 30, (3, 40))
        kmat, kcond = rs.normal(size=(41, 41)) / 40, np.concatenate([rs.normal(size=40), [0.0]])
        spec = dict(kind="Exponential", var=1.0, len_rescaled=6.0)
        kaxes = [np.arange(33.0), np.arange(16.0), np.arange(24.0)]
        (lf, le), (lo, hi) = gdist.krige_evaluate_sharded(spec, kmat, kcond, cpos, axes=kaxes)
        wf, we = gsb.krige_evaluate(spec, kmat, kcond, cpos, axes=kaxes)
        assert np.array_equal(lf, wf[lo:hi]) and np.array_equal(le, we[lo:hi])
        gathered = gdist.gather_field(torch.tensor(lf, device=dev), 33)
        assert torch.equal(gathered.cpu(), torch.tensor(wf))
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_nccl_sharded_structured_and_gather(tmp_path):
    import torch.multiprocessing as mp

    world = min(_ngpu(), 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_one_process_uses_several_devices():
    """Kernel attributes (dynamic shared memory, carve-out) are per device: the same process must be
    able to run every kernel family on device 0 and then on device 1, with identical bits."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import gstools_b200 as gsb
    from conftest import synth_modes

    cov, z1, z2 = synth_modes(3, 200, seed=5)
    axes = [np.arange(20.0), np.arange(140.0), np.arange(150.0)]
    rs = np.random.RandomState(3)
    kmat, kcond, kv = rs.normal(size=(90, 90)), rs.normal(size=90), rs.uniform(-1, 1, (90, 3000))
    cpos = rs.uniform(0, 20, (3, 89))
    spec = dict(kind="Exponential", var=1.0, len_rescaled=5.0)
    results = []
    try:
        for dev in (0, 1):
            gsb.set_device(dev)
            a = gsb.summate_structured(cov, z1, z2, axes)
            b = gsb.summate_incompr_structured(cov, z1, z2, axes)
            c = np.stack(gsb.calc_field_krige_and_variance(kmat, kv, kcond))
            d = np.stack(gsb.krige_evaluate(spec, kmat, kcond, cpos, axes=[a_[:12] for a_ in axes]))
            e = gsb.summate(cov, z1, z2, rs.uniform(0, 50, (3, 0)) if False else np.ones((3, 5000)))
            results.append((a, b, c, d, e))
    finally:
        gsb.set_device(0)
    for x, y in zip(*results):
        assert np.array_equal(x, y)
