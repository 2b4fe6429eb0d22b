"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: partition + gather.

The compute function is injected (the CPU oracle stands in for the CUDA kernels), so what is
covered here is exactly the host logic that runs unchanged with NCCL on the GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist

    import oracle
    from conftest import synth_modes
    from gstools_b200 import dist as gdist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cov, z1, z2 = synth_modes(3, 64, seed=5)
        # ---- flat points, uneven split (n = 1001) ----
        pos = np.random.RandomState(0).uniform(0, 100, (3, 1001))
        local, (lo, hi) = gdist.summate_sharded(cov, z1, z2, pos, compute=oracle.summate)
        assert local.shape == (hi - lo,)
        full = gdist.gather_field(local, 1001)
        want = oracle.summate(cov, z1, z2, pos)
        assert np.array_equal(full, want)
        only0 = gdist.gather_field(local, 1001, dst=0)
        assert (only0 is None) == (rank != 0)
        if rank == 0:
            assert np.array_equal(only0, want)
        # ---- vector field: gather along the point axis (axis 1) ----
        lv, _ = gdist.summate_sharded(cov, z1, z2, pos, incompr=True, compute=oracle.summate_incompr)
        fv = gdist.gather_field(lv, 1001, axis=1)
        assert np.array_equal(fv, oracle.summate_incompr(cov, z1, z2, pos))

        # ---- structured slabs along axis 0 (7 x 5 x 4 mesh: uneven 4 + 3) ----
        axes = [np.linspace(0, 9, 7), np.linspace(-3, 3, 5), np.arange(4.0)]

        def struct_compute(c, a, b, ax, matrix):
            grid = np.stack([g.reshape(-1) for g in np.meshgrid(*ax, indexing="ij")])
            return oracle.summate(c, a, b, grid).reshape([len(x) for x in ax])

        slab, (lo, hi) = gdist.summate_structured_sharded(cov, z1, z2, axes, compute=struct_compute)
        assert slab.shape == (hi - lo, 5, 4)
        field = gdist.gather_field(slab, 7)
        assert np.array_equal(field, struct_compute(cov, z1, z2, axes, None))

        # ---- sum and gather pipelined: row pieces of decreasing size, each sent while the next is computed ----
        def struct_vec(c, a, b, ax, matrix):
            grid = np.stack([g.reshape(-1) for g in np.meshgrid(*ax, indexing="ij")])
            return oracle.summate_incompr(c, a, b, grid).reshape([3] + [len(x) for x in ax])

        for pieces in (1, 3, 8):
            g0 = gdist.summate_structured_gathered(cov, z1, z2, axes, dst=0, pieces=pieces, compute=struct_compute)
            assert (g0 is None) == (rank != 0)
            if rank == 0:
                assert np.array_equal(g0.numpy(), struct_compute(cov, z1, z2, axes, None))
            g1 = gdist.summate_structured_gathered(cov, z1, z2, axes, dst=1, pieces=pieces, incompr=True,
                                                   compute=struct_vec)
            if rank == 1:
                assert np.array_equal(g1.numpy(), struct_vec(cov, z1, z2, axes, None))
            ga = gdist.summate_structured_gathered(cov, z1, z2, axes, dst=None, pieces=pieces, compute=struct_compute)
            assert np.array_equal(ga.numpy(), struct_compute(cov, z1, z2, axes, None))
        # one rank without rows (axis 0 shorter than the world)
        short = [axes[0][:1]] + axes[1:]
        gs = gdist.summate_structured_gathered(cov, z1, z2, short, dst=1, pieces=4, compute=struct_compute)
        if rank == 1:
            assert np.array_equal(gs.numpy(), struct_compute(cov, z1, z2, short, None))

        # ---- ensemble: seeds sharded, no collective ----
        sets = [synth_modes(3, 16, seed=s) for s in range(5)]
        mine, (lo, hi) = gdist.ensemble_sharded(sets, lambda m: oracle.summate(*m, pos[:, :50]))
        assert len(mine) == hi - lo and (lo, hi) == gdist.shard_range(5, rank, world)
        for f, s in zip(mine, range(lo, hi)):
            assert np.array_equal(f, oracle.summate(*sets[s], pos[:, :50]))
        # ---- kriging evaluation: points / slabs sharded, system replicated, drift rows follow ----
        rs = np.random.RandomState(3)
        cpos, K = rs.uniform(0, 9, (3, 12)), 12 + 1 + 1
        kmat, kcond = rs.normal(size=(K, K)), np.concatenate([rs.normal(size=12), [0.0, 0.0]])
        spec = dict(kind="Exponential", var=1.3, len_rescaled=4.0)

        def krige_compute(model, m, c, cp, pos=None, axes=None, matrix=None, unbiased=True, tail_rows=None,
                          return_var=True):
            shape = None
            if axes is not None:
                shape = tuple(len(a) for a in axes)
                pos = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
            f, e = oracle.krige_evaluate(model, m, c, cp, pos, unbiased, tail_rows)
            return (f.reshape(shape), e.reshape(shape)) if shape else (f, e)

        grid = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
        drift = grid[0:1] * 0.5                                            # one drift row on the mesh
        (lf, le), (lo, hi) = gdist.krige_evaluate_sharded(spec, kmat, kcond, cpos, axes=axes, tail_rows=drift,
                                                           compute=krige_compute)
        assert lf.shape == (hi - lo, 5, 4)
        wf, we = oracle.krige_evaluate(spec, kmat, kcond, cpos, grid, True, drift)
        assert np.array_equal(gdist.gather_field(lf, 7).reshape(-1), wf)
        assert np.array_equal(gdist.gather_field(le, 7).reshape(-1), we)
        (pf, pe), (lo, hi) = gdist.krige_evaluate_sharded(spec, kmat, kcond, cpos, pos=pos[:, :101] * 0.09,
                                                           tail_rows=pos[0:1, :101], compute=krige_compute)
        wf, we = oracle.krige_evaluate(spec, kmat, kcond, cpos, pos[:, :101] * 0.09, True, pos[0:1, :101])
        assert np.array_equal(gdist.gather_field(pf, 101), wf) and np.array_equal(gdist.gather_field(pe, 101), we)
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    import oracle

    oracle.build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
