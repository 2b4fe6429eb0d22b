#!/usr/bin/env python
"""Quick GPU-side correctness + timing probe (development aid; the real tests are in tests/)."""
import glob
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import gstools_b200 as gsb  # noqa: E402
import oracle  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


def modes(dim, N, seed, ell=10.0):
    rs = np.random.RandomState(seed)
    z1, z2 = rs.normal(size=N), rs.normal(size=N)
    v = rs.normal(size=(dim, N))
    v /= np.linalg.norm(v, axis=0)
    rad = np.abs(rs.standard_cauchy(N)) / ell if dim > 1 else np.abs(rs.normal(size=N)) / ell
    rad = np.minimum(rad, 5.0)
    return (rad * v), z1, z2


def main():
    print("device:", torch.cuda.get_device_name(0), "count", gsb.device_count())
    ok = True
    # 1. goldens
    for f in sorted(glob.glob(os.path.join(REPO, "tests/golden/*.npz"))):
        if "config_modes" in f:
            continue
        d = np.load(f)
        meta = json.loads(str(d["meta"]))
        fn = gsb.summate if meta["kind"] == "scalar" else gsb.summate_incompr
        got = fn(d["cov_samples"], d["z_1"], d["z_2"], d["pos"])
        err = rel(got, d["raw"])
        scale = np.sqrt(meta["var"] / d["cov_samples"].shape[1])
        flag = err * scale <= 1e-9 * np.sqrt(meta["var"])
        ok &= flag
        print(f"golden {meta['name']:36s} max|d|*scale={err*scale:.3e} {'OK' if flag else 'FAIL'}")
    # 2. random flat cases vs oracle
    for dim in (1, 2, 3, 4, 5):
        for n, N in ((1, 7), (37, 100), (1000, 129), (5000, 1000), (300000, 260)):
            cov, z1, z2 = modes(dim, N, 100 + dim)
            pos = np.random.RandomState(n).uniform(-100, 500, (dim, n))
            want = oracle.summate(cov, z1, z2, pos)
            got = gsb.summate(cov, z1, z2, pos)
            e = rel(got, want) / np.sqrt(N)
            flag = e <= 1e-9
            ok &= flag
            msg = f"flat d={dim} n={n} N={N} err/sqrtN={e:.2e}"
            if dim in (2, 3):
                want = oracle.summate_incompr(cov, z1, z2, pos)
                got = gsb.summate_incompr(cov, z1, z2, pos)
                e2 = rel(got, want) / np.sqrt(N)
                flag2 = e2 <= 1e-9
                ok &= flag2
                msg += f" incompr={e2:.2e}"
                flag &= flag2
            print(msg, "OK" if flag else "FAIL")
    # 3. structured vs oracle on expanded grid
    for dim, lens, N in ((2, (100, 100), 1000), (2, (300, 517), 200), (3, (40, 50, 130), 300),
                         (3, (64, 64, 256), 1000), (4, (9, 10, 11, 140), 64), (3, (5, 6, 7), 50)):
        cov, z1, z2 = modes(dim, N, 7 + dim)
        rs = np.random.RandomState(5)
        axes = [np.sort(rs.uniform(0, 200, L)) for L in lens]
        M = rs.normal(size=(dim, dim))
        grid = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
        pos = M @ grid
        want = oracle.summate(cov, z1, z2, pos).reshape(lens)
        for force in (1, 2):
            gsb.set_option("force_path", force)
            got = gsb.summate_structured(cov, z1, z2, axes, M)
            e = rel(got, want) / np.sqrt(N)
            flag = e <= 1e-9
            ok &= flag
            msg = f"struct d={dim} lens={lens} N={N} path={'direct' if force==1 else 'separable'} err/sqrtN={e:.2e}"
            if dim in (2, 3):
                wv = oracle.summate_incompr(cov, z1, z2, pos).reshape((dim,) + tuple(lens))
                gv = gsb.summate_incompr_structured(cov, z1, z2, axes, M)
                e2 = rel(gv, wv) / np.sqrt(N)
                flag2 = e2 <= 1e-9
                ok &= flag2
                flag &= flag2
                msg += f" incompr={e2:.2e}"
            print(msg, "OK" if flag else "FAIL")
        gsb.set_option("force_path", 0)
    # batched structured
    B, dim, lens, N = 3, 3, (20, 30, 130), 100
    ms = [modes(dim, N, 50 + b) for b in range(B)]
    cov = np.stack([m[0] for m in ms]); z1 = np.stack([m[1] for m in ms]); z2 = np.stack([m[2] for m in ms])
    axes = [np.linspace(0, 50, L) for L in lens]
    grid = np.stack([g.reshape(-1) for g in np.meshgrid(*axes, indexing="ij")])
    gsb.set_option("force_path", 2)
    got = gsb.summate_structured(cov, z1, z2, axes)
    gsb.set_option("force_path", 0)
    for b in range(B):
        e = rel(got[b], oracle.summate(cov[b], z1[b], z2[b], grid).reshape(lens)) / np.sqrt(N)
        ok &= e <= 1e-9
        print(f"batched struct b={b} err/sqrtN={e:.2e}")
    print("ALL OK" if ok else "SOME FAILED")

    # 4. peaks
    for kind, name in ((0, "DFMA"), (1, "DMMA m8n8k4")):
        p = gsb.measure_fp64_peak(0, kind, 0.5)
        print(f"fp64 peak {name}: {p/1e12:.3f} TFMA/s = {2*p/1e12:.2f} TFLOP/s")

    # 5. timings, device-resident
    dev = torch.device("cuda:0")

    def timeit(fn, reps=3):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return min(ts)

    for dim, n, N in ((2, 20_000_000, 1000), (3, 16_777_216, 1000), (2, 2_000_000, 10000)):
        cov, z1, z2 = modes(dim, N, 1)
        tc, t1, t2 = (torch.tensor(a, device=dev) for a in (cov, z1, z2))
        pos = torch.rand((dim, n), device=dev, dtype=torch.float64) * 1000
        t = timeit(lambda: gsb.summate(tc, t1, t2, pos))
        print(f"direct scalar d={dim} n={n} N={N}: {t*1e3:.2f} ms  {n*N/t/1e12:.3f} Tpair/s")
        if dim == 3:
            t = timeit(lambda: gsb.summate_incompr(tc, t1, t2, pos))
            print(f"direct incompr d={dim} n={n} N={N}: {t*1e3:.2f} ms  {n*N/t/1e12:.3f} Tpair/s")
    for lens, N in (((256, 256, 256), 1000), ((512, 512, 512), 1000), ((4096, 4096), 1000)):
        dim = len(lens)
        cov, z1, z2 = modes(dim, N, 1)
        tc, t1, t2 = (torch.tensor(a, device=dev) for a in (cov, z1, z2))
        axes = [torch.arange(L, device=dev, dtype=torch.float64) for L in lens]
        n = int(np.prod(lens))
        t = timeit(lambda: gsb.summate_structured(tc, t1, t2, axes))
        print(f"separable scalar lens={lens} N={N}: {t*1e3:.2f} ms  {n*N/t/1e12:.3f} Tpair/s")
        if dim == 3 and lens[0] == 256:
            t = timeit(lambda: gsb.summate_incompr_structured(tc, t1, t2, axes))
            print(f"separable incompr lens={lens} N={N}: {t*1e3:.2f} ms  {n*N/t/1e12:.3f} Tpair/s")
    # host e2e for C2
    lens, N = (512, 512, 512), 1000
    cov, z1, z2 = modes(3, N, 1)
    axes = [np.arange(L, dtype=np.float64) for L in lens]
    for i in range(3):
        t0 = time.perf_counter(); out = gsb.summate_structured(cov, z1, z2, axes); t = time.perf_counter() - t0
        print(f"host e2e C2 iter {i}: {t*1e3:.1f} ms  {np.prod(lens)*N/t/1e12:.3f} Tpair/s")
        del out
    print("launches:", gsb.get_counter("launches"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
