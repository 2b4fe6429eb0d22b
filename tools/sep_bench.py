#!/usr/bin/env python
"""Correctness spot-check + timing of the separable kernel variants (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc  # noqa: E402
import gstools_b200 as gsb  # noqa: E402
import oracle  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)


peak = gsb.measure_fp64_peak(0, 0, 0.3)
print(f"DFMA peak {peak/1e12:.3f} TFMA/s")
for variant in (0, 1, 2):
    gsb.set_option("sep_path", variant)
    name = ["auto  ", "agen  ", "scaled"][variant]
    # correctness on a sample
    cfg = bc.config2(128)
    tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
    axes = [torch.tensor(a, device=dev) for a in cfg["axes"]]
    out = gsb.summate_structured(tc, t1, t2, axes).reshape(-1).cpu().numpy()
    idx = np.random.RandomState(0).randint(0, out.size, 3000)
    want = oracle.summate(cfg["cov"], cfg["z1"], cfg["z2"], bc.grid_points(cfg["axes"], None, idx))
    print(f"{name}: max|d|*scale = {np.max(np.abs(out[idx]-want))*np.sqrt(1/1000):.2e}")
    for edge in (256, 512):
        cfg = bc.config2(edge)
        axes = [torch.tensor(a, device=dev) for a in cfg["axes"]]
        t = timeit(lambda: gsb.summate_structured(tc, t1, t2, axes))
        pairs = edge ** 3 * 1000
        print(f"{name} scalar {edge}^3: {t*1e3:.2f} ms {pairs/t/1e12:.3f} Tpair/s  {2*pairs/t/peak*100:.1f}% of DFMA peak")
    c4 = bc.config4(256)
    m4 = [torch.tensor(c4[k], device=dev) for k in ("cov", "z1", "z2")]
    axes = [torch.tensor(a, device=dev) for a in c4["axes"]]
    t = timeit(lambda: gsb.summate_incompr_structured(*m4, axes))
    pairs = 256 ** 3 * 1000
    print(f"{name} incompr 256^3: {t*1e3:.2f} ms {pairs/t/1e12:.3f} Tpair/s  {6*pairs/t/peak*100:.1f}% of DFMA peak")
    c5 = bc.config5(128, 64)
    m5 = [torch.tensor(c5[k], device=dev) for k in ("cov", "z1", "z2")]
    axes = [torch.tensor(a, device=dev) for a in c5["axes"]]
    t = timeit(lambda: gsb.summate_structured(*m5, axes), reps=3)
    pairs = 64 * 128 ** 3 * 1000
    print(f"{name} ensemble 64x128^3: {t*1e3:.2f} ms {pairs/t/1e12:.3f} Tpair/s  {2*pairs/t/peak*100:.1f}% of DFMA peak")
