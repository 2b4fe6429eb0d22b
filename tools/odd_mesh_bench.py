#!/usr/bin/env python
"""Meshes whose last axis is not a multiple of 128: contraction with and without partial-tile skipping."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc
import gstools_b200 as gsb
dev = torch.device("cuda:0")
cfg = bc.config2(512)
tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
for shape in [(512, 512, 512), (200, 200, 200), (300, 300, 300), (256, 256, 130), (100, 100, 100), (128, 128, 128), (400, 400, 72)]:
    axes = [torch.arange(float(s), device=dev, dtype=torch.float64) for s in shape]
    pairs = np.prod(shape) * 1000
    res = []
    for name, v in (("full-tile kernel", 0), ("partial-tile kernel", 1)):
        gsb.set_option("partial_tiles", v)
        t = timeit(lambda: gsb.summate_structured(tc, t1, t2, axes))
        res.append(f"{name} {t:.3f} ms ({pairs / t / 1e9:.2f} Tpair/s)")
    gsb.set_option("partial_tiles", 1)
    print(shape, " | ".join(res))
