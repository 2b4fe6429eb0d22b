#!/usr/bin/env python
"""Layered meshes (short last axis): separable path with and without folding the last two axes, and the direct path."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc
import gstools_b200 as gsb
dev = torch.device("cuda:0")
cfg = bc.config2(512)
tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for shape in [(1000, 1000, 10), (1000, 1000, 20), (2000, 500, 5), (512, 512, 48)]:
    axes = [torch.arange(float(s), device=dev, dtype=torch.float64) for s in shape]
    pairs = np.prod(shape) * 1000
    res = []
    for name, opts in [("unfolded", {"fold_axes": 0, "force_path": 2}), ("folded", {"fold_axes": 2, "force_path": 2}),
                       ("direct", {"force_path": 1}), ("auto", {"fold_axes": 1, "force_path": 0})]:
        for k, v in opts.items():
            gsb.set_option(k, v)
        t = timeit(lambda: gsb.summate_structured(tc, t1, t2, axes))
        res.append(f"{name} {t:.2f} ms ({pairs / t / 1e9:.2f} Tpair/s)")
        gsb.set_option("fold_axes", 1); gsb.set_option("force_path", 0)
    print(shape, " | ".join(res))
