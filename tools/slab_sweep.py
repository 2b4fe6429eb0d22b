#!/usr/bin/env python
"""One rank's share of C2 at 8 GPUs (a 64 x 512 x 512 slab): which chunk plan / variant is fastest?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc
import gstools_b200 as gsb
dev = torch.device("cuda:0")
cfg = bc.config2(512)
tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
axes = [torch.tensor(cfg["axes"][0][:planes], device=dev)] + [torch.tensor(a, device=dev) for a in cfg["axes"][1:]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(reps=20):
    for _ in range(3):
        gsb.summate_structured(tc, t1, t2, axes)
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); gsb.summate_structured(tc, t1, t2, axes); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
ideal = planes * 512 * 512 * 1000 * 2 / gsb.measure_fp64_peak(0, 0, 0.3) * 1e3
print(f"slab {planes} x 512 x 512: ideal at the measured FP64 peak {ideal:.3f} ms")
for name, opts in [("default", {}), ("scaled", {"sep_path": 2}), ("min_chunks=2", {"min_chunks": 2}),
                   ("min_chunks=1", {"min_chunks": 1}), ("growth=200", {"chunk_growth_pct": 200}),
                   ("growth=300", {"chunk_growth_pct": 300}), ("min_chunks=8", {"min_chunks": 8})]:
    for k, v in opts.items():
        gsb.set_option(k, v)
    t = timeit()
    print(f"{name:14s} {t:.3f} ms  ({100 * ideal / t:.1f} % of peak)")
    for k in opts:
        gsb.set_option(k, {"sep_path": 0, "min_chunks": 4, "chunk_growth_pct": 140}[k])
