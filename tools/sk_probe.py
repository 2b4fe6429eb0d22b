#!/usr/bin/env python
"""Probe of the stream-K contraction: kernel-only time of one shape for several grid sizes (sk_grid option)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc
import gstools_b200 as gsb
dev = torch.device("cuda:0")
cfg = bc.config2(512)
tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
peak = gsb.measure_fp64_peak(0, 0, 0.3)
shape = tuple(int(v) for v in sys.argv[1].split("x"))
grids = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
axes = [torch.arange(float(s), device=dev, dtype=torch.float64) for s in shape]
c = tc[:len(shape)].contiguous()
pairs = np.prod(shape) * 1000
gsb.set_option("sep_path", 3); gsb.set_option("force_path", 2)
for g in grids:
    gsb.set_option("sk_grid", g)
    for _ in range(3): gsb.summate_structured(c, t1, t2, axes)
    torch.cuda.synchronize()
    gsb.set_option("time_kernels", 1); gsb.kernel_times()
    for _ in range(10): gsb.summate_structured(c, t1, t2, axes)
    torch.cuda.synchronize(); km, kn = gsb.kernel_times(); gsb.set_option("time_kernels", 0)
    print(f"{sys.argv[1]} grid {g or 148}: {km / kn:.4f} ms per launch, {2 * pairs / (km / 10 * 1e-3) / peak * 100:.1f}% of DFMA peak", flush=True)
