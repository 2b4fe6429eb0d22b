#!/usr/bin/env python
"""Meshes off the 512^3 sweet spot through the stream-K contraction (gsb_sepk.cuh): device time per call (bursts of
10 calls), the contraction kernel alone, Tpair/s and the fraction of the measured DFMA peak (2 DFMA per pair).
The first-generation numbers for the same shapes are in profiles/r01_odd_mesh_bench.log and
profiles/r02_odd_mesh_gen1_vs_gen2.log."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc
import gstools_b200 as gsb
dev = torch.device("cuda:0")
cfg = bc.config2(512)
tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
peak = gsb.measure_fp64_peak(0, 0, 0.3)
def timeit(fn, reps=7, burst=10):
    """Device time per call, `burst` calls enqueued back to back per bracket (throughput of a stream of calls:
    the host-side enqueue cost of one call overlaps the device work of the previous one)."""
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(burst): fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / burst)
    return float(np.median(ts))
shapes = [(512, 512, 512), (64, 512, 512), (200, 200, 200), (300, 300, 300), (256, 256, 130), (100, 100, 100),
          (128, 128, 128), (400, 400, 72), (512, 16, 512), (2048, 2048), (1000, 1000, 10)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
print(f"DFMA peak {peak / 1e12:.2f} TFMA/s")
for shape in shapes:
    axes = [torch.arange(float(s), device=dev, dtype=torch.float64) for s in shape]
    c = tc[:len(shape)].contiguous()
    pairs = np.prod(shape) * 1000
    res = []
    for name in ("stream-K",):
        gsb.set_option("force_path", 2)
        t = timeit(lambda: gsb.summate_structured(c, t1, t2, axes))
        gsb.set_option("time_kernels", 1); gsb.kernel_times()
        for _ in range(5): gsb.summate_structured(c, t1, t2, axes)
        torch.cuda.synchronize(); km, kn = gsb.kernel_times(); gsb.set_option("time_kernels", 0)
        res.append(f"{name} {t:.3f} ms ({km / 5:.3f} in {kn // 5} contraction launches) {pairs / t / 1e9:.2f} Tpair/s "
                   f"{2 * pairs / (t * 1e-3) / peak * 100:.1f}% (kernel {2 * pairs / (km / 5 * 1e-3) / peak * 100:.1f}%)")
    gsb.set_option("force_path", 0)
    print("x".join(str(v) for v in shape), " | ".join(res), flush=True)
