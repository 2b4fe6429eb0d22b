"""Direct-kernel launch configurations (option direct_cfg: 2 = 8 points/thread x 128 threads, 1 = 2 x 128, 0 = 1 x 64).
The sweep that chose cfg 2 also tried 4 x 256 (the first version: -3..8 %), 8 x 128 x 3 CTAs, 12 x 128, 16 x 64, 8 x 64, 8 x 256."""
import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench_configs as bc
import gstools_b200 as gsb
dev = torch.device("cuda:0")
def timeit(fn, reps=4):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)
c3 = bc.config3(8_000_000)
pos = torch.tensor(c3["pos"], device=dev)
m = [torch.tensor(np.ascontiguousarray(c3[k][..., :2000]), device=dev) for k in ("cov", "z1", "z2")]
c4 = bc.config4(8)
m4 = [torch.tensor(c4[k], device=dev) for k in ("cov", "z1", "z2")]
pos3 = torch.rand((3, 8_000_000), device=dev, dtype=torch.float64) * 256
for cfg in (2, 1, 0):
    gsb.set_option("direct_cfg", cfg)
    t2 = timeit(lambda: gsb.summate(m[0], m[1], m[2], pos))
    t3 = timeit(lambda: gsb.summate(m4[0], m4[1], m4[2], pos3))
    tv = timeit(lambda: gsb.summate_incompr(m4[0], m4[1], m4[2], pos3))
    print(f"cfg {cfg}: 2D {8e6*2000/t2/1e12:.3f}  3D {8e6*1000/t3/1e12:.3f}  incompr3D {8e6*1000/tv/1e12:.3f} Tpair/s")
