import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc
import gstools_b200 as gsb
dev = torch.device("cuda:0")
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)
peak = gsb.measure_fp64_peak(0, 0, 0.3)
c3 = bc.config3(8_000_000)
pos = torch.tensor(c3["pos"], device=dev)
m = [torch.tensor(np.ascontiguousarray(c3[k][..., :2000]), device=dev) for k in ("cov", "z1", "z2")]
c4 = bc.config4(8)
m4 = [torch.tensor(c4[k], device=dev) for k in ("cov", "z1", "z2")]
pos3 = torch.rand((3, 8_000_000), device=dev, dtype=torch.float64) * 256
for cfg in (2, 1, 0):
    gsb.set_option("direct_cfg", cfg)
    t = timeit(lambda: gsb.summate(m[0], m[1], m[2], pos))
    pr = 8e6 * 2000
    print(f"cfg {cfg} scalar 2D: {t*1e3:.2f} ms {pr/t/1e12:.3f} Tpair/s pipe {pr*15/t/peak*100:.1f}%")
    t = timeit(lambda: gsb.summate(m4[0], m4[1], m4[2], pos3))
    pr = 8e6 * 1000
    print(f"cfg {cfg} scalar 3D: {t*1e3:.2f} ms {pr/t/1e12:.3f} Tpair/s pipe {pr*16/t/peak*100:.1f}%")
    t = timeit(lambda: gsb.summate_incompr(m4[0], m4[1], m4[2], pos3))
    print(f"cfg {cfg} incompr 3D: {t*1e3:.2f} ms {pr/t/1e12:.3f} Tpair/s pipe {pr*19/t/peak*100:.1f}%")
