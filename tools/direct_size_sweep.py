import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench_configs as bc, gstools_b200 as gsb
dev = torch.device("cuda:0")
c3 = bc.config3(4_000_000)
m = [torch.tensor(np.ascontiguousarray(c3[k][..., :2000]), device=dev) for k in ("cov", "z1", "z2")]
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e-3)
    return min(ts)
for n in (10_000, 100_000, 300_000, 400_000, 620_000, 1_000_000, 2_000_000, 4_000_000):
    pos = torch.tensor(c3["pos"][:, :n], device=dev)
    res=[]
    for cfg in (-1, 0, 1, 2):
        gsb.set_option("direct_cfg", cfg)
        t = timeit(lambda: gsb.summate(m[0], m[1], m[2], pos))
        res.append(n*2000/t/1e12)
    print(f"n={n:8d}: auto {res[0]:.3f} | cfg0 {res[1]:.3f} cfg1 {res[2]:.3f} cfg2 {res[3]:.3f} Tpair/s")
