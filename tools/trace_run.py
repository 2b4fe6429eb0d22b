import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc
import gstools_b200 as gsb
dev = torch.device("cuda:0")
cfg = bc.config2(512)
tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
axes = [torch.tensor(a, device=dev) for a in cfg["axes"]]
for _ in range(2):
    out = gsb.summate_structured(tc, t1, t2, axes)
torch.cuda.synchronize()
gsb.set_option("trace", 1)
out = gsb.summate_structured(tc, t1, t2, axes)
torch.cuda.synchronize()
