import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc
import gstools_b200 as gsb
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "3d"
if which == "3d":
    cfg = bc.config2(512)
    tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
    axes = [torch.tensor(a, device=dev) for a in cfg["axes"]]
else:   # same number of rows / columns, but 2-D: the A operand needs no complex product
    cov, z1, z2 = bc.synth_mode_set(2, 1000, 3)
    tc, t1, t2 = (torch.tensor(a, device=dev) for a in (cov, z1, z2))
    axes = [torch.arange(262144.0, device=dev, dtype=torch.float64) * 0.01, torch.arange(512.0, device=dev, dtype=torch.float64)]
for _ in range(2):
    out = gsb.summate_structured(tc, t1, t2, axes)
torch.cuda.synchronize()
gsb.set_option("trace", 1)
out = gsb.summate_structured(tc, t1, t2, axes)
torch.cuda.synchronize()
