#!/usr/bin/env python
"""Tiny driver for ncu captures: structured C2-like call (edge^3 mesh) and direct calls."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc  # noqa: E402
import gstools_b200 as gsb  # noqa: E402

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 256
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
dev = torch.device("cuda:0")
cfg = bc.config2(edge)
tc, t1, t2 = (torch.tensor(cfg[k], device=dev) for k in ("cov", "z1", "z2"))
axes = [torch.tensor(a, device=dev) for a in cfg["axes"]]
for _ in range(2):
    out = gsb.summate_structured(tc, t1, t2, axes)
torch.cuda.synchronize()
c3 = bc.config3(npts)
pos = torch.tensor(c3["pos"], device=dev)
m = [torch.tensor(np.ascontiguousarray(c3[k][..., :2000]), device=dev) for k in ("cov", "z1", "z2")]
for _ in range(2):
    o = gsb.summate(m[0], m[1], m[2], pos)
pos3 = torch.rand((3, npts), device=dev, dtype=torch.float64) * 256
c4 = bc.config4(8)
m4 = [torch.tensor(c4[k], device=dev) for k in ("cov", "z1", "z2")]
o = gsb.summate_incompr(m4[0], m4[1], m4[2], pos3)
torch.cuda.synchronize()
print("done", float(out.sum()), float(o.sum()))
