#!/usr/bin/env python
"""Driver for the round-2 ncu captures of HEAD: one launch each of
  sk_contract_kernel<true,false>   C5-like batch (4 seeds x 128^3, 1000 modes)
  sk_contract_kernel<true,true>    200^3 (partial row and column tiles)
  sk_contract_kernel<true,false>   C4-like vector field (128^3, 3 components)
  sk_contract_kernel<false,*>      2-D mesh 2048 x 2048 (nothing rescaled)
  direct_kernel<2,8,...>           C3 sample (4 M points x 2000 Matern modes, 2-D)
  direct_kernel<3,8,true,...>      incompressible 3-D, unstructured
No warm-up launches: ncu replays every captured kernel anyway."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc  # noqa: E402
import gstools_b200 as gsb  # noqa: E402

dev = torch.device("cuda:0")


def t(a):
    return torch.tensor(np.ascontiguousarray(a), device=dev)


c5 = bc.config5(128, 4)
out = gsb.summate_structured(t(c5["cov"]), t(c5["z1"]), t(c5["z2"]), [t(a) for a in c5["axes"]])
c2 = bc.config2(200)
out = gsb.summate_structured(t(c2["cov"]), t(c2["z1"]), t(c2["z2"]), [t(a) for a in c2["axes"]])
c4 = bc.config4(128)
out = gsb.summate_incompr_structured(t(c4["cov"]), t(c4["z1"]), t(c4["z2"]), [t(a) for a in c4["axes"]])
c1 = bc.config1()
ax2 = [torch.arange(2048.0, device=dev, dtype=torch.float64)] * 2
out = gsb.summate_structured(t(c1["cov"]), t(c1["z1"]), t(c1["z2"]), ax2)
c3 = bc.config3(4_000_000)
m = [t(c3[k][..., :2000]) for k in ("cov", "z1", "z2")]
o = gsb.summate(m[0], m[1], m[2], t(c3["pos"]))
pos3 = torch.rand((3, 2_000_000), device=dev, dtype=torch.float64) * 256
c4s = bc.config4(8)
o = gsb.summate_incompr(t(c4s["cov"]), t(c4s["z1"]), t(c4s["z2"]), pos3)
torch.cuda.synchronize()
print("done", float(out.sum()), float(o.sum()))
