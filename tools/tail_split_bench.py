"""One rank's share of config 3 on eight GPUs (2.5 M points x 10^4 modes, 8.25 waves of the big-CTA configuration):
the tail split (full waves + a second launch of small CTAs) against the single big-CTA launch."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench_configs as bc  # noqa: E402
import gstools_b200 as gsb  # noqa: E402

cfg = bc.config3(2_500_000)
dev = torch.device("cuda:0")
t = [torch.tensor(np.ascontiguousarray(cfg[k]), device=dev) for k in ("cov", "z1", "z2", "pos")]
pairs = cfg["pos"].shape[1] * cfg["cov"].shape[1]
for label, opt in (("tail split (auto)", -1), ("one launch, 8 points per thread", 2)):
    gsb.set_option("direct_cfg", opt)
    for _ in range(2):
        gsb.summate(*t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = gsb.summate(*t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{label}: {ms:.3f} ms  {pairs / ms / 1e9:.3f} Tpair/s")
gsb.set_option("direct_cfg", -1)
