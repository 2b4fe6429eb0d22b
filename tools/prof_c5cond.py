"""cProfile of one conditioned realisation of config 5 through gs.CondSRF with the plugin (GPU box)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import refharness  # noqa: E402

gs = refharness.import_gstools()
import gstools_b200 as gsb  # noqa: E402

edge = 128
gsb.enable()
rs = np.random.RandomState(20170519)
cond_pos = rs.uniform(0, edge - 1, (3, 1000))
cond_val = rs.normal(size=1000)
model = gs.Exponential(dim=3, var=1, len_scale=10)
krige = gs.krige.Ordinary(model, cond_pos, cond_val)
crf = gs.CondSRF(krige)
crf.set_pos([np.arange(float(edge))] * 3, "structured")
seeds = gs.random.MasterRNG(20170519)
for _ in range(5):
    crf(seed=seeds(), store=["fld", False, False])
for kw in (dict(), dict(krige_store=False)):
    t0 = time.perf_counter()
    for _ in range(40):
        crf(seed=seeds(), store=["fld", False, False], **kw)
    print(f"{kw}: {(time.perf_counter() - t0) / 40 * 1e3:.3f} ms per realisation", flush=True)
t0 = time.perf_counter()
for _ in range(40):
    crf.generator.update(model, seeds())
print(f"generator.update alone: {(time.perf_counter() - t0) / 40 * 1e3:.3f} ms", flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(40):
    crf(seed=seeds(), store=["fld", False, False])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
