#!/usr/bin/env python
"""BASELINE.json configs[4] through the UNMODIFIED reference with the B200 backend enabled:
CondSRF ensemble on a 128^3 mesh, ordinary kriging on 1000 synthetic conditioning points."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import refharness
gs = refharness.import_gstools()
import gstools_b200 as gsb
edge = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n_cond = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
gsb.enable()
rs = np.random.RandomState(20170519)
cond_pos = rs.uniform(0, edge - 1, (3, n_cond))
cond_val = rs.normal(size=n_cond)
model = gs.Exponential(dim=3, var=1, len_scale=10)
t0 = time.perf_counter()
krige = gs.krige.Ordinary(model, cond_pos, cond_val)
print(f"Ordinary kriging setup ({n_cond} points, pinv on the host): {time.perf_counter()-t0:.2f} s")
t0 = time.perf_counter()
crf = gs.CondSRF(krige)
print(f"CondSRF construction: {time.perf_counter()-t0:.2f} s")
axes = [np.arange(float(edge))] * 3
seeds = gs.random.MasterRNG(20170519)
crf.set_pos(axes, "structured")
for i in range(n_seeds):
    t0 = time.perf_counter()
    k0 = gsb.get_counter("krige_calls")
    f = crf(seed=seeds(), store=[f"fld{i}", False, False])
    t = time.perf_counter() - t0
    print(f"realisation {i}: {t*1e3:.1f} ms  (kriging evaluations in this call: {gsb.get_counter('krige_calls')-k0})  "
          f"f[0,0,0]={f[0,0,0]:.6f}")
# conditioning honoured at the nearest mesh nodes? (informal check: kriging variance small near data)
idx = tuple(np.clip(np.rint(cond_pos).astype(int), 0, edge - 1))
print("mean |field - cond_val| at the mesh nodes nearest to the data:", float(np.mean(np.abs(f[idx] - cond_val))))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
crf(seed=seeds(), store=["fldx", False, False])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
