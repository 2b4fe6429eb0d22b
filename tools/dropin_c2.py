#!/usr/bin/env python
"""Time the user-visible drop-in call gs.SRF(model)(512^3 structured) with the B200 backend."""
import os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import refharness
gs = refharness.import_gstools()
import gstools_b200 as gsb
edge = int(sys.argv[1]) if len(sys.argv) > 1 else 512
gsb.enable()
t0 = time.perf_counter()
srf = gs.SRF(gs.Exponential(dim=3, var=1, len_scale=10), seed=20170519)
print(f"SRF construction (mode sampling, emcee stand-in): {time.perf_counter()-t0:.2f} s")
axes = [np.arange(float(edge))] * 3
for i in range(3):
    t0 = time.perf_counter()
    f = srf.structured(axes)
    t = time.perf_counter() - t0
    print(f"srf.structured({edge}^3) call {i}: {t*1e3:.1f} ms   field[0,0,0]={f[0,0,0]:.6f}")
# breakdown: the summation alone through the public API
from gstools.field import generator as gen
g = srf.generator
for i in range(2):
    t0 = time.perf_counter()
    raw = gsb.summate_structured(g._cov_sample, g._z_1, g._z_2, axes)
    t1 = time.perf_counter()
    scaled = np.sqrt(g.model.var / g._mode_no) * raw
    t2 = time.perf_counter()
    print(f"summate_structured {1e3*(t1-t0):.1f} ms, host sqrt(var/N)* pass {1e3*(t2-t1):.1f} ms")
