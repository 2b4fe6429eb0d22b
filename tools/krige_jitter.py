"""Wall time of identical krige_evaluate calls (config 5: K = 1001, 128^3), host arrays in and out."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import gstools_b200 as gsb  # noqa: E402

w = bench.make_krige_workload()
ts = []
for _ in range(14):
    t0 = time.perf_counter()
    f, e = gsb.krige_evaluate(w["spec"], w["mat"], w["cond"], w["cond_pos"], axes=w["axes"])
    ts.append((time.perf_counter() - t0) * 1e3)
    del f, e
print("ms per call:", " ".join(f"{t:.1f}" for t in ts))
print(f"after warm-up: min {min(ts[3:]):.1f} median {np.median(ts[3:]):.1f} max {max(ts[3:]):.1f}")
