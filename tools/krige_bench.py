#!/usr/bin/env python
"""Time the kriging evaluation kernel: K conditioning rows, n points, device-resident krig_vecs."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gstools_b200 as gsb
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1001
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128 ** 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
g = torch.Generator(device="cuda").manual_seed(1)
mat = torch.randn(K, K, dtype=torch.float64, device="cuda", generator=g)
kv = torch.rand(K, n, dtype=torch.float64, device="cuda", generator=g)
cond = torch.randn(K, dtype=torch.float64, device="cuda", generator=g)
peak = gsb.measure_fp64_peak(0, 0, 0.3)
for want_var in (True, False):
    fn = gsb.calc_field_krige_and_variance if want_var else gsb.calc_field_krige
    for _ in range(2):
        fn(mat, kv, cond)
    torch.cuda.synchronize()
    gsb.set_option("time_kernels", 1); gsb.kernel_times()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn(mat, kv, cond)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    kms, kn = gsb.kernel_times(); gsb.set_option("time_kernels", 0)
    full = 1.0 * K * K * n            # FMAs of the reference's loop nest
    tri = 0.5 * K * (K + 1) * n + K * n
    if want_var:
        print(f"calc_field_krige_and_variance K={K} n={n}: {ms:.2f} ms/call (kernel {kms/max(kn,1):.2f} ms) | "
              f"reference-loop FMAs {full/ms/1e9:.2f} TFMA/s-equivalent, executed (triangular) {tri/ms/1e9:.2f} TFMA/s "
              f"= {100*tri/ms/1e-3/peak:.1f} % of measured DFMA peak {peak/1e12:.2f} TFMA/s")
    else:
        print(f"calc_field_krige K={K} n={n}: {ms:.2f} ms/call, {8.0*K*n/ms/1e6:.0f} GB/s of krig_vecs")
# the whole evaluation loop with device-generated right-hand sides (BASELINE.json configs[4] kriging step:
# 1000 conditioning points, ordinary kriging, Exponential(dim=3, var=1, len_scale=10), 128^3 mesh)
import time
rs = np.random.RandomState(20170519)
C = K - 1
edge = round(n ** (1 / 3))
if edge ** 3 == n:
    cond_pos = rs.uniform(0, edge - 1, (3, C))
    axes = [np.arange(float(edge))] * 3
    matn, condn = mat.cpu().numpy() / K, cond.cpu().numpy()
    spec = dict(kind="Exponential", var=1.0, len_rescaled=10.0)
    for i in range(3):
        t0 = time.perf_counter()
        gsb.set_option("time_kernels", 1); gsb.kernel_times()
        f, e = gsb.krige_evaluate(spec, matn, condn, cond_pos, axes=axes)
        t = time.perf_counter() - t0
        kms, kn = gsb.kernel_times(); gsb.set_option("time_kernels", 0)
        print(f"krige_evaluate (device right-hand sides) K={K} mesh {edge}^3: {1e3*t:.1f} ms wall, "
              f"contraction kernels {kms:.2f} ms in {kn} launches")
# the reference's loop nest on the host cores (C/OpenMP oracle, test infrastructure) on a bounded sample
import oracle
ns = 16384
kvs = np.random.RandomState(0).uniform(0, 1, (K, ns))
mats, conds = mat.cpu().numpy(), cond.cpu().numpy()
oracle.calc_field_krige_and_variance(mats, kvs[:, :256], conds)
t0 = time.perf_counter()
oracle.calc_field_krige_and_variance(mats, kvs, conds)
t = time.perf_counter() - t0
print(f"CPU oracle calc_field_krige_and_variance K={K} on {ns} points, {oracle.max_threads()} threads: {t:.2f} s "
      f"-> {t / ns * n:.1f} s extrapolated to n={n} ({1.0 * K * K * ns / t / 1e9:.2f} GFMA/s)")
