"""Timeline (option "trace") of the host route for the whole 512^3 field and for half of it on one GPU."""
import sys
import time

sys.path.insert(0, ".")
import bench_configs as bc  # noqa: E402
import gstools_b200 as gsb  # noqa: E402
import torch  # noqa: E402

cfg = bc.config2(512)
cov, z1, z2, axes = cfg["cov"], cfg["z1"], cfg["z2"], cfg["axes"]
for frac in (1, 2, 4, 8):
    part = [axes[0][: 512 // frac]] + axes[1:]
    for _ in range(3):
        out = gsb.summate_structured(cov, z1, z2, part)
        del out
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = gsb.summate_structured(cov, z1, z2, part)
    t = time.perf_counter() - t0
    print(f"--- 1/{frac} of the field: {t * 1e3:.2f} ms, pinned={torch.from_numpy(out).is_pinned()}", file=sys.stderr, flush=True)
    del out
    gsb.set_option("trace", 1)
    out = gsb.summate_structured(cov, z1, z2, part)
    gsb.set_option("trace", 0)
    del out
