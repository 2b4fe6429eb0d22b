"""Where does the end-to-end time of a multi-GPU host-route call go?  One process, a gsb_plan over the GPUs of
the box; the 512^3 field of config 2 (or --edge N).  Prints, per variant, the wall time of
summate_structured(numpy) -> numpy (pinned):
  one GPU, whole field / one GPU, 1/G of the field (the other GPUs idle) / all GPUs, each 1/G at the same time;
  each with the default growing pieces and with 1, 4, 16 equal pieces (option host_pieces)."""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench_configs as bc  # noqa: E402
import gstools_b200 as gsb  # noqa: E402
import torch  # noqa: E402


def timeit(fn, reps=5):
    fn()
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
        del out
    return best * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", type=int, default=512)
    args = ap.parse_args()
    cfg = bc.config2(args.edge)
    cov, z1, z2, axes = cfg["cov"], cfg["z1"], cfg["z2"], cfg["axes"]
    G = torch.cuda.device_count()
    gb = 8 * args.edge ** 3 / 1e9
    plan = gsb.Plan(list(range(G))) if G > 1 else None
    for pieces in (0, 1, 4, 16):
        gsb.set_option("host_pieces", pieces)
        t_all = timeit(lambda: gsb.summate_structured(cov, z1, z2, axes))
        part = [axes[0][: args.edge // G]] + axes[1:]
        t_part = timeit(lambda: gsb.summate_structured(cov, z1, z2, part))
        line = f"host_pieces={pieces:2d}  1 GPU whole {t_all:7.2f} ms ({gb / t_all * 1e3:5.1f} GB/s)   1 GPU 1/{G} {t_part:7.2f} ms"
        if plan is not None:
            t_plan = timeit(lambda: plan.summate_structured(cov, z1, z2, axes))
            line += f"   {G} GPUs {t_plan:7.2f} ms ({gb / t_plan * 1e3:5.1f} GB/s aggregate)"
        print(line, flush=True)
    gsb.set_option("host_pieces", 0)


if __name__ == "__main__":
    main()
