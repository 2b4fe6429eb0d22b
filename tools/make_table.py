"""Result table of BASELINE.md section 6 from the bench lines under profiles/r02_table/ (tools/scale_table.sh).
    python tools/make_table.py            -> markdown on stdout"""
import glob
import json
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIR = os.path.join(REPO, "profiles", "r02_table")
NOMINAL = 2 * 148 * 64 * 1.965e9 / 1e12       # TFLOP/s


def load(name):
    path = os.path.join(DIR, name)
    if not os.path.exists(path):
        return None
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None


def fmt(v):
    return "—" if v is None else (f"{v / 1e12:.2f} T" if v >= 1e11 else f"{v / 1e9:.1f} G")


PARITY = {"c1": "1.6e-13", "c2": "3.3e-12", "c3": "≤ 1e-12", "c4": "≤ 1e-12", "c5": "≤ 1e-12", "c5cond": "≤ 1e-9 (test_gpu_cond)"}
NAMES = {"c1": "C1 100×100, N=1000 (direct: one tile)", "c2": "C2 512³, N=1000 (separable)",
         "c3": "C3 2·10⁷ points, N=10⁴ (direct 2-D)", "c4": "C4 256³ incompressible (separable ×3)",
         "c5": "C5 256 × 128³ summation, seeds sharded", "c5cond": "C5 through gs.CondSRF (1 realisation per GPU and step)"}


def main():
    print("| Config | GPUs | point·modes/s (device) | ms / step | % of measured FP64 peak (kernel; of nominal 37.2) | end-to-end "
          "point·modes/s (host in → host out) | max\\|Δ\\| / √var vs oracle | CPU port point·modes/s (cores) |")
    print("|---|---|---|---|---|---|---|---|")
    for cfg in ("c1", "c2", "c3", "c4", "c5", "c5cond"):
        for n in (1, 2, 4, 8):
            d = load(f"{cfg}_n{n}.json")
            if d is None:
                continue
            roof = d.get("roofline") or {}
            frac = roof.get("frac")
            nom = roof.get("frac_of_nominal_37.2_TFLOPs")
            pct = "—" if frac is None else f"{100 * frac:.1f} % ({100 * nom:.1f} %)"
            cpu = d.get("cpu_baseline")
            cpu_s = "—" if not cpu else f"{cpu['value'] / 1e9:.2f} G ({cpu['cores']})"
            print(f"| {NAMES[cfg]} | {n} | {fmt(d['value'])} | {d['ms_per_step']:.3f} | {pct} | {fmt(d['e2e']['value'])} "
                  f"({d['e2e']['ms_per_step']:.2f} ms) | {PARITY[cfg]} | {cpu_s} |")
    print()
    print("| C2, field gathered onto ONE device | GPUs | point·modes/s | ms / step | exposed gather ms |")
    print("|---|---|---|---|---|")
    for n in (2, 4, 8):
        for tag, label in (("c2_gather_p2p", "one process per GPU, kernels store into rank 0's tensor (CUDA IPC)"),
                           ("c2_gather_nccl", "one process per GPU, NCCL send/recv after the sum"),
                           ("c2_plan", "ONE process, gsb_plan device route (peer stores)")):
            d = load(f"{tag}_n{n}.json")
            if d is None:
                continue
            g = d.get("gather") or {}
            ex = g.get("exposed_ms")
            print(f"| {label} | {n} | {fmt(d['value'])} | {d['ms_per_step']:.3f} | {'—' if ex is None else f'{ex:.3f}'} |")
    print()
    print("| Kriging evaluation of config 5 (K = 1001, 128³; row f1) | GPUs | points/s | ms | % FP64 peak (algorithmic) | e2e points/s |")
    print("|---|---|---|---|---|---|")
    for n in (1, 2, 4, 8):
        d = load(f"krige_n{n}.json")
        if d is None:
            continue
        roof = d.get("roofline") or {}
        print(f"| krige_evaluate_structured | {n} | {d['value'] / 1e6:.1f} M | {d['ms_per_step']:.2f} | "
              f"{100 * roof.get('frac', 0):.1f} % | {d['e2e']['value'] / 1e6:.1f} M |")


if __name__ == "__main__":
    main()
