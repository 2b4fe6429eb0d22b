"""Opcode digest of the shipped library (cuobjdump -sass): per kernel, how many DMMA / DFMA / DMUL / DADD, bulk-copy
(UBLKCP), mbarrier (SYNCS), shared / global load-store and MUFU instructions the SASS holds, plus the architectures
present.  Static counts (not executed counts): evidence that the hot loops are fp64 tensor / FP64-pipe code fed by
bulk async copies.    python tools/sass_digest.py > profiles/rNN_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "gstools_b200", "libgsb200.so")
KEYS = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "MUFU", "UTCMMA", "UTMALDG", "LDTM",
        "SHFL", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    per = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            per[name]["total"] += 1
            for k in KEYS:
                if op.split(".")[0] == k or op.startswith(k + "."):
                    per[name][k] += 1
    print(f"library: gstools_b200/libgsb200.so   architectures in the fatbin: {', '.join(archs)}")
    print(f"{'kernel':78s} {'total':>7s} " + " ".join(f"{k:>7s}" for k in KEYS))
    tot = collections.Counter()
    for name, c in per.items():
        print(f"{name[:78]:78s} {c['total']:7d} " + " ".join(f"{c[k]:7d}" for k in KEYS))
        tot.update(c)
    print(f"{'ALL KERNELS':78s} {tot['total']:7d} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
    print("\nNotes: fp64 has no tcgen05 kind (no UTCMMA / LDTM expected): Blackwell's fp64 tensor path is mma.sync m8n8k4 ="
          " DMMA.8x8x4.\nOperands are pre-tiled in HBM, so 1-D bulk copies (cp.async.bulk = UBLKCP, completion on mbarriers"
          " = SYNCS) replace tensor-map TMA (UTMALDG).\nMUFU in the contraction kernels is the reciprocal seed of the integer"
          " divisions that decode a tile index (once per tile, not per stage); the stage loops hold DMMA / LDS / SYNCS only."
          "\nStatic counts: the 4096 DADD of sk_contract_kernel are the unrolled, predicated epilogue terms (up to 8 per"
          " output value), not loop work.")


if __name__ == "__main__":
    sys.exit(main())
