#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
]
STALLS = "smsp__average_warps_issue_stalled_"


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    lines = [f"ncu summary of {rep}", ""]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        lines.append(f"== kernel: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for k in KEYS:
            if k in d and d[k] != "":
                lines.append(f"  {k:82s} {d[k]:>18s} {u.get(k, '')}")
        lines.append("  warp stall reasons (stalled warps per issue-active cycle):")
        st = sorted(((float(v), k) for k, v in d.items() if k.startswith(STALLS) and k.endswith("per_issue_active.ratio") and v),
                    reverse=True)
        for v, k in st[:8]:
            lines.append(f"    {k[len(STALLS):].replace('_per_issue_active.ratio', ''):28s} {v:8.3f}")
        try:
            dr = float(d["dram__bytes_read.sum"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_read.sum"]]
            dw = float(d["dram__bytes_write.sum"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_write.sum"]]
            lines.append(f"  DRAM traffic per launch: {dr + dw:.4g} bytes (read {dr:.4g}, write {dw:.4g})")
        except Exception:
            pass
        lines.append("")
    src = page(rep, "source")
    # the source page holds one block per profiled launch: ["Kernel Name", name], header row, data rows
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1] if len(r) > 1 else "?", header=None, rows=[])
            blocks.append(cur)
        elif cur is not None and cur["header"] is None:
            cur["header"] = r
        elif cur is not None and len(r) == len(cur["header"]):
            cur["rows"].append(r)
    seen = set()
    for b in blocks:
        if b["name"] in seen or not b["header"]:
            continue          # one block per distinct kernel is enough
        seen.add(b["name"])
        h, data = b["header"], b["rows"]
        idx = {n: i for i, n in enumerate(h)}
        ops = Counter()
        for r in data:
            s = r[idx["Source"]].split()
            if not s:
                continue
            op = s[1] if s[0].startswith("@") and len(s) > 1 else s[0]
            try:
                ops[op.split(".")[0]] += int(float(r[idx["Instructions Executed"]] or 0))
            except ValueError:
                pass
        tot = sum(ops.values())
        lines.append(f"[{b['name']}] executed warp-instructions by opcode (source page):")
        for op, c in ops.most_common(14):
            lines.append(f"  {op:12s} {c:14d}  {100.0 * c / max(tot, 1):5.1f}%")
        lines.append("")
        lines.append(f"[{b['name']}] top sampled instructions (stall samples):")

        def samples(r):
            try:
                return float(r[idx["# Samples"]] or 0)
            except ValueError:
                return 0.0

        for r in sorted(data, key=lambda r: -samples(r))[:12]:
            lines.append(f"  {r[idx['Source']][:60]:60s} samples {r[idx['# Samples']]:>7s}  long_sb {r[idx['stall_long_sb']]:>6s}"
                         f" math {r[idx['stall_math']]:>6s} wait {r[idx['stall_wait']]:>6s} short_sb {r[idx['stall_short_sb']]:>6s}")
        lines.append("")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
