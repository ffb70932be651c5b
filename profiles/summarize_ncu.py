#!/usr/bin/env python
"""Turn an .ncu-rep (brought back in gpurun_out/) into the small text summary kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_cfg2.ncu-rep profiles/r01/ncu_cfg2_fused.md

Reads the report with `ncu -i ... --page raw --csv` (kernel-level metrics) and
`--page source --csv --print-source sass,cuda` (per-line instruction counts; needs -lineinfo).
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "mio_throttle", "wait", "math_pipe_throttle", "not_selected",
          "no_instruction", "lg_throttle", "branch_resolving", "dispatch_stall"]


def run(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main(rep, out):
    raw = list(csv.reader(io.StringIO(run("-i", rep, "--page", "raw", "--csv"))))
    hdr, units, data = raw[0], raw[1], raw[2:]
    lines = [f"# ncu summary of `{rep.split('/')[-1]}`", "", "`ncu --set full --clock-control none --import-source on` (replayed, cold-cache: compare shares, not absolutes).", ""]
    name_i = hdr.index("Kernel Name")
    for r in data:
        lines += [f"## {r[name_i]}", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"| {k} | {r[i]} | {units[i]} |")
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in hdr:
                lines.append(f"| stall {s} (warps per issue) | {r[hdr.index(k)]} | |")
        lines.append("")
    # per source line: instructions and warp-stall samples (the "sass,cuda" view lists every CUDA line with its totals)
    src = list(csv.reader(io.StringIO(run("-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"))))
    per_line, samples, text, why = collections.Counter(), collections.Counter(), {}, {}
    cur, h = None, None
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur, h = r[1].split("/")[-1], None
        elif r[0] == "Line No":
            h = r
        elif h and cur and r[0].isdigit() and r[2] == "-":
            try:
                n, smp = int(r[h.index("Instructions Executed")]), int(r[h.index("# Samples")])
            except (ValueError, IndexError):
                continue
            key = (cur, int(r[0]))
            per_line[key] += n
            samples[key] += smp
            text[key] = r[1].strip()
            st = {c[6:]: int(r[i]) for i, c in enumerate(h) if c.startswith("stall_") and "(" not in c and r[i].isdigit() and int(r[i])}
            why[key] = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    tot = sum(per_line.values()) or 1
    lines += ["## hottest source lines (share of executed warp instructions, all captured launches)", "", "| file:line | % | source |", "|---|---|---|"]
    for k, v in per_line.most_common(25):
        lines.append(f"| {k[0]}:{k[1]} | {100 * v / tot:.1f} | `{text[k][:110]}` |")
    tot_s = sum(samples.values()) or 1
    lines += ["", "## source lines by warp-stall samples (where the warps wait; top stall reasons of the line)", "",
              "| file:line | % of samples | top stalls | source |", "|---|---|---|---|"]
    for k, v in samples.most_common(25):
        lines.append(f"| {k[0]}:{k[1]} | {100 * v / tot_s:.1f} | {why.get(k, '')} | `{text[k][:90]}` |")
    # SASS opcode mix (the "sass" view lists every instruction once per kernel)
    sass = list(csv.reader(io.StringIO(run("-i", rep, "--page", "source", "--csv", "--print-source", "sass"))))
    ops, h = collections.Counter(), None
    for r in sass:
        if not r:
            continue
        if r[0] == "Address":
            h = r
        elif h and r[0].startswith("0x"):
            try:
                n = int(r[h.index("Instructions Executed")])
            except (ValueError, IndexError):
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1])
            if m:
                ops[m.group(2)] += n
    tot_ops = sum(ops.values()) or 1
    lines += ["", "## SASS opcode mix (share of executed warp instructions)", "", ", ".join(f"{k} {100 * v / tot_ops:.1f}%" for k, v in ops.most_common(30)), ""]
    open(out, "w").write("\n".join(lines))
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
