"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + opcode histogram + top stall lines.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
]  # fmt: skip


def run(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, units = rows[0], rows[1]
print(f"# {rep}\n", file=out)
for r in rows[2:]:
    print(f"## {r[hdr.index('Kernel Name')]}  (launch id {r[0]})\n", file=out)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"- `{k}` = {r[i]} {units[i]}", file=out)
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.05:
                print(f"- `{h}` = {v:.3f}", file=out)
    print(file=out)

rows = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "sass"))))
hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hidx:
    h = rows[hidx[0]]
    body = rows[hidx[0] + 1 : (hidx[1] - 1 if len(hidx) > 1 else len(rows))]
    ie, so, ss = h.index("Instructions Executed"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
    byop, stall, tot, tots = collections.Counter(), collections.Counter(), 0, 0
    lines = []
    for r in body:
        try:
            n = int(r[ie])
        except (ValueError, IndexError):
            continue
        m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[so].strip())
        op = m.group(2).split(".")[0] if m else r[so][:10]
        byop[op] += n
        tot += n
        try:
            s = int(r[ss])
        except ValueError:
            s = 0
        stall[op] += s
        tots += s
        lines.append((s, n, r[so].strip()))
    print(f"## SASS: {len(body)} instructions, {tot} warp-instructions executed, {tots} stall samples\n", file=out)
    print("| opcode | warp-instr | % | stall samples |\n|---|---|---|---|", file=out)
    for op, n in byop.most_common(28):
        print(f"| {op} | {n} | {100 * n / tot:.1f} | {stall[op]} |", file=out)
    print("\n### top stall-sample instructions\n", file=out)
    for s, n, src in sorted(lines, reverse=True)[:25]:
        print(f"- {s} samples, executed {n}: `{src}`", file=out)
