"""Summarise an .ncu-rep (raw page + source page) into text: key metrics, stall reasons,
dynamic opcode mix, hottest source regions.  Usage: python tools/ncu_summary.py file.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # which captured launch
hdr, units, vals = rows[0], rows[1], rows[2 + which]
m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_register_spilling",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in m:
        print(f"{k:70s} {m[k]} {u.get(k, '')}")
print("-- stall reasons (warps per issue-active cycle) --")
for k in sorted(m):
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
        print(f"  {k.split('stalled_')[1].split('_per_issue')[0]:28s} {float(m[k]):.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot, samp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    f = r[ia].split()
    op = (f[1] if f[0].startswith("@") else f[0]).split(".")[0]
    try:
        tot[op] += int(r[ie]); samp[op] += int(r[isamp])
    except (ValueError, IndexError):
        continue
T, S = sum(tot.values()), sum(samp.values())
print(f"-- dynamic opcode mix: {T} warp instructions, {len(rows) - 2} static ({(len(rows) - 2) * 16} B) --")
for op, c in tot.most_common(22):
    print(f"  {op:10s} {100 * c / T:5.1f}% of issued   {100 * samp[op] / max(S, 1):5.1f}% of samples")
