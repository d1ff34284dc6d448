"""Compact text summary of an ncu report for profiles/: per kernel the headline metrics and the SASS instructions with the most
stall samples.  usage: ncu_top.py report.ncu-rep [n_top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__cluster_dim_x",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = raw[0], raw[1]
for r in raw[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")][:100], "| launch id", r[0])
    for k in KEYS:
        if k in hdr:
            print(f"   {k:95s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout.splitlines()))
cur, h, rows = None, None, []


def flush():
    if not rows:
        return
    ia, isamp = h.index("Instructions Executed"), h.index("# Samples")
    ts = sum(int(x[isamp]) for x in rows) or 1
    ti = sum(int(x[ia]) for x in rows) or 1
    ops = {}
    for x in rows:
        t = x[1].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + int(x[ia])
    print("== source:", cur, "| executed warp instructions", ti, "| samples", ts)
    print("   opcode mix (executed):", ", ".join(f"{k} {100 * v / ti:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:12]))
    base = int(rows[0][0], 16)
    for x in sorted(sorted(rows, key=lambda x: -int(x[isamp]))[:ntop], key=lambda x: int(x[0], 16)):
        print(f"   {int(x[0], 16) - base:06x}  {100 * int(x[isamp]) / ts:5.1f}% samples  {x[1].strip()[:90]}")


for r in src:
    if r and r[0] == "Kernel Name":
        flush()
        cur, rows = r[1][:100], []
    elif r and r[0] == "Address":
        h = r
    elif h and len(r) > 5:
        rows.append(r)
flush()
