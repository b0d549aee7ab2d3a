#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU needed): key raw metrics per captured launch and the hottest
source lines by warp-stall samples.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/x.md]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_imc_miss_per_warp_active.pct",
]


def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units = raw[0], raw[1]
idx = {n: i for i, n in enumerate(hdr)}
print(f"# ncu summary of `{rep}`\n")
for r in raw[2:]:
    print(f"## {r[idx['Kernel Name']]}  (grid {r[idx.get('launch__grid_size', 0)]}, block {r[idx.get('launch__block_size', 0)]})\n")
    print("| metric | value | unit |\n|---|---|---|")
    for m in METRICS:
        if m in idx:
            print(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |")
    print()

src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv"]))))
if src and src[0] and src[0][0] == "Kernel Name":
    src = src[1:]
if src:
    h = src[0]
    col = {n: i for i, n in enumerate(h)}
    samp = next((c for c in h if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling" in c), None)
    srccol = next((c for c in h if c in ("Source",)), None)
    if samp and srccol:
        rows = []
        for r in src[1:]:
            try:
                rows.append((float(r[col[samp]] or 0), r[col[srccol]].strip(), r))
            except (ValueError, IndexError):
                pass
        tot = sum(x[0] for x in rows) or 1
        rows.sort(key=lambda x: -x[0])
        print(f"## hottest lines by `{samp}` (total {tot:.0f})\n")
        print("| % | samples | source |\n|---|---|---|")
        for s, text, _ in rows[:30]:
            print(f"| {100 * s / tot:.1f} | {s:.0f} | `{text[:110]}` |")
        if "Instructions Executed" in col:
            ie = col["Instructions Executed"]
            tot_i = sum(float(r[ie] or 0) for _, _, r in rows) or 1
            by = sorted(rows, key=lambda x: -float(x[2][ie] or 0))
            print(f"\n## most executed SASS instructions (total warp instructions {tot_i:.0f})\n")
            print("| % | executed | SASS |\n|---|---|---|")
            for _, text, r in by[:25]:
                print(f"| {100 * float(r[ie]) / tot_i:.2f} | {float(r[ie]):.0f} | `{text[:100]}` |")
