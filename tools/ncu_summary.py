#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.txt: per-launch key
metrics of the `--set full` capture plus the hottest source lines by stall samples."""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# ncu summary of {rep}", ""]
for r in rows[2:]:
    lines.append(f"== launch id {r[hdr.index('ID')]}: {r[hdr.index('Kernel Name')][:110]}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"  {k:75s} {r[i]:>16s} {units[i]}")
    stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled_" in h and h.endswith("_per_issue_active.ratio")]
    lines.append("  stall reasons (warps stalled per issue-active cycle):")
    for v, h in sorted(stalls, reverse=True)[:8]:
        lines.append(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:60]))
