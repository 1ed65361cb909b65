#!/usr/bin/env python
"""Turn an .ncu-rep (read here, no GPU needed) into a short text summary under profiles/.

    python profiles/summarize.py gpurun_out/<name>.ncu-rep > profiles/<name>.summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    rec = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("kernel:", rec.get("Kernel Name"))
    for k in KEYS:
        if k in rec:
            print(f"  {k:72s} {rec[k]} {u[k]}")
    stalls = [(float(v), h) for h, v in rec.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v]
    print("  warp stall reasons (cycles per issued instruction):")
    for v, h in sorted(stalls, reverse=True)[:7]:
        name = h.split("issue_stalled_")[1].split("_per_issue")[0]
        print(f"    {name:28s} {v:.2f}")
    print()
