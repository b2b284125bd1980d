#!/usr/bin/env python
"""dev helper: print the metrics DESIGN/profiles quote from an .ncu-rep as a markdown table (one column per launch).

    python tools/ncu_table.py gpurun_out/r02_gemm.ncu-rep [--limit 8]
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("time", "gpu__time_duration.sum"),
    ("dram read", "dram__bytes_read.sum"),
    ("dram write", "dram__bytes_write.sum"),
    ("dram % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("XU pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("FMA pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("ALU pipe %", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("SM clock", "sm__cycles_elapsed.avg.per_second"),
    ("regs/thread", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("occupancy limit regs", "launch__occupancy_limit_registers"),
    ("occupancy limit smem", "launch__occupancy_limit_shared_mem"),
    ("local load sectors", "lts__t_sectors_srcunit_tex_aperture_device_op_read_lookup_miss.sum"),
]


def main():
    rep = sys.argv[1]
    limit = int(sys.argv[sys.argv.index("--limit") + 1]) if "--limit" in sys.argv else 8
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    data = rows[2:2 + limit]
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col["Kernel Name"]][:60] for r in data]
    print("| metric | " + " | ".join(f"#{i} {n}" for i, n in enumerate(names)) + " |")
    print("|---|" + "---:|" * len(names))
    for label, key in METRICS:
        if key not in col:
            continue
        i = col[key]
        print(f"| {label} (`{key}`) | " + " | ".join(f"{r[i]} {units[i]}" for r in data) + " |")


if __name__ == "__main__":
    main()
