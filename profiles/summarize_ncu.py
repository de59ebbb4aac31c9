#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of per-launch numbers DESIGN.md / bench.py cite.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe % (F2F cvt, MUFU)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of `{path}` (ncu --set full --clock-control none; per launch, cold cache, serialised)\n")
    for r in rows[2:]:
        print(f"## {r[hdr.index('Kernel Name')]}\n")
        print("| metric | value |\n|---|---|")
        for m, label in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {label} (`{m}`) | {r[i]} {units[i]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if "warps_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                try:
                    stalls.append((float(r[i]), h.split("warps_issue_stalled_")[1].split("_per_warp")[0]))
                except ValueError:
                    pass
        top = ", ".join(f"{n} {v:.1f}%" for v, n in sorted(stalls, reverse=True)[:5])
        print(f"| top stall reasons (per active warp) | {top} |\n")


if __name__ == "__main__":
    main(sys.argv[1])
