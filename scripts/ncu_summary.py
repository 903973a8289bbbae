#!/usr/bin/env python
"""Condense one `ncu --set full` report of wg_flow_kernel into a small text summary for profiles/.

usage: scripts/ncu_summary.py <report.ncu-rep> <out.txt> [kernel-symbol-substring]
"""
import csv
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    sym = sys.argv[3] if len(sys.argv) > 3 else "wg_flow_kernel"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none summary of {os.path.basename(rep)} (first captured launch)"]
    r = rows[2]
    kn = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    if kn is not None:
        lines.append(f"kernel: {r[kn]}")
    for h, u, v in zip(hdr, units, r):
        if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            lines.append(f"{h} [{u}] = {v}")
    here = os.path.dirname(os.path.abspath(__file__))
    by = subprocess.run([sys.executable, os.path.join(here, "ncu_by_line.py"), rep, sym, "0.7"],
                        capture_output=True, text=True).stdout
    lines.append("")
    lines.append("# executed warp instructions / stall samples by CUDA source line (>= 0.7 %)")
    lines.append(by)
    with open(out, "w") as fh:
        fh.write("\n".join(lines))
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
