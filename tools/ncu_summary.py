#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full capture) into the metrics DESIGN.md / bench.py cite.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt
Runs `ncu -i <rep> --page raw --csv` (no GPU needed) and prints one block per captured launch."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"== {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}")
        for k in KEYS:
            if k in d:
                print(f"{k:75s} {d[k]:>18s} {u[k]}")
        stalls = sorted(((float(d[h]), h[len(STALL):-len('_per_issue_active.ratio')]) for h in hdr
                         if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and d[h]), reverse=True)
        print("stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:8]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
