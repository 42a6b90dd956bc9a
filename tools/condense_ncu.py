"""Key metrics of an `ncu --page raw --csv` export (one kernel launch) as a three-column CSV: what profiles/ keeps of
a full capture (the .ncu-rep files stay on the GPU box).

    python tools/condense_ncu.py raw.csv > profiles/r2_ncu_full_<name>.csv
"""
import csv
import sys

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))
    out = csv.writer(sys.stdout)
    out.writerow(["metric", "unit", "value"])
    out.writerow(["Kernel Name", "", d["Kernel Name"][1]])
    for k in KEYS:
        if k in d:
            out.writerow([k, d[k][0], d[k][1]])
    stalls = [(h, d[h]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
              and "not_issued" not in h and d[h][1] not in ("", "n/a")]
    for h, (u, v) in sorted(stalls, key=lambda x: -float(x[1][1])):
        out.writerow([h, u, v])


if __name__ == "__main__":
    main()
