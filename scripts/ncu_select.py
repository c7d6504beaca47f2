"""Condense an .ncu-rep (ncu --set full) into the metrics the profile notes quote: one column per captured launch.
usage: python scripts/ncu_select.py <report.ncu-rep> <out.csv> "<comment line>" """
import csv
import subprocess
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
]


def main():
    rep, out, comment = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units, launches = rows[0], rows[1], rows[2:]
    stall = [h for h in head if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    with open(out, "w", newline="") as f:
        if comment:
            f.write("# " + comment + "\n")
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch %d" % i for i in range(len(launches))])
        for m in KEEP + stall:
            if m in head:
                j = head.index(m)
                w.writerow([m, units[j]] + [r[j] for r in launches])


if __name__ == "__main__":
    main()
