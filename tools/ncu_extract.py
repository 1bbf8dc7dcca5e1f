"""Extracts the columns DESIGN.md / profiles/ quote from an ncu report: python tools/ncu_extract.py in.ncu-rep out.csv"""
import csv
import io
import subprocess
import sys

COLS = [
    "ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = [head.index(c) for c in COLS if c in head]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in body:
            w.writerow([r[i] for i in idx])
    for r in body:
        print(" | ".join(r[i][:40] for i in idx))


if __name__ == "__main__":
    main()
