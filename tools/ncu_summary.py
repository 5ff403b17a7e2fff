"""Compact summary of an .ncu-rep (raw page): python tools/ncu_summary.py <report> [substring ...]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct", "launch__registers_per_thread", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum ",
        "sm__pipe_tensor_cycles_active.avg.pct", "sm__inst_executed_pipe_xu.avg.pct", "sm__pipe_xu_cycles_active",
        "sm__inst_executed_pipe_fma.avg.pct", "sm__pipe_fma_cycles_active.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_uniform", "smsp__average_warp",
        "smsp__average_warps_issue_stalled", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct",
        "lts__t_bytes.sum ", "sm__inst_executed_pipe_tmem", "smsp__inst_executed_op_", "sm__cycles_elapsed.avg ",
        "smsp__cycles_active.avg ", "sm__cycles_active.avg "]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("==", name[:100])
        for h, u, v in zip(hdr, units, vals):
            if any(k.strip() in h for k in KEYS + extra):
                print("  %-90s %s %s" % (h, v, u))


if __name__ == "__main__":
    main()
