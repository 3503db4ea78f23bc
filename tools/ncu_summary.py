"""Print the metrics this repo's profiles/ summaries quote from an `ncu --set full` report (one block per kernel launch):
    python tools/ncu_summary.py gpurun_out/foo.ncu-rep [kernel-name-substring]
Needs the `ncu` binary on PATH (it only reads the report; no GPU)."""
import csv
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "launch__cluster_dim_x", "gpu__time_duration.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
]


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if want and want not in r[col["Kernel Name"]]:
            continue
        for k in KEYS:
            if k in col:
                print(f"{k:86s} {r[col[k]]} {units[col[k]]}")
        print()


if __name__ == "__main__":
    main()
