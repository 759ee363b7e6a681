"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/: one row per profiled launch."""
import csv, subprocess, sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
        for r in rows[2:]:
            name = r[idx[0]].split("(")[0]
            w.writerow([name] + [r[i] for i in idx[1:]])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
