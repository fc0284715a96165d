"""Compact per-launch summary of `ncu --page raw --csv` exports (the .ncu-rep files are too large to commit).
    python scripts/ncu_summary.py out.csv raw1.csv raw2.csv ...
ncu picks the unit of every column PER FILE (us / ms, Mbyte / Gbyte ...): values are converted to fixed units here
(time in us, bytes in MB = 1e6 bytes) and the unit is part of the column name, so rows of different captures compare."""
import csv, sys
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle"]
SCALE = {"ns": ("us", 1e-3), "us": ("us", 1.0), "ms": ("us", 1e3), "s": ("us", 1e6),
         "byte": ("MB", 1e-6), "Kbyte": ("MB", 1e-3), "Mbyte": ("MB", 1.0), "Gbyte": ("MB", 1e3)}


def main(argv):
    out = csv.writer(open(argv[1], "w", newline=""))
    header = None
    for path in argv[2:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        idx = [hdr.index(k) if k in hdr else -1 for k in KEYS]
        conv = [SCALE.get(units[i], (units[i], None)) if i >= 0 else ("", None) for i in idx]
        if header is None:
            header = ["source"] + [k + (" [%s]" % c[0] if c[0] else "") for k, c in zip(KEYS, conv)]
            out.writerow(header)
        for r in rows[2:]:
            vals = []
            for i, c in zip(idx, conv):
                v = r[i][:110] if i >= 0 else ""
                if c[1] is not None and v:
                    v = "%.6g" % (float(v.replace(",", "")) * c[1])
                vals.append(v)
            out.writerow([path.split("/")[-1]] + vals)


if __name__ == "__main__":
    main(sys.argv)
