"""Compact per-kernel summary of an .ncu-rep (run where ncu is installed):  python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("time_us", "gpu__time_duration.sum"), ("dram_rd_GB", "dram__bytes_read.sum"), ("dram_wr_GB", "dram__bytes_write.sum"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2hit%", "lts__t_sector_hit_rate.pct"), ("warps", "sm__warps_active.avg.per_cycle_active"), ("regs", "launch__registers_per_thread"),
        ("inst", "smsp__inst_executed.sum"), ("cycles", "sm__cycles_elapsed.avg"), ("grid", "launch__grid_size"),
        ("st_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
        ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
        ("st_bar", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("st_short", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("st_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio")]
for d in data:
    name = d[hdr.index("Kernel Name")]
    name = name.split("(")[0][-70:]
    out = []
    for label, key in want:
        if key in hdr:
            i = hdr.index(key)
            try:
                v = float(d[i].replace(",", ""))
            except ValueError:
                continue
            u = units[i]
            if label == "time_us":
                v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
            if label.endswith("_GB"):
                v = {"byte": v / 1e9, "Kbyte": v / 1e6, "Mbyte": v / 1e3, "Gbyte": v}.get(u, v)
            out.append(f"{label}={v:.3g}")
    print(name, "|", " ".join(out))
