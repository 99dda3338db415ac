"""Prints selected metrics of every kernel in an .ncu-rep (ncu -i ... --page raw --csv piped in)."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = [("Kernel Name", "kernel", 70), ("gpu__time_duration.sum", "us", 9), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 7),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 7), ("dram__bytes_read.sum", "rd", 9), ("dram__bytes_write.sum", "wr", 9),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 7), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 7),
        ("launch__registers_per_thread", "regs", 5), ("sm__cycles_elapsed.avg", "cycles", 10)]
idx = [(hdr.index(w) if w in hdr else -1, n, wd) for w, n, wd in want]
units = rows[1]
print(" ".join(n.ljust(wd) for _, n, wd in idx))
for r in rows[2:]:
    print(" ".join(((r[i] + (units[i] if n in ("rd", "wr") else ""))[:wd] if i >= 0 else "-").ljust(wd) for i, n, wd in idx))
