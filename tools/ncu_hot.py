"""Top stall-sample instructions per kernel from `ncu -i X.ncu-rep --page source --csv` (stdin).  Usage: ... | python tools/ncu_hot.py <kernel substring> [topN]"""
import csv, sys
pat = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = csv.reader(sys.stdin)
cur, hdr, data, done = None, None, [], False
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        if cur and pat in cur and data:
            break
        cur, hdr, data = r[1], None, []
        continue
    if r[0] == "Address":
        hdr = r
        continue
    if cur and pat in cur and hdr:
        data.append(r)
if not data:
    sys.exit("kernel not found")
si = hdr.index("# Samples")
ii = hdr.index("Instructions Executed")
tot = sum(int(d[si] or 0) for d in data)
print(cur[:120], "total samples", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda k: -int(data[k][si] or 0))[:top]
for k in sorted(order):
    d = data[k]
    print(f"{k:5d} {int(d[si]):7d} {100.0*int(d[si])/tot:5.1f}%  exec {d[ii]:>9}  {d[1].strip()[:90]}")
