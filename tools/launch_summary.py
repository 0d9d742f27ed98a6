"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python tools/launch_summary.py <csv> [first_id last_id]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
agg = collections.OrderedDict()
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            i = int(d["ID"])
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if lo <= i <= hi:
            agg.setdefault(d["Kernel Name"][:70], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} n={len(v):5d} total={sum(v)/1e6:9.3f} ms  mean={sum(v)/len(v)/1e3:8.1f} us  min={min(v)/1e3:8.1f} max={max(v)/1e3:8.1f}")
print(f"all: {tot/1e6:.3f} ms in {sum(len(v) for v in agg.values())} launches")
