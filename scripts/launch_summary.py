"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total, per-kernel count and time."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for r in rows[1:]:
    v, u = float(r[vi].replace(",", "")), r[ui]
    us = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
    seq.append((r[ki][:70], us))
agg = collections.OrderedDict()
for k, u in seq:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += u
print("launches", len(seq), "total us", round(sum(u for _, u in seq), 1))
for k, (c, u) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{u:10.1f} us  x{c:4d}  {k}")
