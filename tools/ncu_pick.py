"""Picks the launch to capture from an ncu launch list (csv of gpu__time_duration.sum per launch).
usage: python tools/ncu_pick.py list.csv <kernel substring> <nth heavy launch to pick> [min_us]
prints the 0-based index (among the listed launches) of the nth launch of that kernel lasting >= min_us"""
import csv
import sys

path, sub, nth = sys.argv[1], sys.argv[2], int(sys.argv[3])
min_us = float(sys.argv[4]) if len(sys.argv) > 4 else 500.0
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ik, iv, iu, iid = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
seen = 0
for r in rows[hdr + 1:]:
    if len(r) != len(h):
        continue
    v = float(r[iv].replace(",", ""))
    us = v / 1000.0 if r[iu].startswith("n") else (v * 1000.0 if r[iu].startswith("m") else v)
    if sub in r[ik] and us >= min_us:
        seen += 1
        if seen == nth:
            print(int(r[iid]))
            sys.exit(0)
sys.exit("no such launch")
