"""Reads an ncu --set full report and prints / stores per-launch DRAM traffic of k_trace.
usage: python tools/ncu_traffic.py <report.ncu-rep> <workload> [--store]"""
import csv
import json
import os
import subprocess
import sys

rep, workload = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    def val(metric):
        i = hdr.index(metric)
        return float(r[i]) * mult.get(units[i], 1.0)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    dur = float(r[hdr.index("gpu__time_duration.sum")])
    kind = "extend" if "k_trace<0" in name.replace("(bool)", "") else ("shadow" if "k_trace<1" in name.replace("(bool)", "") else name[:40])
    print(f"{kind}: dram read {rd / 1e6:.1f} MB write {wr / 1e6:.1f} MB duration {dur} {units[hdr.index('gpu__time_duration.sum')]}")
    res.setdefault(kind, []).append(rd + wr)
if "--store" in sys.argv:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r1_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[workload] = {"extend_dram_bytes_per_launch": sum(res.get("extend", [0])) / max(1, len(res.get("extend", []))),
                      "shadow_dram_bytes_per_launch": sum(res.get("shadow", [0])) / max(1, len(res.get("shadow", []))),
                      "source": os.path.basename(rep)}
    json.dump(data, open(path, "w"), indent=1)
    print("stored", path)
