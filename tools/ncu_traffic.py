"""Per-kernel means over the launches of an ncu report -> profiles/r2_traffic.json (feeds bench.py's
roofline block: DRAM / L2 / local-memory bytes per launch, lanes per instruction, occupancy, issue
utilisation of the closest-hit and any-hit traversal kernels, launch-weighted over every launch the
report holds -- i.e. over all bounces of the captured batches, not just the first).
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
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}
WANT = {
    "ms": "gpu__time_duration.sum",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "l2_bytes": "lts__t_bytes.sum",
    "local_ld_inst": "smsp__sass_inst_executed_op_local_ld.sum",
    "local_st_inst": "smsp__sass_inst_executed_op_local_st.sum",
    "local_ld_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "local_st_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "shared_inst": "smsp__sass_inst_executed_op_shared.sum",
    "warp_inst": "smsp__inst_executed.sum",
    "lanes_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "dram_pct_of_peak": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "registers": "launch__registers_per_thread",
    "l1_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
}
acc = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].replace("(bool)", "")
    kind = "extend" if "k_trace<0" in name else "shadow" if "k_trace<1" in name else name.split("(")[0].split("::")[-1][:40]
    rec = {}
    for key, metric in WANT.items():
        if metric in hdr:
            i = hdr.index(metric)
            try:
                rec[key] = float(r[i].replace(",", "")) * mult.get(units[i], 1.0)
            except ValueError:
                pass
    acc.setdefault(kind, []).append(rec)
res = {}
for kind, recs in acc.items():
    n = len(recs)
    tot_ms = sum(x.get("ms", 0.0) for x in recs)
    m = {"launches": n}
    for key in WANT:
        vals = [x[key] for x in recs if key in x]
        if not vals:
            continue
        if key in ("lanes_per_inst", "occupancy_pct", "issue_active_pct", "l1_hit_pct", "l2_hit_pct", "dram_pct_of_peak", "l1_pipe_pct"):
            w = [x.get("ms", 1.0) for x in recs if key in x]  # time-weighted
            m[key] = sum(v * t for v, t in zip(vals, w)) / max(sum(w), 1e-30)
        elif key == "registers":
            m[key] = vals[0]
        else:
            m[key + "_per_launch"] = sum(vals) / n
    if "local_ld_sectors_per_launch" in m:
        m["local_mem_bytes_per_launch"] = 32.0 * (m["local_ld_sectors_per_launch"] + m.get("local_st_sectors_per_launch", 0.0))
    m["dram_bytes_per_launch"] = m.get("dram_bytes_read_per_launch", 0.0) + m.get("dram_bytes_write_per_launch", 0.0)
    res[kind] = m
    print(kind, json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in m.items()}))
if "--store" in sys.argv:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[workload] = dict(res, source=os.path.basename(rep))
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print("stored", path)
