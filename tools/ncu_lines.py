"""Top source lines of a kernel in an .ncu-rep captured with --import-source on (-lineinfo build).
usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top-n] [launch-index]"""
import csv, subprocess, sys
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
sel = ["--launch-skip", sys.argv[4], "--launch-count", "1"] if len(sys.argv) > 4 else []
out = subprocess.run(["ncu", "-i", rep] + sel + ["--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = cur_fn = None; hdr = None; agg = {}; tot = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr) or not r[0] or want not in (cur_fn or ""): continue
    try:
        s = float(r[hdr.index("# Samples")]); n = float(r[hdr.index("Instructions Executed")]); t = float(r[hdr.index("Thread Instructions Executed")])
    except ValueError:
        continue
    key = (cur_fn[:60], cur_file, int(r[0]), r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0, 0]); a[0] += s; a[1] += n; a[2] += t
for fn in sorted({k[0] for k in agg}):
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    ts = sum(v[0] for _, v in items); ti = sum(v[1] for _, v in items)
    print(f"== {fn}: samples {ts:.0f}, warp instructions {ti:.0f}, threads/inst {sum(v[2] for _, v in items) / max(ti, 1):.1f}")
    for k, v in sorted(items, key=lambda x: -x[1][0])[:top]:
        print(f"  {100 * v[0] / max(ts, 1):5.1f}% smp {100 * v[1] / max(ti, 1):5.1f}% inst thr {v[2] / max(v[1], 1):4.1f}  {k[1]}:{k[2]}  {k[3]}")
