"""Static SASS attribution of one kernel by source line (no GPU needed): which inlined helpers are
repeated at many call sites.  usage: python tools/sass_static.py <lib.so> <kernel-substring> [top]"""
import collections, glob, os, re, subprocess, sys, tempfile
lib, want = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for cubin in glob.glob(tmp + "/*.cubin"):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next((i for i, l in enumerate(dis) if l.startswith(".text.") and want in l), None)
    if start is None: continue
    end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith(".text.") or dis[i].startswith(".section")), len(dis))
    fre = re.compile(r'//## File "([^"]+)", line (\d+)'); ire = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);')
    cur = ("?", 0); cnt = collections.Counter(); sites = collections.Counter(); prev = None
    for l in dis[start:end]:
        m = fre.search(l)
        if m:
            if "inlined at" in l: continue
            cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        if ire.match(l):
            cnt[cur] += 1
            if cur != prev: sites[cur] += 1
            prev = cur
    total = sum(cnt.values()); print("kernel", dis[start][:90], "instructions", total, "=", total * 16 // 1024, "KB")
    byfile = collections.Counter()
    for (f, _), n in cnt.items(): byfile[f] += n
    print(byfile.most_common(8))
    for (f, ln), n in cnt.most_common(top):
        path = [p for p in glob.glob(root + "/pbrs_b200/csrc/*") if p.endswith("/" + f)]
        text = open(path[0]).read().splitlines()[ln - 1].strip()[:84] if path else ""
        print(f"{n:5d} in {sites[(f, ln)]:3d} runs  {f}:{ln}  {text}")
