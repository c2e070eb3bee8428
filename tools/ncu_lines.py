"""Per-source-line view of an ncu capture of one kernel: joins the SASS source
page (sampling, executed instructions) with nvdisasm's line table of the built
library.  python tools/ncu_lines.py gpurun_out/commit_v2.ncu-rep k_commit2 [top]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cranesched_b200", "csrc", "libcrane_sched.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
line_of = {}
cur, inside = None, False
for l in dis.splitlines():
    if l.startswith(".text."):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*)", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix["Address"]], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(lambda: defaultdict(int))
tot = 0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    ln = line_of.get(off, ("?", 0))
    s = int(r[ix["# Samples"]])
    tot += s
    a = agg[ln]
    a["samples"] += s
    a["inst"] += int(r[ix["Instructions Executed"]])
    for h in stalls:
        a[h] += int(r[ix[h]])
# NCU_LINES_WORK=1: rank by the samples that are NOT barrier waits — with one warp (or half the CTA)
# working while the rest waits at a barrier, these are the critical path
work = bool(os.environ.get("NCU_LINES_WORK"))
bar = [h for h in stalls if "barrier" in h]
if work:
    for a in agg.values():
        a["samples_all"] = a["samples"]
        a["samples"] = a["samples"] - sum(a[h] for h in bar)
    tot = sum(a["samples"] for a in agg.values())
src = {}
for (f, n), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    if f not in src:
        p = os.path.join(ROOT, "cranesched_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
    st = sorted(((a[h], h[6:]) for h in stalls), reverse=True)[:3]
    print("%5.1f%% %9d inst  %-16s:%-4d %-34s | %s" % (100.0 * a["samples"] / tot, a["inst"], f, n,
          " ".join("%s %d%%" % (h, 100 * v // max(a["samples"], 1)) for v, h in st), text))
print("total samples", tot)
if len(sys.argv) > 5:  # listing of a line range in source order: lo hi
    lo, hi = int(sys.argv[4]), int(sys.argv[5])
    print("---- lines %d..%d of commit_v2.cuh in source order" % (lo, hi))
    f = "commit_v2.cuh"
    lines = open(os.path.join(ROOT, "cranesched_b200", "csrc", f)).read().splitlines()
    for n in range(lo, hi + 1):
        a = agg.get((f, n))
        if not a:
            continue
        st = sorted(((a[h], h[6:]) for h in stalls), reverse=True)[:3]
        print("%5.2f%% %9d inst :%-4d %-40s | %s" % (100.0 * a["samples"] / tot, a["inst"], n,
              " ".join("%s %d%%" % (h, 100 * v // max(a["samples"], 1)) for v, h in st), lines[n - 1].strip()[:100]))
