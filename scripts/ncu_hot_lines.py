#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction samples with nvdisasm line info -> hottest source lines.
usage: ncu_hot_lines.py report.ncu-rep mangled_kernel_substr cubin_glob_substr [topn]   (NCU_K=<regex on the demangled name> when they differ; BY_INS=1 sorts by instructions)"""
import collections, csv, glob, os, re, subprocess, sys, tempfile
rep, kern, cub = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("HOT_LIB", os.path.join(root, "dlsc_gc_planner_b200", "libdlsc_b200.so"))], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in glob.glob(tmp + "/*.cubin") if cub in f][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
cur, inside, a2l = None, False, {}
for l in sass:
    if l.startswith(".text."):
        inside = kern in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m and inside:
        a2l[int(m.group(1), 16)] = (cur, m.group(2))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + os.environ.get("NCU_K", kern)], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[hi + 1:] if len(r) > iex and r[ia]]
base = int(data[0][ia], 16) if data[0][ia].startswith("0x") else int(data[0][ia])
S, I = collections.Counter(), collections.Counter()
stalls = collections.defaultdict(collections.Counter)
ts = ti = 0
for r in data:
    try:
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except ValueError:
        continue
    ln = a2l.get(a - base, (None, ""))[0]
    s, ie = int(r[isamp] or 0), int(r[iex] or 0)
    S[ln] += s; I[ln] += ie; ts += s; ti += ie
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v:
            stalls[ln][h] += v
print("total samples", ts, "warp-instructions", ti)
src_cache = {}
order = I.most_common(topn) if os.environ.get("BY_INS") else S.most_common(topn)
for ln, _ in order:
    s = S[ln]
    txt = ""
    if ln:
        for d in ("dlsc_gc_planner_b200/csrc",):
            p = os.path.join(root, d, ln[0])
            if os.path.exists(p):
                src_cache.setdefault(p, open(p).read().split("\n"))
                txt = src_cache[p][ln[1] - 1].strip()[:90]
    top = ",".join("%s:%d%%" % (k[6:], 100 * v / max(s, 1)) for k, v in stalls[ln].most_common(3))
    print("%-22s smp %5.1f%% ins %5.1f%% [%s] %s" % ("%s:%s" % ln if ln else "?", 100 * s / ts, 100 * I[ln] / ti, top, txt))
