#!/usr/bin/env python
"""Copy one gpurun result set (gpurun_out/<tag>_bench.json, _launches.csv, _prof.ncu-rep, _memcheck.log, _racecheck.log) into
profiles/ as the round-2 evidence.   usage: python scripts/refresh_profiles.py <tag> [round-prefix, default r2]"""
import csv, json, os, shutil, subprocess, sys

def kname(full):
    """'void k_lsc<0>(DevParams, ...)' -> 'k_lsc' (the <0> instantiations are the swarm-only hot path); '<1>' -> 'k_lsc<dyn>'"""
    n = full.split("(")[0].strip()
    if n.startswith("void "):
        n = n[5:]
    n = n.replace("dlsc::", "")
    if n.endswith("<0>") or n.endswith("<(bool)0>"):
        n = n[:n.index("<")]
    elif n.endswith("<1>") or n.endswith("<(bool)1>"):
        n = n[:n.index("<")] + "<dyn>"
    return n


root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rp = sys.argv[2] if len(sys.argv) > 2 else "r2"
g = lambda s: os.path.join(root, "gpurun_out", tag + s)
P = lambda s: os.path.join(root, "profiles", s)
if os.path.exists(g("_bench.json")):
    shutil.copy(g("_bench.json"), P(rp + "_bench_1gpu.json"))
if os.path.exists(g("_launches.csv")):
    out = subprocess.check_output([sys.executable, os.path.join(root, "scripts", "summarise_launches.py"), g("_launches.csv"),
                                   "ncu launch list, round 2: DLSC_OVERLAP=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mc-missions 0",
                                   "ncu --metrics gpu__time_duration.sum --clock-control none -c 700 (cold-cache, serialised: compare SHARES)"], text=True)
    open(P(rp + "_launch_summary.csv"), "w").write(out)
if os.path.exists(g("_prof.ncu-rep")):
    subprocess.check_call([sys.executable, os.path.join(root, "scripts", "ncu_summary.py"), g("_prof.ncu-rep"), P(rp + "_ncu_full_summary.csv")],
                          stdout=subprocess.DEVNULL)
    raw = subprocess.run(["ncu", "-i", g("_prof.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.split("\n")))
    hdr, units = rows[0], rows[1]
    sc = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}
    f = lambda x: float(x.replace(",", ""))
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    traffic = {}
    for r in rows[2:]:
        if len(r) > max(ir, iw):
            traffic[kname(r[ik])] = f(r[ir]) * sc[units[ir]] + f(r[iw]) * sc[units[iw]]
    json.dump(traffic, open(P("traffic.json"), "w"), indent=1)
    print(traffic)
for kind in ("memcheck", "racecheck"):
    if os.path.exists(g("_%s.log" % kind)):
        lines = [l for l in open(g("_%s.log" % kind)) if l.startswith("=========") or l.startswith("smoke ok")]
        open(P("%s_sanitizer_%s.txt" % (rp, kind)), "w").write(
            "# compute-sanitizer --tool %s --error-exitcode 7 python -c 'import __graft_entry__ as g; g.smoke()'  (gpurun, B200)\n" % kind + "".join(lines))
