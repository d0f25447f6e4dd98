#!/usr/bin/env python
"""ncu_summary.py report.ncu-rep [out.csv]: per-kernel summary of a --set full capture: time, occupancy, issue, pipes, DRAM bytes
and the top stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active)."""
import csv, subprocess, sys

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


rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
want = [w for w in want if w in hdr]
stall = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
out = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = kname(r[hdr.index("Kernel Name")])
    d = {"kernel": name}
    for w in want:
        d[w] = r[hdr.index(w)].replace(",", "") + " " + units[hdr.index(w)]
    st = sorted(((float(r[hdr.index(h)].replace(",", "") or 0), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)
    d["stalls (warps per issue)"] = " ".join("%s=%.2f" % (n, v) for v, n in st[:7])
    out.append(d)
    print(name)
    for k, v in d.items():
        if k != "kernel":
            print("   %-60s %s" % (k, v))
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        w = csv.writer(f)
        keys = list(out[0].keys())
        w.writerow(keys)
        for d in out:
            w.writerow([d[k] for k in keys])
