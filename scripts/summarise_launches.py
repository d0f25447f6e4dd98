#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, mean us and the
share of the per-step kernels.   python scripts/summarise_launches.py gpurun_out/launches.csv "<comment>" > profiles/..."""
import collections
import csv
import sys


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


STEP = ("k_predict", "k_neighbours", "k_nbr_bin", "k_nbr_search", "k_lsc", "k_lsc_rest", "k_sfc", "k_goal", "k_qp_fast", "k_qp_gi", "k_qp", "k_advance")
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ki, vi = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    n = kname(r[ki])
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e3
step_total = sum(v[1] / v[0] for k, v in agg.items() if k.split("::")[-1] in STEP)
for c in sys.argv[2:]:
    print("# " + c)
print("kernel,launches,mean_us,share_of_step_kernels")
for k, v in agg.items():
    m = v[1] / v[0]
    print("%s,%d,%.1f,%s" % (k, v[0], m, "%.3f" % (m / step_total) if k.split("::")[-1] in STEP else "-"))
