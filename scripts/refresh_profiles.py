#!/usr/bin/env python
"""Copy one gpurun result set (gpurun_out/<tag>_bench.json, _ref.json, _launches.csv, _prof.ncu-rep) into profiles/.
usage: python scripts/refresh_profiles.py <tag>"""
import csv, json, os, shutil, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
g = lambda s: os.path.join(root, "gpurun_out", tag + s)
shutil.copy(g("_bench.json"), os.path.join(root, "profiles", "r1_bench_1gpu.json"))
shutil.copy(g("_ref.json"), os.path.join(root, "profiles", "r1_bench_reference_arm.json"))
out = subprocess.check_output([sys.executable, os.path.join(root, "scripts", "summarise_launches.py"), g("_launches.csv"),
                               "ncu launch list, round 1 (final kernels): DLSC_OVERLAP=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline",
                               "ncu --metrics gpu__time_duration.sum --clock-control none -c 700 (cold-cache, serialised: compare SHARES)"], text=True)
open(os.path.join(root, "profiles", "r1_launch_summary.csv"), "w").write(out)
raw = subprocess.run(["ncu", "-i", g("_prof.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [hdr.index(w) for w in want]
sc = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}
f = lambda x: float(x.replace(",", ""))
traffic = {}
with open(os.path.join(root, "profiles", "r1_ncu_full_summary.csv"), "w") as o:
    o.write("# ncu --set full --clock-control none, one launch of each kernel at step 30 of the 4096-agent forest rollout (in transit), "
            "final round-1 code; DLSC_OVERLAP=0 so that kernels are captured alone\n")
    o.write(",".join(want) + "\n" + ",".join(units[i] for i in idx) + "\n")
    for r in rows[2:]:
        if len(r) <= max(idx):
            continue
        o.write(",".join('"%s"' % r[i] if "," in r[i] else r[i] for i in idx) + "\n")
        name = r[idx[0]].split("(")[0]
        traffic[name] = f(r[idx[2]]) * sc[units[idx[2]]] + f(r[idx[3]]) * sc[units[idx[3]]]
        print(name, r[idx[1]], "us inst", r[idx[9]], "issue", r[idx[11]], "regs", r[idx[5]])
json.dump(traffic, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
d = json.load(open(g("_bench.json")))
print({k: d[k] for k in ("value", "ms_per_step", "p50_step_ms")}, d["e2e"]["ms_per_step"], d["stages_ms"], d["edt_build"]["ms"], d["cpu_baseline"]["value"])
