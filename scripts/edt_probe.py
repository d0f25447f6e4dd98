import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
from dlsc_gc_planner_b200 import capi, missions
cfg = missions.PlannerConfig.forest3d(); m = missions.synthetic_forest(n_agents=4096, half_extent=32.0, seed=4096)
pl = capi.SwarmPlanner(cfg, m, max_nbr=8)
for i in range(2):
    pl.build_edt(m.boxes)
if len(sys.argv) > 1:
    for i in range(3):
        pl.build_edt(m.boxes); print("edt_build_ms", pl.edt_build_ms())
