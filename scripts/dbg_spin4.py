import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, _parity
from dlsc_gc_planner_b200 import missions as ms, capi
lib = capi.load_library() if len(sys.argv) < 3 else _parity.hostsim_lib()
cfg, m = _parity.load_case("forest10")
sw = _parity.make_oracle(cfg, m, 14, n_threads=os.cpu_count())
pl = capi.SwarmPlanner(cfg, m, max_nbr=14, lib=lib)
pl.set_edt(sw.edt.dist, sw.edt.obst, sw.edt.dims, sw.edt.min_key, sw.edt.res)
wf = _parity.default_waypoints(cfg, m)
for s in range(int(sys.argv[1])):
    st = ms.obstacle_states(ms.SPIN4, s*cfg.dt)
    kw = dict(radius=st["radius"], downwash=st["downwash"], max_acc=st["max_acc"], slack_weight=100.0)
    sw.waypoint = wf(sw)
    sw.set_obstacles(st["pos"], st["vel"], **kw); pl.set_obstacles(st["pos"], st["vel"], **kw)
    _parity.force_state(pl, sw)
    sw.step(); pl.plan()
    d = np.abs(pl.slack()-sw.qp_slack).max(axis=(1,2))
    bad = np.where(d > 1e-5)[0]
    if len(bad):
        a = bad[0]
        print(s, "bad agents", bad.tolist(), "status", pl.status()[bad].tolist(), sw.status[bad].tolist(), "iters", pl.qp_iters()[bad].tolist(),
              "dx", np.abs(pl.qp_x()-sw.qp_x).reshape(10,-1).max(axis=1)[bad].tolist())
        print("  gpu slack", pl.slack()[a].round(4).tolist()); print("  orc slack", sw.qp_slack[a].round(4).tolist())
    sw.advance()
print("done")
