#!/usr/bin/env python
"""Diagnostic: distribution of dual-active-set iterations / neighbour counts over the bench swarm in transit."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dlsc_gc_planner_b200 import capi, missions

class A: pass
args = A(); args.agents = int(os.environ.get("AGENTS", 4096)); args.half_extent = None; args.max_nbr = 96; args.settle = 25
cfg, m = bench.make_world(args)
pl = capi.SwarmPlanner(cfg, m, max_nbr=args.max_nbr, device=0)
pl.build_edt(m.boxes)
occupied = missions.occupied_nodes(m.boxes, cfg.grid_res)
wp = pl.start.copy(); goal_des = m.goal.astype(np.float32); traj = None
for t in range(args.settle + 6):
    pos, vel, acc = pl.state(); goal_cur = pl.goal()
    wp = missions.next_waypoints(wp, goal_cur, goal_des, traj, pos, cfg, occupied)
    pl.set_agents(waypoint=wp); pl.plan(); traj = pl.traj()
    if t >= args.settle:
        it = pl.qp_iters(); _, cnt = pl.neighbours(); st = pl.status()
        h = np.bincount(np.minimum(it, 60), minlength=61)
        print("step", t, "iters sum", it.sum(), "max", it.max(), "zero", int((it == 0).sum()), "1-5", int(((it > 0) & (it <= 5)).sum()),
              "6-20", int(((it > 5) & (it <= 20)).sum()), ">20", int((it > 20).sum()), "ipm", int(((st & capi.QP_IPM_USED) != 0).sum()) if hasattr(capi, "QP_IPM_USED") else "?",
              "nbr mean", cnt.mean(), "max", cnt.max(), "counters", pl.counters())
        print("   hist", h.tolist())
        if os.environ.get("DLSC_B200_LIB", "").find("cycles") >= 0:      # -DDLSC_QP_CYCLES build: viol = CTA residency in clocks
            cyc = pl.violation()
            import ctypes as C
            raw = (C.c_int64 * 16)(); pl.lib.dlsc_get_counters(pl.ctx, raw)
            print("   GI kernel, Mcycles of thread 0 summed over its agents: prologue %.1f map_x %.1f scan(+far2) %.1f cand+Hinv %.1f w=Hinv*a %.1f "
                  "serial %.1f yupdate %.1f | total residency %.1f  (agents %d, iterations %d)" % (
                      tuple(raw[k] / 1e6 for k in range(9, 16)) + (cyc[it > 0].sum() / 1e6 if os.environ.get("DLSC_QP_FAST") != "0" else cyc.sum() / 1e6,
                                                                    int((it > 0).sum()), int(it.sum()))))
            hv = int(os.environ.get("HEAVY", "0"))
            if hv:
                print("   (phase sums above are over the %d agents with >= %d iterations: %d iterations, residency %.1f Mcycles)" % (
                    int((it >= hv).sum()), hv, int(it[it >= hv].sum()), cyc[it >= hv].sum() / 1e6))
            for lo, hi in ((0, 0), (1, 1), (2, 2), (3, 5), (6, 20), (21, 1000)):
                sel = (it >= lo) & (it <= hi)
                if sel.any():
                    print("   iters %d-%d: agents %d, mean kcycles %.1f, p50 %.1f, max %.1f, share of cycles %.3f" % (
                        lo, hi, sel.sum(), cyc[sel].mean() / 1e3, np.median(cyc[sel]) / 1e3, cyc[sel].max() / 1e3, cyc[sel].sum() / cyc.sum()))
    pl.advance()
