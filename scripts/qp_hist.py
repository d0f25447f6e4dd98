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
cfg, m, edt = bench.make_world(args)
pl = capi.SwarmPlanner(cfg, m, max_nbr=args.max_nbr, device=0)
pl.set_edt(*edt, cfg.world_res)
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
    pl.advance()
