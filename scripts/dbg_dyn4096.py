"""Debug: 4096-agent forest with dynamic obstacles -- which agents fail on the GPU, and what does the oracle say about them?"""
import sys, os, time, argparse
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import bench
from dlsc_gc_planner_b200 import capi, missions
args = argparse.Namespace(agents=int(os.environ.get("AGENTS", 4096)), half_extent=None, max_nbr=96, settle=12, steps=0)
cfg, m = bench.make_world(args)
rec, snap = bench.pilot_rollout(cfg, m, args, 0, int(os.environ.get("STEPS", 4)))
pl = capi.SwarmPlanner(cfg, m, max_nbr=args.max_nbr)
pl.build_edt(m.boxes)
bench.restore(pl, snap, slice(0, m.n_agents))
nd = 8
rng = np.random.default_rng(1234)
he = float(m.world_max[0]) * 0.7
opos = np.stack([rng.uniform(-he, he, nd), rng.uniform(-he, he, nd), np.full(nd, 1.0)], axis=1).astype(np.float32)
ang = rng.uniform(0, 2 * np.pi, nd)
ovel = np.stack([np.cos(ang), np.sin(ang), np.zeros(nd)], axis=1).astype(np.float32)
kw = dict(radius=0.3, downwash=1.0, max_acc=2.0, slack_weight=100.0)
sw = bench.oracle_swarm(cfg, m, snap["edt"], snap, rec, args, os.cpu_count())
for t in range(int(os.environ.get("STEPS", 4))):
    wp = rec["wp"][t]
    pl.set_agents(waypoint=wp); pl.set_obstacles(opos, ovel, **kw)
    pl.enable_timing(True)
    t0 = time.perf_counter(); pl.plan(); pl.sync(); t1 = time.perf_counter()
    ms, _ = pl.timings(); pl.enable_timing(False)
    st = pl.status(); it = pl.qp_iters()
    fail = np.where((st & 3) != 0)[0]; ipm = np.where((st & 64) != 0)[0]
    print("step", t, "ms", round(1e3 * (t1 - t0), 2), {k: round(v, 3) for k, v in ms.items()}, "fails", len(fail), "ipm", len(ipm), "max it", it.max(), "it of ipm", it[ipm][:12].tolist(), flush=True)
    # oracle on the same state for the failing / handed-over agents
    rcd = pl.get_records()
    sw.waypoint[...] = wp
    sw.set_obstacles(opos, ovel, **kw)
    chk = list(fail[:6]) + [a for a in ipm if a not in fail][:4]
    gx, gc, gs = pl.qp_x(), pl.cost(), pl.slack()
    for a in chk:
        sw_state = (sw.traj.copy(), sw.goal_cur.copy(), sw.sfc.copy())
        sw.seq = pl.seq - 1
        sw.step(int(a), int(a) + 1)
        print("   agent", a, "gpu status", st[a], "cost", gc[a], "| oracle status", sw.status[a], "cost", sw.cost[a], "it", sw.qp_iters[a], "viol", sw.max_violation[a],
              "nbr", sw.nbr_cnt[a], "minslack", sw.qp_slack[a].min().round(3), gs[a].min().round(3), flush=True)
        sw.traj[...], sw.goal_cur[...], sw.sfc[...] = sw_state
    pl.advance()
    # keep the oracle state in lock step with the GPU
    r = pl.get_records(); o = cfg.M * (cfg.n + 1) * 3
    sw.traj[...] = r[:, :o].reshape(sw.traj.shape); sw.pos[...] = r[:, o:o + 3]; sw.vel[...] = r[:, o + 3:o + 6]; sw.goal_cur[...] = r[:, o + 6:o + 9]
    sw.acc[...] = pl.state()[2]; sw.sfc[...] = pl.sfc(); sw.sfc_init[...] = 0
    opos = opos + ovel * np.float32(cfg.dt)
