"""Shared parity harness: drives the CPU oracle (oracle/) and a SwarmPlanner (C-ABI library: the CUDA
build or the test-only host simulator) in lock step on identical inputs and reports the differences
of every stage.  Test infrastructure.

Every step is "teacher forced": the planner's device state (records, SFC boxes, planner_seq) is
overwritten with the oracle's state before the step, so each step is an independent comparison and
a 1e-9 difference in one QP cannot leak into the bit-exact comparisons of the next step.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from dlsc_gc_planner_b200 import capi, missions  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")
HOSTSIM_SO = os.path.join(HOSTSIM_DIR, "libdlsc_hostsim.so")


def hostsim_lib():
    subprocess.check_call(["make", "-C", HOSTSIM_DIR, "all"], stdout=subprocess.DEVNULL)
    return capi.load_library(HOSTSIM_SO)


def oracle_params(cfg, mission):
    return O.make_params(M=cfg.M, n=cfg.n, phi=cfg.phi, dim=cfg.dim, use_sfc=cfg.use_sfc, dt=cfg.dt,
                         world_min=mission.world_min, world_max=mission.world_max, world_res=cfg.world_res,
                         grid_res=cfg.grid_res, z_2d=cfg.z_2d, comm_range=cfg.comm_range,
                         w_control=cfg.w_control, w_terminal=cfg.w_terminal, reset_threshold=cfg.reset_threshold)


def make_oracle(cfg, mission, max_nbr, n_threads=1):
    p = oracle_params(cfg, mission)
    edt = O.edt_build(p, mission.boxes) if cfg.use_sfc else None
    sw = O.Swarm(p, mission.start, mission.goal, mission.radius, mission.downwash, mission.max_vel,
                 mission.max_acc, mission.nominal_vel, edt=edt, max_nbr=max_nbr, n_threads=n_threads)
    return sw


def records_from_oracle(sw, rec_floats):
    """The per-agent record the planner keeps on the device, built from the oracle's state."""
    N, M, P = sw.N, sw.p.M, sw.p.n + 1
    rec = np.zeros((N, rec_floats), np.float32)
    o = M * P * 3
    rec[:, :o] = sw.traj.reshape(N, -1)
    rec[:, o:o + 3] = sw.pos
    rec[:, o + 3:o + 6] = sw.vel
    rec[:, o + 6:o + 9] = sw.goal_cur
    rec[:, o + 9] = sw.radius.astype(np.float32)
    rec[:, o + 10] = sw.downwash.astype(np.float32)
    return rec


def force_state(pl, sw):
    """Overwrite the planner's state with the oracle's (before a step)."""
    pl.set_records(0, records_from_oracle(sw, pl.rec_floats))
    sl = slice(pl.begin, pl.begin + pl.NL)
    pl.set_agents(acc=sw.acc[sl], waypoint=sw.waypoint[sl], disturbed=sw.disturbed[sl])
    if sw.p.use_sfc:
        pl.set_sfc(sw.sfc[sl], sw.sfc_init[sl])
    pl.seq = sw.seq


OBJ_REL = 1e-5     # north_star: QP objective within 1e-5 relative ...
OBJ_ABS = 5e-8     # ... plus the duality gap the stopping rule admits: 2 solvers x <=6000 rows x mu_tol 1e-12, DESIGN.md


def compare_step(pl, sw):
    """After both sides ran one step from the same state.  Returns a dict of maximum differences."""
    sl = slice(pl.begin, pl.begin + pl.NL)
    out = {}
    out["init_traj"] = float(np.max(np.abs(pl.init_traj() - sw.init_traj[sl])))
    out["pred_traj"] = float(np.max(np.abs(pl.pred_traj() - sw.pred_traj)))
    idx, cnt = pl.neighbours()
    out["nbr_cnt"] = int(np.max(np.abs(cnt - sw.nbr_cnt[sl])))
    mask = np.arange(pl.K)[None, :] < cnt[:, None]
    out["nbr_idx"] = int(np.max(np.abs((idx - sw.nbr_idx[sl]) * mask))) if mask.any() else 0
    normal, anchor, d = pl.lsc()
    m4 = mask[:, :, None, None]
    out["lsc_normal"] = float(np.max(np.abs((normal - sw.lsc_normal[sl]) * m4))) if mask.any() else 0.0
    out["lsc_d"] = float(np.max(np.abs((d - sw.lsc_d[sl]) * m4))) if mask.any() else 0.0
    out["lsc_anchor"] = float(np.max(np.abs((anchor - sw.lsc_anchor[sl]) * m4[..., None]))) if mask.any() else 0.0
    if sw.p.use_sfc:
        out["sfc"] = float(np.max(np.abs(pl.sfc() - sw.sfc[sl])))
    out["goal"] = float(np.max(np.abs(pl.goal() - sw.goal_cur[sl])))
    st_p, st_o = pl.status(), sw.status[sl]
    # NBR_OVERFLOW is compared through the neighbour lists themselves (the oracle raises the flag for the whole swarm,
    # the kernels per agent); the other failure bits must agree agent by agent
    mask = capi.FAIL_MASK & ~capi.NBR_OVERFLOW
    out["status_mismatch"] = int(np.sum((st_p & mask) != (st_o & mask)))
    ok = ((st_p | st_o) & mask) == 0
    cost_p, cost_o = pl.cost(), sw.cost[sl]
    excess = np.abs(cost_p - cost_o) - OBJ_REL * np.abs(cost_o)
    out["obj_excess"] = float(np.max(excess[ok])) if ok.any() else 0.0   # must stay <= OBJ_ABS
    out["violation"] = float(np.max(pl.violation()[ok])) if ok.any() else 0.0
    dx = np.abs(pl.qp_x() - sw.qp_x[sl]).reshape(pl.NL, -1).max(axis=1)
    out["x"] = float(np.max(dx[ok])) if ok.any() else 0.0
    # control points are unique only off degenerate (weakly active) rows, where the oracle's interior point is defined to
    # O(sqrt(mu_tol)) = 1e-6: count the agents beyond 1e-6 instead of loosening the bound for everyone
    out["x_loose_agents"] = int(np.sum(dx[ok] > 1e-6)) if ok.any() else 0
    out["ok_agents"] = int(ok.sum())
    out["traj"] = float(np.max(np.abs(pl.traj() - sw.traj[sl])))
    out["iters_planner"] = int(pl.qp_iters().max())
    out["iters_oracle"] = int(sw.qp_iters[sl].max())
    return out


SUMMED = ("x_loose_agents", "ok_agents")


def merge_max(acc, new):
    for k, v in new.items():
        acc[k] = acc.get(k, 0) + v if k in SUMMED else max(acc.get(k, 0), v)
    return acc


def run_lockstep(pl, sw, mission, steps, waypoint_fn=None, check=None):
    """Teacher-forced lock-step rollout.  Returns the per-key maxima over the rollout."""
    worst = {}
    desired = sw.goal_des
    for _ in range(steps):
        if waypoint_fn is not None:
            sw.waypoint = waypoint_fn(sw)
        force_state(pl, sw)
        sw.step()
        pl.plan()
        diff = compare_step(pl, sw)
        if check is not None:
            check(diff)
        merge_max(worst, diff)
        sw.advance()
    return worst


def default_waypoints(cfg, mission):
    occupied = missions.occupied_nodes(mission.boxes, cfg.grid_res) if len(mission.boxes) else None
    router = None
    if occupied and mission.n_agents <= 128:
        goal = mission.goal
        router = missions.LatticeRouter(mission.world_min, mission.world_max, cfg.grid_res, occupied, goal)

    def fn(sw):
        return missions.next_waypoints(sw.waypoint, sw.goal_cur, sw.goal_des, sw.traj if sw.seq > 0 else None,
                                       sw.pos, cfg, occupied, router)
    return fn


_MISSION_FILES = {}


def load_case(name, index=None):
    """Reference missions named by the BASELINE configs, from the committed fixtures tests/golden/missions.npz (mission
    #1 of every family) and missions_all.npz (index = 1..30), generated from /root/reference by
    tests/golden/make_fixtures.py."""
    fn = "missions.npz" if index is None else "missions_all.npz"
    if fn not in _MISSION_FILES:
        _MISSION_FILES[fn] = np.load(os.path.join(ROOT, "tests", "golden", fn))
    z = _MISSION_FILES[fn]
    cfg = {"empty10": missions.PlannerConfig.empty, "empty50": missions.PlannerConfig.empty,
           "empty70": missions.PlannerConfig.empty, "forest10": missions.PlannerConfig.forest3d,
           "maze10": missions.PlannerConfig.maze2d}[name]()
    g = lambda f: z[name + "/" + f] if index is None else z[f"{name}/{index}/{f}"]
    m = missions.Mission(g("world_min"), g("world_max"), g("start"), g("goal"), g("radius"), g("downwash"),
                         g("max_vel"), g("max_acc"), g("nominal_vel"), g("boxes"))
    return cfg, m


def subset(mission, n):
    """First n agents of a mission."""
    return missions.Mission(mission.world_min, mission.world_max, mission.start[:n], mission.goal[:n],
                            mission.radius[:n], mission.downwash[:n], mission.max_vel[:n], mission.max_acc[:n],
                            mission.nominal_vel[:n], mission.boxes)
