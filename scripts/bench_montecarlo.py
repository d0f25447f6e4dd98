#!/usr/bin/env python
"""BASELINE configs[4]: Monte-Carlo batch of independent empty50 missions replanning in lockstep (K = 49, no map).

    python scripts/bench_montecarlo.py [--missions 128] [--steps 30] [--warmup 5] [--settle 10]

One GPU's share of the 1024-mission / 8-GPU configuration is 128 missions = 6400 agents (the missions are
independent: replicas only, no data-path collective).  Every mission is the reference's
missions/empty50 #1 with seeded goal noise (Mission::addNoise, max_noise 0.2).  Waypoints come from an untimed
pilot rollout with the host-side stand-in provider applied per mission; the timed pass replays them on the device
(plan -> advance chained, CUDA events).  Prints one JSON line (same keys as bench.py where they apply).
Not the driver's bench line (that is bench.py on configs[3]); numbers go to profiles/."""
import argparse
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from dlsc_gc_planner_b200 import capi, missions  # noqa: E402


def load_empty50():
    z = np.load(os.path.join(ROOT, "tests", "golden", "missions.npz"))
    g = lambda f: z["empty50/" + f]
    return missions.PlannerConfig.empty(), missions.Mission(g("world_min"), g("world_max"), g("start"), g("goal"), g("radius"),
                                                            g("downwash"), g("max_vel"), g("max_acc"), g("nominal_vel"), g("boxes"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--missions", type=int, default=128)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--settle", type=int, default=10)
    a = ap.parse_args()
    import torch
    torch.cuda.set_device(0)
    cfg, base = load_empty50()
    ms = [missions.add_goal_noise(base, 0.2, cfg.dim, seed=i) for i in range(a.missions)]
    batch, group = missions.concat_missions(ms)
    n_each = base.n_agents
    N, K = batch.n_agents, n_each - 1
    W, T = max(a.warmup, 3), a.steps

    def planner():
        pl = capi.SwarmPlanner(cfg, batch, max_nbr=K, device=0)
        pl.set_groups(group)
        return pl

    # ---- pilot: records the waypoints of every step and the state at the start of the timed region ----
    pl = planner()
    wp = pl.start.copy()
    goal_des = batch.goal.astype(np.float32)
    traj, rec_wp, snap, fails = None, [], None, 0
    for t in range(a.settle + W + T):
        pos, vel, acc = pl.state()
        goal_cur = pl.goal()
        for i in range(a.missions):
            s = slice(i * n_each, (i + 1) * n_each)
            wp[s] = missions.next_waypoints(wp[s], goal_cur[s], goal_des[s], None if traj is None else traj[s], pos[s], cfg)
        if t == a.settle:
            snap = {"records": pl.get_records(), "acc": acc.copy(), "seq": pl.seq}
        if t >= a.settle:
            rec_wp.append(wp.copy())
        pl.set_agents(waypoint=wp)
        pl.plan()
        traj = pl.traj()
        fails += int(((pl.status() & capi.FAIL_MASK) != 0).sum())
        pl.advance()
    final = traj
    dist_goal = float(np.mean(np.max(np.abs(pl.state()[0] - goal_des), axis=1)))
    pl.close()

    pl = planner()
    pl.set_stream(torch.cuda.current_stream().cuda_stream)
    wp_dev = torch.from_numpy(np.ascontiguousarray(np.array(rec_wp))).cuda()

    def restore():
        pl.set_records(0, snap["records"])
        pl.set_agents(acc=snap["acc"])
        pl.seq = snap["seq"]

    def run(n0, n):
        for t in range(n0, n0 + n):
            pl.set_waypoints_device(wp_dev[t].data_ptr())
            pl.plan(); pl.advance()

    restore(); torch.cuda.synchronize()
    run(0, W); torch.cuda.synchronize()
    l0 = pl.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(T + 1)]
    ev[0].record()
    for t in range(T):
        run(W + t, 1)
        ev[t + 1].record()
    torch.cuda.synchronize()
    launches = pl.launch_count() - l0
    total_ms = ev[0].elapsed_time(ev[T])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(T)]
    exact = bool(np.array_equal(pl.traj(), final))
    restore(); torch.cuda.synchronize()
    run(0, W); torch.cuda.synchronize()
    pl.enable_timing(True)
    run(W, T); torch.cuda.synchronize()
    stage_ms, _ = pl.timings()
    pl.enable_timing(False)
    c = pl.counters()
    it = pl.qp_iters()
    pl.close()
    print(json.dumps({
        "metric": "agent-replans/sec (LSC+QP), Monte-Carlo batch of empty50 missions", "value": N * T / (total_ms * 1e-3),
        "unit": "agent-replans/s", "n_gpus": 1, "steps": T, "warmup": W, "ms_per_step": total_ms / T,
        "p50_step_ms": statistics.median(step_ms), "higher_is_better": True, "scaling": "weak", "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Monte-Carlo batch: %d independent empty50 missions (%d agents, K = %d, M = 5, no map), goal noise 0.2, "
                               "lockstep replans (BASELINE configs[4]: one GPU's share of 1024 missions / 8 GPUs)" % (a.missions, N, K),
                   "agents": N, "settle_steps": a.settle},
        "gpu_launches": int(launches), "stages_ms": stage_ms,
        "work_last_step": {"pairs": c["pairs"], "gjk_iters": c["gjk_iters"], "qp_iters": c["qp_iters"],
                           "agents_with_active_rows": int((it > 0).sum())},
        "pilot": {"qp_failsafe_agents": fails, "mean_dist_to_goal_m": dist_goal, "replay_exact": exact}}))


if __name__ == "__main__":
    main()
