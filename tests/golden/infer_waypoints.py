#!/usr/bin/env python
"""Recover the waypoints the reference's PIBT provider issued during its only recorded run, from the run's log.

The reference log (log/result_1742185870.978562_DLSCGC_10agents.csv, maze10_dense #1, CPLEX, 34 s) records every
agent's state each 0.1 s but not the waypoints, which come from the MAPF layer (out of the hot path's scope, SURVEY
s8(f) rank 3).  A waypoint is a node of the 0.5 m lattice and moves by at most one 4-connected step per replan
(src/multi_sync_simulator.cpp:407-447), and within a step every agent's replan depends on the other agents only
through previous-step data.  So the waypoint of every agent and step can be found independently: try the five moves
from every waypoint still compatible with the log so far, keep those whose replan (oracle: LSC + SFC + goal + QP +
state step) reproduces the two logged rows of that step, continue the rollout with the closest one.  Wrong moves
miss by >= 0.07 m; the right one matches position to 1e-6 for the first 20 chained replans and to 1e-4 (the drift
between the oracle's interior point and CPLEX in closed loop) for 137 replans = 27.4 s, where one oracle QP hits
its iteration cap.

Output: tests/golden/inferred_waypoints.npz (waypoints [steps][10][3], moves matching per agent and step) -- the
input of tests/test_golden_log_chain.py.  Reads only tests/golden/golden_log_full.npz.
Run from the repository root:  python tests/golden/infer_waypoints.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import _parity  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

FIELDS = ("pos", "vel", "acc", "goal_cur", "waypoint", "traj", "sfc", "sfc_init", "disturbed")
DELTAS = [(0.0, 0.0), (0.5, 0.0), (-0.5, 0.0), (0.0, 0.5), (0.0, -0.5)]
SCALE = 100.0        # search tolerance: 100 x (position 1e-6, velocity 3e-5, acceleration 6e-4) + the print resolution


def tolerance(ref, scale):
    """one unit of the last printed digit (6 significant) + the solver difference allowed per quantity
    (control points agree to ~1e-6 x scale; velocity x n/dt, acceleration x n(n-1)/dt^2)"""
    return scale * np.array([1e-6] * 3 + [3e-5] * 3 + [6e-4] * 3) + 10.0 ** (np.floor(np.log10(np.maximum(np.abs(ref), 1e-30))) - 5)


def score(sw, a, step, state):
    """(within tolerance, largest position error) of agent a's trajectory against the two logged rows of `step`"""
    ok, worst = True, 0.0
    for k, t in ((2 * step + 1, 0.1), (2 * step + 2, 0.2)):
        s = O.state_at(sw.p, sw.traj[a], t).reshape(9)
        ok = ok and bool(np.all(np.abs(s - state[k, a]) <= tolerance(state[k, a], SCALE)))
        worst = max(worst, float(np.abs(s - state[k, a])[:3].max()))
    return ok, worst


def main():
    cfg, m = _parity.load_case("maze10")
    sw = _parity.make_oracle(cfg, m, 9, n_threads=os.cpu_count() or 1)
    z = np.load(os.path.join(HERE, "golden_log_full.npz"))
    state = z["state"]
    n_steps = (len(z["t"]) - 1) // 2
    N = m.n_agents
    hyp = [[tuple(np.round(sw.pos[a, :2].astype(np.float64), 3))] for a in range(N)]
    chosen, n_hyp = [], []
    for step in range(n_steps):
        s0 = ({k: getattr(sw, k).copy() for k in FIELDS}, sw.seq)

        def restore():
            for k, v in s0[0].items():
                getattr(sw, k)[...] = v
            sw.seq = s0[1]
        cand = []
        for a in range(N):
            c = []
            for w in hyp[a]:
                for d in DELTAS:
                    q = (round(w[0] + d[0], 3), round(w[1] + d[1], 3))
                    if q not in c:
                        c.append(q)
            cand.append(c)
        match = [[] for _ in range(N)]
        for r in range(max(len(c) for c in cand)):
            restore()
            w = sw.waypoint.copy()
            for a in range(N):
                w[a, :2] = cand[a][min(r, len(cand[a]) - 1)]
            sw.waypoint = w
            sw.step()
            for a in range(N):
                if r < len(cand[a]):
                    ok, e = score(sw, a, step, state)
                    if ok:
                        match[a].append((e, cand[a][r]))
        if any(not x for x in match):
            print("step", step, "no move reproduces the log for agents", [a for a in range(N) if not match[a]], "-- stopping")
            break
        match = [[q for _, q in sorted(x, key=lambda y: y[0])] for x in match]
        restore()
        w = sw.waypoint.copy()
        for a in range(N):
            w[a, :2] = match[a][0]
        sw.waypoint = w
        st = sw.step()
        if (st & ~16).max() != 0:
            print("step", step, "an oracle QP did not converge -- stopping before it")
            break
        sw.advance()
        hyp = match
        chosen.append(w.copy())
        n_hyp.append([len(x) for x in match])
        print("step", step, "moves matching per agent", n_hyp[-1])
    np.savez_compressed(os.path.join(HERE, "inferred_waypoints.npz"), waypoints=np.array(chosen, np.float32),
                        n_hyp=np.array(n_hyp, np.int32), scale=SCALE)
    print("steps recovered:", len(chosen))


if __name__ == "__main__":
    main()
