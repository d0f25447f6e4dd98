"""Generates tests/golden/pibt_ref.npz: known-answer plans of the REFERENCE'S OWN PIBT object code (oracle/_ref/libmapf_ref.so =
/root/reference/src/mapf/{pibt,solver,problem,plan,paths,lib_cbs}.cpp + third_party/grid-pathfinding/graph compiled
unmodified by oracle/Makefile) on random lattices: 2-D and 3-D, with and without missing nodes, 2..14 agents, including
corridor swaps and agents already at their goals.  Run in the build container:  python tests/golden/make_pibt_fixtures.py"""
import ctypes as C
import os
import sys
from collections import deque

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ref_pibt(lib, w, d, h, exists, start, cur, goal, max_t=6000):
    plan = np.zeros((max_t, len(cur)), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    T = lib.ref_pibt_solve(w, d, h, p(exists), len(cur), p(start), p(cur), p(goal), max_t, p(plan))
    assert T > 0
    return plan[:T].copy()


def ref_pibt_obs(lib, w, d, h, exists, warning, start, cur, goal, obs_node, obs_dist, max_t=6000):
    plan = np.zeros((max_t, len(cur)), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    T = lib.ref_pibt_solve_obs(w, d, h, p(exists), p(warning), len(cur), p(start), p(cur), p(goal), p(obs_node), p(obs_dist), max_t, p(plan))
    assert T > 0
    return plan[:T].copy()


def obs_problems(rng, count):
    """problems() plus the dynamic-obstacle inputs of the planner: warning flags on a random blob of nodes (the reachable
    region of an obstacle) and, for some agents, a closest obstacle of interest (node + distance)."""
    out = []
    for (w, d, h, exists, start, cur, goal) in problems(rng, count):
        ids = np.flatnonzero(exists)
        warning = np.zeros(w * d * h, np.uint8)
        for _ in range(int(rng.integers(1, 4))):
            c = int(rng.choice(ids)); z, r = divmod(c, w * d); y, x = divmod(r, w)
            rad = float(rng.uniform(0.8, 2.6))
            for v in ids:
                vz, vr = divmod(int(v), w * d); vy, vx = divmod(vr, w)
                if (vx - x) ** 2 + (vy - y) ** 2 + (vz - z) ** 2 <= rad * rad:
                    warning[v] = 1
        n = len(cur)
        obs_node = np.where(rng.random(n) < 0.5, rng.choice(ids, size=n), -1).astype(np.int32)
        obs_dist = rng.uniform(0.3, 6.0, n).astype(np.float32)
        out.append((w, d, h, exists, warning, start, cur, goal, obs_node, obs_dist))
    return out


def component(w, d, h, exists, seed):
    """node ids of the connected component of `seed` (6-neighbourhood)"""
    seen = {seed}
    q = deque([seed])
    while q:
        v = q.popleft()
        z, r = divmod(v, w * d); y, x = divmod(r, w)
        for dx, dy, dz in ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)):
            X, Y, Z = x + dx, y + dy, z + dz
            if 0 <= X < w and 0 <= Y < d and 0 <= Z < h:
                u = w * d * Z + w * Y + X
                if exists[u] and u not in seen:
                    seen.add(u); q.append(u)
    return sorted(seen)


def problems(rng, count):
    out = []
    while len(out) < count:
        three_d = rng.random() < 0.3
        w, d, h = (int(rng.integers(3, 8)), int(rng.integers(3, 8)), int(rng.integers(2, 4))) if three_d else (int(rng.integers(3, 18)), int(rng.integers(2, 12)), 1)
        exists = (rng.random(w * d * h) > rng.choice([0.0, 0.1, 0.25])).astype(np.uint8)
        free = np.flatnonzero(exists)
        if len(free) < 4:
            continue
        comp = component(w, d, h, exists, int(rng.choice(free)))
        n = int(rng.integers(2, 15))
        if len(comp) < n + 1:
            continue
        cur = rng.choice(comp, size=n, replace=False).astype(np.int32)
        goal = rng.choice(comp, size=n, replace=False).astype(np.int32)
        if rng.random() < 0.2:
            goal[: n // 2] = cur[: n // 2]                  # some agents already at their goals
        start = cur.copy()
        if rng.random() < 0.5:                             # start point differs from the current waypoint (later replans)
            start = rng.choice(comp, size=n, replace=False).astype(np.int32)
        out.append((w, d, h, exists, start, cur, goal))
    return out


def main():
    O.build()
    so = os.path.join(ROOT, "oracle", "_ref", "libmapf_ref.so")
    if not os.path.exists(so):
        raise SystemExit("oracle/_ref/libmapf_ref.so not built (needs /root/reference)")
    lib = C.CDLL(so)
    rng = np.random.default_rng(20261019)
    data = {}
    probs = problems(rng, 80)
    for i, (w, d, h, exists, start, cur, goal) in enumerate(probs):
        plan = ref_pibt(lib, w, d, h, exists, start, cur, goal)
        data["%d/dims" % i] = np.array([w, d, h], np.int32)
        data["%d/exists" % i] = exists
        data["%d/start" % i] = start; data["%d/cur" % i] = cur; data["%d/goal" % i] = goal
        data["%d/plan" % i] = plan
    data["count"] = np.array(len(probs))
    np.savez_compressed(os.path.join(OUT, "pibt_ref.npz"), **data)
    obs = {}
    oprobs = obs_problems(np.random.default_rng(20261020), 40)
    for i, (w, d, h, exists, warning, start, cur, goal, obs_node, obs_dist) in enumerate(oprobs):
        obs["%d/dims" % i] = np.array([w, d, h], np.int32)
        for k, v in (("exists", exists), ("warning", warning), ("start", start), ("cur", cur), ("goal", goal), ("obs_node", obs_node),
                     ("obs_dist", obs_dist)):
            obs["%d/%s" % (i, k)] = v
        obs["%d/plan" % i] = ref_pibt_obs(lib, w, d, h, exists, warning, start, cur, goal, obs_node, obs_dist)
    obs["count"] = np.array(len(oprobs))
    np.savez_compressed(os.path.join(OUT, "pibt_obs_ref.npz"), **obs)
    print("obstacle problems", len(oprobs), "plan lengths", sorted(len(obs["%d/plan" % i]) for i in range(len(oprobs)))[-5:])
    print("problems", len(probs), "plan lengths", sorted(len(data["%d/plan" % i]) for i in range(len(probs)))[-5:])


if __name__ == "__main__":
    main()
