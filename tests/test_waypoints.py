"""The waypoint provider behind the C ABI (dlsc_wp_*, dlsc_gc_planner_b200/csrc/dlsc_waypoints.cpp): comm-range groups + PIBT on
the lattice + the waypoint update rules of the reference's MultiSyncSimulator::decentralizedMAPP.

  * PIBT against the REFERENCE'S OWN object code (src/mapf/*.cpp + third_party/grid-pathfinding compiled unmodified into
    oracle/_ref/libmapf_ref.so): 80 committed known-answer plans (tests/golden/pibt_ref.npz, 2-D / 3-D lattices, missing
    nodes, up to 14 agents, unsolvable instances that run to the 5000-step cap) and, when oracle/_ref is built, live on
    fresh random instances.  Plans must be identical, node by node.
  * the whole layer against the reference's recorded run: the 137 waypoint sets recovered from the CPLEX log
    (tests/golden/inferred_waypoints.npz) are reproduced step by step when the provider is fed the rollout's states --
    i.e. the golden-log chain now runs with GENERATED waypoints.
Host code: runs on the CPU tier through the product library itself (no GPU needed for dlsc_wp_*)."""
import ctypes as C
import os

import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi, missions

GOLD = os.path.join(_parity.ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def lib():
    capi.build_library()
    return capi.load_library()


def _provider(lib, n):
    cfg = missions.PlannerConfig.maze2d()
    z = np.zeros((n, 3), np.float32)
    one = np.ones(n)
    m = missions.Mission(np.array([-2, -2, 0], np.float32), np.array([2, 2, 2], np.float32), z, z.copy(), 0.15 * one, 2.0 * one,
                         one, 2 * one, one, np.zeros((0, 6), np.float32))
    return capi.WaypointProvider(cfg, m, lib=lib)


def test_pibt_matches_reference_known_answers(lib):
    z = np.load(os.path.join(GOLD, "pibt_ref.npz"))
    long_plans = 0
    for i in range(int(z["count"])):
        g = lambda f: z["%d/%s" % (i, f)]
        wp = _provider(lib, len(g("cur")))
        wp.set_nodes(g("dims"), g("exists"))
        plan = wp.pibt(g("start"), g("cur"), g("goal"))
        assert plan.shape == g("plan").shape and np.array_equal(plan, g("plan")), i
        long_plans += len(plan) > 1000
        wp.close()
    assert long_plans >= 3          # the iteration cap is part of the pinned behaviour


def test_pibt_matches_reference_object_code_live(lib):
    so = os.path.join(_parity.ROOT, "oracle", "_ref", "libmapf_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    import sys
    sys.path.insert(0, GOLD)
    import make_pibt_fixtures as mk
    ref = C.CDLL(so)
    rng = np.random.default_rng(77)
    for (w, d, h, exists, start, cur, goal) in mk.problems(rng, 40):
        want = mk.ref_pibt(ref, w, d, h, exists, start, cur, goal)
        wp = _provider(lib, len(cur))
        wp.set_nodes((w, d, h), exists)
        got = wp.pibt(start, cur, goal)
        assert got.shape == want.shape and np.array_equal(got, want)
        wp.close()


def test_pibt_with_dynamic_obstacles_matches_reference(lib):
    """The dynamic-obstacle side of the planner's PIBT: warning nodes (directed edges) and the closest obstacle of interest
    per agent (priority, tie-breaking) -- 40 committed known-answer plans of the reference's object code, and live when
    oracle/_ref is built."""
    z = np.load(os.path.join(GOLD, "pibt_obs_ref.npz"))
    differs = 0
    for i in range(int(z["count"])):
        g = lambda f: z["%d/%s" % (i, f)]
        wp = _provider(lib, len(g("cur")))
        wp.set_nodes(g("dims"), g("exists"))
        plain = wp.pibt(g("start"), g("cur"), g("goal"))
        wp.set_warning(g("warning"))
        plan = wp.pibt_obs(g("start"), g("cur"), g("goal"), g("obs_node"), g("obs_dist"))
        assert plan.shape == g("plan").shape and np.array_equal(plan, g("plan")), i
        differs += plan.shape != plain.shape or not np.array_equal(plan, plain)
        wp.set_warning(None)
        assert np.array_equal(wp.pibt(g("start"), g("cur"), g("goal")), plain)          # clearing restores the plain lattice
        wp.close()
    assert differs >= 20                     # the obstacle inputs do change the plans
    so = os.path.join(_parity.ROOT, "oracle", "_ref", "libmapf_ref.so")
    if os.path.exists(so):
        import sys
        sys.path.insert(0, GOLD)
        import make_pibt_fixtures as mk
        ref = C.CDLL(so)
        for (w, d, h, exists, warning, start, cur, goal, on, od) in mk.obs_problems(np.random.default_rng(78), 25):
            want = mk.ref_pibt_obs(ref, w, d, h, exists, warning, start, cur, goal, on, od)
            wp = _provider(lib, len(cur))
            wp.set_nodes((w, d, h), exists)
            wp.set_warning(warning)
            got = wp.pibt_obs(start, cur, goal, on, od)
            assert got.shape == want.shape and np.array_equal(got, want)
            wp.close()


def test_lattice_nodes_follow_the_distance_grid(lib, oracle):
    """updateGridMap: nodes inside inflated obstacles are dropped; maze10 #1 keeps its start and goal nodes."""
    cfg, m = _parity.load_case("maze10")
    sw = _parity.make_oracle(cfg, m, 9)
    wp = capi.WaypointProvider(cfg, m, lib=lib, edt=(sw.edt.dist, sw.edt.obst, sw.edt.dims, sw.edt.min_key, sw.edt.res))
    w, d, h = wp.dims()
    assert (w, d, h) == (17, 9, 1)                                  # world [-2, 6] x [-0.3, 4.3] at 0.5 m: x -2..6, y 0..4
    ex = wp.nodes()
    assert 0 < ex.sum() < ex.size
    # a node is dropped exactly when its point is closer (L-infinity) than the radius to the nearest obstacle cell
    dist, obst = sw.edt.dist, sw.edt.obst
    for y in range(d):
        for x in range(w):
            q = np.array([-2.0 + 0.5 * x, 0.0 + 0.5 * y, 1.0], np.float32)
            cell = np.floor(q.astype(np.float64) / 0.1).astype(int) - np.array(sw.edt.min_key)
            c = (cell[0] * sw.edt.dims[1] + cell[1]) * sw.edt.dims[2] + cell[2]
            blocked = False
            if obst[c][0] >= 0:
                centre = ((obst[c] + np.array(sw.edt.min_key) + 0.5) * 0.1).astype(np.float32)
                near = np.clip(q, centre - np.float32(0.05), centre + np.float32(0.05))
                blocked = np.abs(near - q).max() < 0.15 - 1e-5
            assert ex[w * y + x] == (0 if blocked else 1), (x, y)
    for a in range(m.n_agents):                                     # start and goal nodes of the mission are free
        for pt in (m.start[a], m.goal[a]):
            assert ex[w * int(round((pt[1] - 0.0) / 0.5)) + int(round((pt[0] + 2.0) / 0.5))] == 1
    wp.close()


def test_generated_waypoints_reproduce_the_reference_run(lib, oracle):
    """Golden-log chain with generated waypoints: at every one of the 137 steps the provider, fed the oracle rollout's
    states, issues exactly the waypoints recovered from the reference's log."""
    cfg, m = _parity.load_case("maze10")
    sw = _parity.make_oracle(cfg, m, 9, n_threads=os.cpu_count() or 1)
    wp = capi.WaypointProvider(cfg, m, lib=lib, edt=(sw.edt.dist, sw.edt.obst, sw.edt.dims, sw.edt.min_key, sw.edt.res))
    want = np.load(os.path.join(GOLD, "inferred_waypoints.npz"))["waypoints"]
    for step in range(len(want)):
        got = wp.step(sw.pos, sw.goal_cur, sw.traj if sw.seq > 0 else None, sw.waypoint)
        assert np.array_equal(got, want[step]), (step, np.flatnonzero((got != want[step]).any(axis=1)), got, want[step])
        sw.waypoint = got
        st = sw.step()
        assert (st & ~16).max() == 0
        sw.advance()
    assert wp.pibt_timesteps() > 0
    wp.close()


def _open_world(lib, starts, goals, comm_range=-1.0):
    cfg = missions.PlannerConfig.maze2d()
    cfg.comm_range = comm_range
    n = len(starts)
    one = np.ones(n)
    s = np.array([[x, y, 1.0] for x, y in starts], np.float32); g = np.array([[x, y, 1.0] for x, y in goals], np.float32)
    m = missions.Mission(np.array([-4, -4, 0], np.float32), np.array([4, 4, 2], np.float32), s, g, 0.15 * one, 2.0 * one, one, 2 * one,
                         one, np.zeros((0, 6), np.float32))
    return cfg, m, capi.WaypointProvider(cfg, m, lib=lib)


def _walk(wp, m, steps, on_step=None):
    """agents that follow their waypoints exactly (position = current goal = waypoint): the waypoint sequence the layer issues"""
    cur = m.start.copy()
    seq = [cur.copy()]
    for s in range(steps):
        if on_step:
            on_step(s, cur)
        cur = wp.step(cur, cur, None, cur)
        seq.append(cur.copy())
    return np.array(seq)


def test_waypoints_go_around_an_obstacle_region(lib):
    """GridBasedPlanner with a dynamic obstacle (src/grid_based_planner.cpp:64-150): the lattice nodes the obstacle can reach
    within the horizon are warning nodes, no edge leads from a clear node into one, so the issued waypoints route around
    the region; without the obstacle the agent walks the straight row."""
    cfg, m, wp = _open_world(lib, [(-3.0, 0.0)], [(3.0, 0.0)])
    w, d, h = wp.dims()
    straight = _walk(wp, m, 14)
    assert np.all(straight[:, 0, 1] == 0.0) and straight[-1, 0, 0] == 3.0
    wp.close()
    cfg, m, wp = _open_world(lib, [(-3.0, 0.0)], [(3.0, 0.0)])
    wp.set_obstacles([[0.0, 0.0, 1.0]], [[0.0, 0.0, 0.0]], radius=0.6, max_acc=0.0)
    around = _walk(wp, m, 20)
    warn = wp.warning().reshape(d, w)
    # reachable region: |node - obstacle| < 0.15 + 0.6: the 3 x 3 block of nodes around the origin
    ys, xs = np.nonzero(warn)
    assert set(zip((xs * 0.5 - 4).tolist(), (ys * 0.5 - 4).tolist())) == {(x, y) for x in (-0.5, 0.0, 0.5) for y in (-0.5, 0.0, 0.5)}
    for q in around[:, 0]:
        assert not warn[int(round((q[1] + 4) / 0.5)), int(round((q[0] + 4) / 0.5))], q       # never steps into the region
    assert np.allclose(around[-1, 0, :2], [3.0, 0.0]) and np.any(around[:, 0, 1] != 0.0)     # arrives, by a detour
    wp.set_obstacles(None)
    again = wp.step(m.start, m.start, None, m.start)
    assert not wp.warning().any() and again[0, 1] == 0.0
    wp.close()


def test_agent_inside_the_region_escapes(lib):
    """updateDOI / updateGoal (:192-299): an agent whose waypoint lies inside an obstacle's reachable region takes that
    obstacle as its obstacle of interest and gets an escape goal down the obstacle-cost slope: its next waypoints move away
    from the obstacle although its desired goal lies on the other side; a collision alert makes an obstacle one of interest
    even when the waypoint is out of its reach."""
    cfg, m, wp = _open_world(lib, [(-0.5, 0.0)], [(3.0, 0.0)])
    wp.set_obstacles([[0.0, 0.0, 1.0]], [[0.0, 0.0, 0.0]], radius=0.6, max_acc=0.0)
    seq = _walk(wp, m, 4)
    d0 = np.linalg.norm(seq[:, 0, :2], axis=1)
    assert d0[1] > d0[0] and d0[2] >= d0[1], seq[:, 0]                   # away from the obstacle at the origin first
    wp.close()
    # alert: the obstacle is far from the waypoint, yet the agent treats it as its obstacle of interest
    cfg, m, wp = _open_world(lib, [(-3.0, 0.0)], [(3.0, 0.0)])
    wp.set_obstacles([[-1.5, 0.0, 1.0]], [[0.0, 0.0, 0.0]], radius=0.2, max_acc=0.0)
    plain = wp.step(m.start, m.start, None, m.start)
    wp.close()
    cfg, m, wp = _open_world(lib, [(-3.0, 0.0)], [(3.0, 0.0)])
    wp.set_obstacles([[-1.5, 0.0, 1.0]], [[0.0, 0.0, 0.0]], radius=0.2, max_acc=0.0)
    wp.set_alerts([[0]])
    alerted = wp.step(m.start, m.start, None, m.start)
    assert plain[0, 0] == -2.5 and alerted[0, 0] <= -3.0, (plain, alerted)          # towards the goal vs. away from the obstacle
    wp.close()
