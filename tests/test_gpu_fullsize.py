"""GPU tier at BASELINE.json's full size (the synthetic 4096-agent 3-D forest of bench.py), through properties that
do not need the oracle to replay 4096 agents for many steps:
  * the three SFC answer paths (summed-area query + vertex mask, vertex mask alone, 16-byte EDT records: the
    reference's own per-vertex arithmetic) give identical boxes, bit for bit, along a rollout;
  * the two QP solvers (dual active set, interior point) agree on every agent within the north_star tolerances and
    every returned trajectory satisfies every constraint to 1e-6 m;
  * a sampled block of agents matches the oracle bit for bit (geometry) / within tolerance (QP) at full density;
  * the rollout is deterministic."""
import os

import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi, edt as edtmod, missions

pytestmark = pytest.mark.gpu

N_AGENTS = 4096


@pytest.fixture(scope="module")
def world():
    cfg = missions.PlannerConfig.forest3d()
    m = missions.synthetic_forest(n_agents=N_AGENTS, half_extent=0.5 * float(np.sqrt(N_AGENTS)), seed=4096)
    grid = edtmod.build_edt(m.world_min, m.world_max, cfg.world_res, m.boxes)
    return cfg, m, grid


def make(cuda_lib, world, env=None, **kw):
    cfg, m, grid = world
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        pl = capi.SwarmPlanner(cfg, m, max_nbr=96, lib=cuda_lib, **kw)
        pl.set_edt(*grid, cfg.world_res)
        pl.run_stages(capi.STAGE_SFC)               # builds the vertex mask / summed-area table under `env` ...
        pl.sync()
        pl._ck(pl.lib.dlsc_reset(pl.ctx, capi._p(pl.start)))                    # ... and back to the clean start state
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return pl


def rollout(pls, world, steps, on_step=None, after_advance=None):
    cfg, m, _ = world
    occupied = missions.occupied_nodes(m.boxes, cfg.grid_res)
    lead = pls[0]
    wp = lead.start.copy()
    goal_des = m.goal.astype(np.float32)
    traj = None
    for t in range(steps):
        pos, vel, acc = lead.state()
        wp = missions.next_waypoints(wp, lead.goal(), goal_des, traj, pos, cfg, occupied)
        for pl in pls:
            pl.set_agents(waypoint=wp)
            pl.plan()
        traj = lead.traj()
        if on_step:
            on_step(t)
        for pl in pls:
            pl.advance()
        if after_advance:
            after_advance(t)
    return wp, traj


def test_sfc_paths_identical_at_full_size(cuda_lib, world):
    a = make(cuda_lib, world)
    b = make(cuda_lib, world, env={"DLSC_SFC_SAT": "0"})
    c = make(cuda_lib, world, env={"DLSC_SFC_MASK": "0"})
    seen = {"sat": 0, "mask": 0, "rec": 0}

    def check(t):
        sa = a.sfc()
        assert np.array_equal(sa, b.sfc()) and np.array_equal(sa, c.sfc()), "SFC boxes differ between answer paths at step %d" % t
        assert np.array_equal(a.traj(), b.traj()) and np.array_equal(a.traj(), c.traj())
        ca, cb, cc = a.counters(), b.counters(), c.counters()
        seen["sat"] += ca["sfc_tests_sat"]; seen["mask"] += cb["sfc_tests_mask"]; seen["rec"] += cc["sfc_tests_records"]
        assert cb["sfc_tests_sat"] == 0 and cc["sfc_tests_mask"] == 0 and ca["sfc_tests_records"] == 0
        assert ca["sfc_vertices_alg"] == cb["sfc_vertices_alg"]
        assert cc["sfc_vertices_alg"] <= ca["sfc_vertices_alg"]      # the record path stops counting at the first hit of a test

    rollout([a, b, c], world, 14, check)
    assert seen["sat"] > 0 and seen["mask"] > 0 and seen["rec"] > 0
    for pl in (a, b, c):
        pl.close()


def test_qp_solvers_agree_and_are_feasible_at_full_size(cuda_lib, world):
    gi = make(cuda_lib, world)
    ipm = make(cuda_lib, world, qp_solver=1)
    worst = {"obj": 0.0, "viol": 0.0, "x": 0.0, "active": 0}

    def check(t):
        for pl in (gi, ipm):
            assert (pl.status() & capi.FAIL_MASK).max() == 0
            assert pl.violation().max() <= 1e-6                      # north_star: constraint violation <= 1e-6 m
        cg, ci = gi.cost(), ipm.cost()
        excess = np.abs(cg - ci) - _parity.OBJ_REL * np.abs(ci)      # north_star: objective within 1e-5 relative
        worst["obj"] = max(worst["obj"], float(excess.max()))
        worst["x"] = max(worst["x"], float(np.abs(gi.qp_x() - ipm.qp_x()).max()))
        worst["active"] += int((gi.qp_iters() > 0).sum())

    def adopt(t):
        # keep the two rollouts on the same states (the two solutions differ by ~1e-9 before the float32 store):
        # the interior-point planner adopts the active-set planner's records, accelerations and boxes
        ipm.set_records(0, gi.get_records())
        ipm.set_agents(acc=gi.state()[2])
        ipm.set_sfc(gi.sfc())

    rollout([gi, ipm], world, 10, check, adopt)
    # absolute allowance: the interior point may stop at its "acceptable" level (mu <= 1e-11 over up to ~6000 rows:
    # duality gap <= 6e-8, dlsc_qp.cuh) -- it only matters for agents resting at their goal (objective ~ 0)
    assert worst["obj"] <= 2e-7, worst
    assert worst["x"] <= 1e-4, worst
    assert worst["active"] > 1000                                    # the comparison saw plenty of constrained QPs
    gi.close(); ipm.close()


def test_oracle_block_at_full_density(cuda_lib, world):
    """Agents [0, 64) of the full swarm against the oracle, which replans just that block from the same 4096 records."""
    cfg, m, grid = world
    pl = make(cuda_lib, world)
    wp_prev, traj_prev = rollout([pl], world, 8)                     # get the swarm into transit
    from oracle import oracle_py as O
    p = _parity.oracle_params(cfg, m)
    e = O.Edt(p, grid[0], grid[1], grid[2], grid[3])
    sw = O.Swarm(p, m.start, m.goal, m.radius, m.downwash, m.max_vel, m.max_acc, m.nominal_vel, edt=e, max_nbr=96, n_threads=8)
    o = cfg.M * (cfg.n + 1) * 3
    rec = pl.get_records()
    sw.traj[...] = rec[:, :o].reshape(sw.traj.shape)
    sw.pos[...] = rec[:, o:o + 3]; sw.vel[...] = rec[:, o + 3:o + 6]; sw.goal_cur[...] = rec[:, o + 6:o + 9]
    _, _, acc = pl.state()
    sw.acc[...] = acc; sw.sfc[...] = pl.sfc(); sw.sfc_init[...] = 0
    occupied = missions.occupied_nodes(m.boxes, cfg.grid_res)
    wp = missions.next_waypoints(wp_prev, pl.goal(), m.goal.astype(np.float32), traj_prev, sw.pos, cfg, occupied)
    sw.waypoint[...] = wp
    sw.seq = pl.seq
    nb = 64
    sw.step(0, nb)
    pl.set_agents(waypoint=wp)
    pl.plan()
    idx, cnt = pl.neighbours()
    assert np.array_equal(cnt[:nb], sw.nbr_cnt[:nb])
    valid = np.arange(pl.K)[None, :] < cnt[:nb, None]
    assert np.array_equal(idx[:nb][valid], sw.nbr_idx[:nb][valid])
    normal, anchor, d = pl.lsc()
    assert np.array_equal(normal[:nb][valid], sw.lsc_normal[:nb][valid])
    assert np.array_equal(d[:nb][valid], sw.lsc_d[:nb][valid])
    assert np.array_equal(anchor[:nb][valid], sw.lsc_anchor[:nb][valid])
    assert np.array_equal(pl.sfc()[:nb], sw.sfc[:nb])
    assert np.array_equal(pl.goal()[:nb], sw.goal_cur[:nb])
    ok = ((pl.status()[:nb] | sw.status[:nb]) & capi.FAIL_MASK) == 0
    assert ok.all()
    excess = np.abs(pl.cost()[:nb] - sw.cost[:nb]) - _parity.OBJ_REL * np.abs(sw.cost[:nb])
    assert excess.max() <= _parity.OBJ_ABS
    assert pl.violation()[:nb].max() <= 1e-6
    pl.close()


def test_full_size_rollout_is_deterministic(cuda_lib, world):
    outs = []
    for rep in range(2):
        pl = make(cuda_lib, world)
        rollout([pl], world, 6)
        outs.append((pl.get_records().copy(), pl.sfc().copy()))
        pl.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_qp_row_screen_is_exact_at_full_size(cuda_lib, world):
    """The QP row screen (k_lsc flags + the ball check in gi_solve) only skips rows that cannot be violated: with
    and without it every agent gets the same solution, bit for bit."""
    a = make(cuda_lib, world)
    b = make(cuda_lib, world, qp_screen_slack=-1.0)              # screen off: every scan evaluates every row

    def check(t):
        assert np.array_equal(a.qp_x(), b.qp_x()), "row screen changed a QP solution at step %d" % t
        assert np.array_equal(a.traj(), b.traj()) and np.array_equal(a.qp_iters(), b.qp_iters())
        assert np.array_equal(a.cost(), b.cost())

    rollout([a, b], world, 14, check)
    a.close(); b.close()
