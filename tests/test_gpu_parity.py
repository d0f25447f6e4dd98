"""GPU tier (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same seeded
inputs.  Bit-exact for the float32/float64 geometry stages, QP within the north_star tolerances
(objective 1e-5 relative, violation <= 1e-6 m).  Nothing here reads /root/reference."""
import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi, missions
from test_hostsim_parity import check_worst

pytestmark = pytest.mark.gpu


def make_pair(cuda_lib, cfg, m, K, qp_solver=0):
    sw = _parity.make_oracle(cfg, m, K, n_threads=8)
    pl = capi.SwarmPlanner(cfg, m, max_nbr=K, lib=cuda_lib, qp_solver=qp_solver)
    if cfg.use_sfc:
        pl.set_edt(sw.edt.dist, sw.edt.obst, sw.edt.dims, sw.edt.min_key, sw.edt.res)
    return sw, pl


# qp_solver 2 / 3: the warp-per-agent first scan (k_qp_fast + seeded k_qp_gi) forced on / off -- the default picks by block
# size and would take the second for every mission of this size
@pytest.mark.parametrize("qp_solver", [2, 3])
@pytest.mark.parametrize("name,steps,n", [("empty10", 30, 10), ("maze10", 60, 10), ("forest10", 40, 10),
                                           ("empty70", 12, 70), ("empty50", 8, 50)])
def test_lockstep_parity_reference_missions(cuda_lib, name, steps, n, qp_solver):
    cfg, m = _parity.load_case(name)
    m = _parity.subset(m, n)
    sw, pl = make_pair(cuda_lib, cfg, m, n - 1, qp_solver)
    w = _parity.run_lockstep(pl, sw, m, steps, _parity.default_waypoints(cfg, m))
    check_worst(w)
    assert pl.launch_count() > 0
    pl.close()


@pytest.mark.parametrize("name", ["empty10", "forest10", "maze10"])
def test_lockstep_parity_more_missions(cuda_lib, name):
    """Missions #4, #9, #15, #22, #30 of the family (world k for mission k), 10 lock-step steps each."""
    worst = {}
    for index in (4, 9, 15, 22, 30):
        cfg, m = _parity.load_case(name, index)
        sw, pl = make_pair(cuda_lib, cfg, m, m.n_agents - 1)
        _parity.merge_max(worst, _parity.run_lockstep(pl, sw, m, 10, _parity.default_waypoints(cfg, m)))
        pl.close()
    check_worst(worst)


def test_synthetic_forest_256(cuda_lib):
    """A 256-agent cut of the synthetic forest (BASELINE config 4 shape: M=10, 3-D, SFC, range 3)."""
    cfg = missions.PlannerConfig.forest3d()
    m = missions.synthetic_forest(n_agents=256, half_extent=8.0, seed=11)
    for qp_solver in (2, 0):
        sw, pl = make_pair(cuda_lib, cfg, m, 64, qp_solver)
        w = _parity.run_lockstep(pl, sw, m, 12, _parity.default_waypoints(cfg, m))
        check_worst(w)
        pl.close()


def test_neighbour_overflow(cuda_lib):
    cfg, m = _parity.load_case("empty10")
    sw, pl = make_pair(cuda_lib, cfg, m, 4)
    _parity.force_state(pl, sw)
    sw.step(); pl.plan()
    d = _parity.compare_step(pl, sw)
    assert d["nbr_cnt"] == 0 and d["nbr_idx"] == 0 and d["lsc_d"] == 0
    assert (pl.status() & capi.NBR_OVERFLOW).all()


def test_disturbed_agent_resets(cuda_lib):
    cfg, m = _parity.load_case("forest10")
    sw, pl = make_pair(cuda_lib, cfg, m, 9)
    wf = _parity.default_waypoints(cfg, m)
    for step in range(8):
        sw.waypoint = wf(sw)
        sw.disturbed[:] = 0
        if step == 6:
            sw.disturbed[[1, 4]] = 1
            sw.pos[1] += np.float32(0.7)
        _parity.force_state(pl, sw)
        sw.step(); pl.plan()
        check_worst(_parity.compare_step(pl, sw))
        sw.advance()


def test_state_step_bit_exact(cuda_lib, oracle):
    cfg, m = _parity.load_case("forest10")
    sw, pl = make_pair(cuda_lib, cfg, m, 9)
    wf = _parity.default_waypoints(cfg, m)
    for _ in range(5):
        sw.waypoint = wf(sw)
        _parity.force_state(pl, sw)
        sw.step(); pl.plan()
        traj = pl.traj()
        pl.advance()
        pos, vel, acc = pl.state()
        for a in range(m.n_agents):
            st = oracle.state_at(sw.p, traj[a], cfg.dt)
            assert (st[0] == pos[a]).all() and (st[1] == vel[a]).all() and (st[2] == acc[a]).all()
        sw.advance()


def test_free_running_rollout_and_determinism(cuda_lib):
    """Device-chained rollout (plan -> advance, no host state) twice: identical bits both times, close to
    the oracle's own rollout, collision free (safety ratio >= 1) and every QP feasible to 1e-6 m."""
    cfg, m = _parity.load_case("empty10")
    sw = _parity.make_oracle(cfg, m, 9)
    wf = _parity.default_waypoints(cfg, m)
    runs = []
    wps = []
    for rep in range(2):
        pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=cuda_lib)
        trajs = []
        for step in range(40):
            if rep == 0:
                sw.waypoint = wf(sw)
                wps.append(sw.waypoint.copy())
                sw.step()
            pl.set_agents(waypoint=wps[step])
            pl.plan()
            assert pl.violation().max() <= 1e-6
            assert (pl.status() & capi.FAIL_MASK).max() == 0
            trajs.append(pl.traj())
            pl.advance()
            if rep == 0:
                sw.advance()
                pos, _, _ = pl.state()
                assert np.max(np.abs(pos - sw.pos)) < 5e-5
                for i in range(m.n_agents):
                    d = pos[i + 1:] - pos[i]
                    d[:, 2] /= 2.0
                    if len(d):
                        assert np.min(np.linalg.norm(d, axis=1)) / 0.3 >= 1.0 - 1e-4
        runs.append(np.array(trajs))
        pl.close()
    assert np.array_equal(runs[0], runs[1])


def test_sharded_contexts_match_single_context(cuda_lib):
    """Two contexts on one GPU, each owning half of the agents (the multi-GPU layout), exchanging records
    through host memory, reproduce the single-context result bit for bit."""
    cfg, m = _parity.load_case("forest10")
    p = _parity.oracle_params(cfg, m)
    from oracle import oracle_py as O
    edt = O.edt_build(p, m.boxes)
    whole = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=cuda_lib)
    halves = [capi.SwarmPlanner(cfg, m, max_nbr=9, begin=b, n_local=5, lib=cuda_lib) for b in (0, 5)]
    for pl in [whole] + halves:
        pl.set_edt(edt.dist, edt.obst, edt.dims, edt.min_key, edt.res)
    wp = m.start.copy()
    for h in halves:                                  # initial all-gather of the reset records
        for o in halves:
            if o is not h:
                h.set_records(o.begin, o.get_records(o.begin, o.NL))
    for step in range(10):
        wp[:, 0] += np.float32(0.1) * np.sign(m.goal[:, 0] - wp[:, 0])
        whole.set_agents(waypoint=wp)
        whole.plan(); whole.advance()
        for h in halves:
            h.set_agents(waypoint=wp[h.begin:h.begin + h.NL])
            h.plan(); h.advance()
        for h in halves:                              # "all-gather"
            for o in halves:
                if o is not h:
                    h.set_records(o.begin, o.get_records(o.begin, o.NL))
        t = whole.traj()
        assert np.array_equal(t[:5], halves[0].traj()) and np.array_equal(t[5:], halves[1].traj())
        assert np.array_equal(whole.get_records(), halves[0].get_records())


def test_bound_host_trajectories_equal_the_device_copy(cuda_lib):
    """dlsc_bind_traj_host: the mapped host buffer holds exactly dlsc_get_traj's result after every step (pinned through
    torch, and a pageable numpy array that the library pins itself), also on the CUDA-graph path and after unbinding."""
    import torch
    cfg, m = _parity.load_case("forest10")
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=cuda_lib)
    pl.build_edt(m.boxes)
    shape = (m.n_agents, cfg.M, cfg.n + 1, 3)
    pinned = torch.empty(shape, dtype=torch.float32).pin_memory().numpy()
    pageable = np.zeros(shape, np.float32)
    wf = _parity.default_waypoints(cfg, m)
    wp = m.start.copy()
    for step in range(12):
        buf = pinned if step < 8 else pageable
        if step in (0, 8):
            pl.bind_traj_host(buf)
        wp[:, 0] += np.float32(0.1) * np.sign(m.goal[:, 0] - wp[:, 0])
        pl.set_agents(waypoint=wp)
        buf[...] = -7.0
        pl.plan(); pl.sync()
        assert np.array_equal(buf, pl.traj()), step
        pl.advance()
    pl.bind_traj_host(None)
    pageable[...] = -7.0
    pl.plan(); pl.sync()
    assert np.all(pageable == -7.0)
    pl.close()
