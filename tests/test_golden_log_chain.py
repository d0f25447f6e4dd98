"""The reference's only recorded run (maze10_dense #1, CPLEX, 34 s; log/result_1742185870.978562_DLSCGC_10agents.csv)
against the composed path LSC + SFC + goal + QP + state step, chained.  The waypoints the reference's PIBT layer issued
are not logged; they were recovered from the log itself (tests/golden/infer_waypoints.py: every wrong lattice move
misses the logged states by >= 0.07 m, the right one reproduces them) and are committed as
tests/golden/inferred_waypoints.npz.  With them a FREE-RUNNING rollout -- no state is ever reset from the log -- must
stay on the logged CPLEX trajectory at both save times of every step:
  steps 0-19   : position to the printed digits + 1e-6, velocity 3e-5, acceleration 6e-4
  steps 20-136 : 100 x that (position 1e-4 m: the closed-loop drift between our QP solvers and CPLEX)
137 chained replans x 10 agents = 27.4 s of the mission."""
import os

import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi

GOLD = os.path.join(_parity.ROOT, "tests", "golden")
TIGHT_STEPS = 20


def fixtures():
    z = np.load(os.path.join(GOLD, "golden_log_full.npz"))
    w = np.load(os.path.join(GOLD, "inferred_waypoints.npz"))
    return z["state"], w["waypoints"], float(w["scale"])


def tolerance(ref, scale):
    return scale * np.array([1e-6] * 3 + [3e-5] * 3 + [6e-4] * 3) + 10.0 ** (np.floor(np.log10(np.maximum(np.abs(ref), 1e-30))) - 5)


def check_step(oracle, p, traj, step, state, scale):
    for k, t in ((2 * step + 1, 0.1), (2 * step + 2, 0.2)):
        for a in range(traj.shape[0]):
            s = oracle.state_at(p, traj[a], t).reshape(9)
            assert np.all(np.abs(s - state[k, a]) <= tolerance(state[k, a], scale)), (step, k, a, s, state[k, a])


def test_inferred_waypoints_are_lattice_moves():
    _, wps, scale = fixtures()
    assert wps.shape == (137, 10, 3) and scale == 100.0
    assert np.allclose(wps[..., :2] * 2, np.round(wps[..., :2] * 2))                   # lattice nodes
    assert np.abs(np.diff(wps[..., :2], axis=0)).sum(axis=2).max() <= 0.5 + 1e-6        # one 4-connected move per replan
    assert np.all(wps[..., 2] == 1.0)


def test_oracle_rollout_stays_on_the_cplex_log(oracle):
    cfg, m = _parity.load_case("maze10")
    sw = _parity.make_oracle(cfg, m, 9, n_threads=os.cpu_count() or 1)
    state, wps, scale = fixtures()
    for step in range(len(wps)):
        sw.waypoint = wps[step]
        st = sw.step()
        assert (st & ~16).max() == 0
        check_step(oracle, sw.p, sw.traj, step, state, 1.0 if step < TIGHT_STEPS else scale)
        sw.advance()


@pytest.mark.gpu
def test_gpu_rollout_stays_on_the_cplex_log(cuda_lib, oracle):
    """The same chain on the CUDA path: plan -> advance on the device for 137 steps, distance grid built on the device."""
    cfg, m = _parity.load_case("maze10")
    p = _parity.oracle_params(cfg, m)
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=cuda_lib)
    pl.build_edt(m.boxes)
    state, wps, scale = fixtures()
    for step in range(len(wps)):
        pl.set_agents(waypoint=wps[step])
        pl.plan()
        assert (pl.status() & capi.FAIL_MASK).max() == 0
        check_step(oracle, p, pl.traj(), step, state, 1.0 if step < TIGHT_STEPS else scale)
        pl.advance()
    pl.close()


@pytest.mark.gpu
def test_gpu_rollout_with_generated_waypoints(cuda_lib, oracle):
    """Closed loop entirely through the C ABI: waypoints from dlsc_wp_step (comm-range groups + PIBT + update rules), replans
    and state steps on the GPU -- no recorded input at all besides the mission -- stay on the reference's CPLEX log for
    all 137 steps, and the generated waypoints equal the ones recovered from that log."""
    cfg, m = _parity.load_case("maze10")
    p = _parity.oracle_params(cfg, m)
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=cuda_lib)
    pl.build_edt(m.boxes)
    dist, obst, dims, mk = pl.get_edt()
    wp = capi.WaypointProvider(cfg, m, lib=cuda_lib, edt=(dist, obst, dims, mk, cfg.world_res))
    state, wps, scale = fixtures()
    wcur = pl.start.copy()
    traj = None
    for step in range(len(wps)):
        pos, _, _ = pl.state()
        wcur = wp.step(pos, pl.goal() if step else pl.start, traj, wcur)
        assert np.array_equal(wcur, wps[step]), step
        pl.set_agents(waypoint=wcur)
        pl.plan()
        assert (pl.status() & capi.FAIL_MASK).max() == 0
        traj = pl.traj()
        check_step(oracle, p, traj, step, state, 1.0 if step < TIGHT_STEPS else scale)
        pl.advance()
    pl.close(); wp.close()
