"""The reference's recorded run (maze10_dense #1, CPLEX; first 4 s = 20 chained replans of 10 agents) against the
composed path LSC + SFC + goal + QP + state step.  The waypoints the reference's PIBT layer issued are not logged;
they were recovered from the log itself (tests/golden/infer_waypoints.py: every wrong lattice move misses the logged
states by >= 0.07, the right one reproduces them) and are committed as tests/golden/inferred_waypoints.npz.  With
them, a free-running rollout -- no state is ever reset from the log -- must stay on the logged CPLEX trajectory:
position to the printed digits + 1e-6, velocity 3e-5, acceleration 6e-4, at both save times of all 20 steps."""
import os

import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi, resultlog

GOLD = os.path.join(_parity.ROOT, "tests", "golden")


def logged_states():
    t, pos, vel, acc, _ = resultlog.read(os.path.join(GOLD, "result_rows_0_4s.csv"))
    return t, np.concatenate([pos, vel, acc], axis=2).astype(np.float64)


def tolerance(ref, scale=1.0):
    return scale * np.array([1e-6] * 3 + [3e-5] * 3 + [6e-4] * 3) + 10.0 ** (np.floor(np.log10(np.maximum(np.abs(ref), 1e-30))) - 5)


def check_step(oracle, p, traj, step, state, scale=1.0):
    worst = 0.0
    for k, t in ((2 * step + 1, 0.1), (2 * step + 2, 0.2)):
        for a in range(traj.shape[0]):
            s = oracle.state_at(p, traj[a], t).reshape(9)
            err = np.abs(s - state[k, a]) / tolerance(state[k, a], scale)
            assert err.max() <= 1.0, (step, k, a, s, state[k, a])
            worst = max(worst, float(err.max()))
    return worst


def test_oracle_rollout_stays_on_the_cplex_log(oracle):
    cfg, m = _parity.load_case("maze10")
    sw = _parity.make_oracle(cfg, m, 9, n_threads=os.cpu_count() or 1)
    wps = np.load(os.path.join(GOLD, "inferred_waypoints.npz"))["waypoints"]
    _, state = logged_states()
    assert wps.shape == (20, 10, 3)
    assert np.allclose(wps[..., :2] * 2, np.round(wps[..., :2] * 2))                   # lattice nodes
    assert np.abs(np.diff(wps[..., :2], axis=0)).sum(axis=2).max() <= 0.5 + 1e-6        # one 4-connected move per replan
    for step in range(20):
        sw.waypoint = wps[step]
        st = sw.step()
        assert (st & ~16).max() == 0
        check_step(oracle, sw.p, sw.traj, step, state)
        sw.advance()


@pytest.mark.gpu
def test_gpu_rollout_stays_on_the_cplex_log(cuda_lib, oracle):
    """Same chain on the CUDA path: plan -> advance on the device for 20 steps, grid built on the device."""
    cfg, m = _parity.load_case("maze10")
    p = _parity.oracle_params(cfg, m)
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=cuda_lib)
    pl.build_edt(m.boxes)
    wps = np.load(os.path.join(GOLD, "inferred_waypoints.npz"))["waypoints"]
    _, state = logged_states()
    for step in range(20):
        pl.set_agents(waypoint=wps[step])
        pl.plan()
        assert (pl.status() & capi.FAIL_MASK).max() == 0
        check_step(oracle, p, pl.traj(), step, state)
        pl.advance()
    pl.close()
