"""Failure paths of the replan: the kernels and the oracle must agree on WHEN a replan fails and on what
is returned then.  The reference knows one QP failure (any exception -> desired_traj := initial_traj,
src/traj_planner.cpp:749-777), the goal LP's QPFAILED (src/goal_optimizer.cpp:122,132) and "Invalid initial SFC"
(src/collision_constraints.cpp:445-447); the status bits QP_MAXITER / QP_NUMERIC both mean "QP failed".

  * infeasible QP            a fast agent 5 cm from the world boundary (control-point bounds cannot be met)
  * infeasible QP, recorded  step 137 of the reference's CPLEX log: the only move of agent 1 compatible with the log
                             makes its QP infeasible at 1e-10 (CPLEX, feasibility tolerance 1e-6, returned a point)
  * GOAL_INFEASIBLE          two agents whose current goals are closer than the collision distance
  * SFC_INIT_FAILED          an agent that starts inside an obstacle
  * active set -> interior point hand-over (DLSC_QP_IPM_USED), forced with qp_active_max = 1

CPU tier: kernel cores through tests/hostsim.  GPU tier (-m gpu): libdlsc_b200.so.
"""
import os

import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi

QP_FAIL = capi.QP_MAXITER | capi.QP_NUMERIC
EXACT = ("init_traj", "pred_traj", "nbr_cnt", "nbr_idx", "lsc_normal", "lsc_d", "lsc_anchor", "sfc", "goal")


def _classes(st):
    """status -> (QP failed, SFC init failed, goal infeasible): the distinctions the reference makes"""
    st = np.asarray(st)
    return (st & QP_FAIL) != 0, (st & capi.SFC_INIT_FAILED) != 0, (st & capi.GOAL_INFEASIBLE) != 0


def _step_and_compare(pl, sw):
    _parity.force_state(pl, sw)
    st_o = sw.step().copy()
    pl.plan()
    st_p = pl.status()
    d = _parity.compare_step(pl, sw)
    for k in EXACT:
        if k in d:
            assert d[k] == 0, (k, d[k])
    for a, b in zip(_classes(st_p), _classes(st_o)):
        assert np.array_equal(a, b), (st_p, st_o)
    # failsafe: every agent whose QP failed flies its initial trajectory, bit for bit, on both sides
    failed = _classes(st_o)[0]
    traj_p, init_p = pl.traj(), pl.init_traj()
    for a in np.flatnonzero(failed):
        assert np.array_equal(traj_p[a], init_p[a])
        assert np.array_equal(sw.traj[a], sw.init_traj[a])
        assert np.array_equal(traj_p[a], sw.traj[a])
    # the others are ordinary replans
    ok = ~failed & ~_classes(st_o)[1]
    if ok.any():
        assert np.abs(pl.traj()[ok] - sw.traj[ok]).max() <= 1e-5
        excess = np.abs(pl.cost() - sw.cost) - _parity.OBJ_REL * np.abs(sw.cost)
        assert excess[ok].max() <= _parity.OBJ_ABS
        assert pl.violation()[ok].max() <= 1e-6
    return st_p, st_o


def _warm(cfg, m, lib, steps, edt=False, **kw):
    sw = _parity.make_oracle(cfg, m, m.n_agents - 1, n_threads=os.cpu_count() or 1)
    pl = capi.SwarmPlanner(cfg, m, max_nbr=m.n_agents - 1, lib=lib, **kw)
    if edt:
        pl.set_edt(sw.edt.dist, sw.edt.obst, sw.edt.dims, sw.edt.min_key, sw.edt.res)
    wf = _parity.default_waypoints(cfg, m)
    for _ in range(steps):
        sw.waypoint = wf(sw)
        sw.step()
        sw.advance()
    sw.waypoint = wf(sw)
    return sw, pl


def _infeasible_qp(lib):
    cfg, m = _parity.load_case("empty10")
    sw, pl = _warm(cfg, m, lib, 3)
    sw.pos[0] = np.float32([m.world_max[0] - 0.05, 0.0, 1.0])      # 1 m/s towards a wall 5 cm away, 2 m/s^2 available
    sw.vel[0] = np.float32([1.0, 0.0, 0.0])
    sw.acc[0] = 0
    st_p, st_o = _step_and_compare(pl, sw)
    assert st_o[0] & QP_FAIL and st_p[0] & QP_FAIL
    assert (_classes(st_o)[0].sum()) >= 1
    pl.close()


def _goal_infeasible(lib):
    cfg, m = _parity.load_case("empty10")
    sw, pl = _warm(cfg, m, lib, 3)
    g = sw.goal_cur[0].copy()
    sw.goal_cur[1] = g + np.float32([0.1, 0, 0])                    # closer than r_i + r_j = 0.3
    sw.waypoint[0] = g + np.float32([0, 0.05, 0])
    st_p, st_o = _step_and_compare(pl, sw)
    assert st_o[0] & capi.GOAL_INFEASIBLE and st_p[0] & capi.GOAL_INFEASIBLE
    pl.close()


def _sfc_init_failed(lib):
    cfg, m = _parity.load_case("forest10")
    sw, pl = _warm(cfg, m, lib, 0, edt=True)
    sw.pos[2] = np.float32([m.boxes[0][0], m.boxes[0][1], 1.0])     # inside the first tree
    st_p, st_o = _step_and_compare(pl, sw)
    assert st_o[2] & capi.SFC_INIT_FAILED and st_p[2] & capi.SFC_INIT_FAILED
    assert not (np.delete(st_o, 2) & capi.SFC_INIT_FAILED).any()
    pl.close()


def _handover(lib):
    """qp_active_max = 1: every agent that needs two simultaneously active rows is handed to the interior point."""
    cfg, m = _parity.load_case("maze10")
    sw = _parity.make_oracle(cfg, m, 9, n_threads=os.cpu_count() or 1)
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=lib, qp_active_max=1)
    ref = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=lib)
    for q in (pl, ref):
        q.set_edt(sw.edt.dist, sw.edt.obst, sw.edt.dims, sw.edt.min_key, sw.edt.res)
    wf = _parity.default_waypoints(cfg, m)
    used = 0
    for _ in range(30):
        sw.waypoint = wf(sw)
        _parity.force_state(ref, sw)
        ref.plan()
        st_p, st_o = _step_and_compare(pl, sw)
        ipm = (st_p & capi.QP_IPM_USED) != 0
        used += int(ipm.sum())
        assert not (ref.status() & capi.QP_IPM_USED).any()
        if ipm.any():                                                # both solvers of the product agree with each other
            assert np.abs(pl.cost()[ipm] - ref.cost()[ipm]).max() <= 1e-5 * np.abs(ref.cost()[ipm]).max() + _parity.OBJ_ABS
            assert np.abs(pl.qp_x()[ipm] - ref.qp_x()[ipm]).max() <= 1e-5
        sw.advance()
    assert used >= 20, used
    pl.close(); ref.close()


def _golden_step_137(lib):
    cfg, m = _parity.load_case("maze10")
    sw = _parity.make_oracle(cfg, m, 9, n_threads=os.cpu_count() or 1)
    wps = np.load(os.path.join(_parity.ROOT, "tests", "golden", "inferred_waypoints.npz"))["waypoints"]
    for step in range(len(wps)):
        sw.waypoint = wps[step]
        sw.step()
        sw.advance()
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=lib)
    pl.set_edt(sw.edt.dist, sw.edt.obst, sw.edt.dims, sw.edt.min_key, sw.edt.res)
    keep = {k: getattr(sw, k).copy() for k in ("pos", "vel", "acc", "goal_cur", "traj", "sfc", "sfc_init", "disturbed")}
    seq = sw.seq
    found = 0
    for dx, dy in ((0.0, 0.0), (0.5, 0.0), (-0.5, 0.0), (0.0, 0.5), (0.0, -0.5)):
        for k, v in keep.items():
            getattr(sw, k)[...] = v
        sw.seq = seq
        w = wps[-1].copy()
        w[1, 0] += dx; w[1, 1] += dy
        sw.waypoint = w
        st_p, st_o = _step_and_compare(pl, sw)
        found += int((st_o[1] & QP_FAIL) != 0)
    assert found >= 1            # the recorded infeasible replan exists, and the kernels fail on exactly the same moves
    pl.close()


CASES = [_infeasible_qp, _goal_infeasible, _sfc_init_failed, _handover, _golden_step_137]


@pytest.mark.parametrize("case", CASES, ids=lambda f: f.__name__.strip("_"))
def test_failure_path_hostsim(hostsim, case):
    case(hostsim)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda f: f.__name__.strip("_"))
def test_failure_path_gpu(cuda_lib, case):
    case(cuda_lib)
