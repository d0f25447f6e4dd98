"""Result log writer against the head of the reference's own recorded run (tests/golden/result_head.csv, the first
rows of log/result_1742185870.978562_DLSCGC_10agents.csv): parsing it and writing it back must give the same text."""
import os

import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi, resultlog

HEAD = os.path.join(_parity.ROOT, "tests", "golden", "result_head.csv")


def test_round_trip_reproduces_the_reference_text(tmp_path):
    t, pos, vel, acc, ptime = resultlog.read(HEAD)
    assert pos.shape == (3, 10, 3)
    out = tmp_path / "result.csv"
    log = resultlog.ResultLog(str(out), 10)
    for k in range(len(t)):
        log.record(t[k], pos[k], vel[k], acc[k], ptime[k])
    log.close()
    assert open(out).read() == open(HEAD).read()


def test_header_and_number_format():
    assert resultlog.header(2) == resultlog.AGENT_COLUMNS + "," + resultlog.AGENT_COLUMNS
    row = resultlog.format_row(0.2, np.array([[-1.0, 2.5, 1.0]], np.float32), np.array([[0.123456789, -0.0, 1e-7]], np.float32),
                               np.zeros((1, 3), np.float32), [0.0068735])
    assert row == "0,0.2,-1,2.5,1,0.123457,-0,1e-07,0,0,0,0.0068735"


def test_oracle_first_replan_writes_the_reference_log(oracle):
    """The oracle's first replan of maze10_dense #1, sampled at the save times and formatted by the writer, is the
    head of the reference's log character for character (every state column; planning_time is wall time, copied)."""
    cfg, m = _parity.load_case("maze10")
    sw = _parity.make_oracle(cfg, m, 9)
    wp = sw.pos.copy()
    wp[:5, 0] += 0.5
    wp[5:, 0] -= 0.5
    sw.waypoint = wp
    t_ref, _, _, _, ptime_ref = resultlog.read(HEAD)
    rows = [resultlog.header(10), resultlog.format_row(0.0, sw.pos, sw.vel, sw.acc, ptime_ref[0])]
    sw.step()
    for k in (1, 2):
        st = np.array([oracle.state_at(sw.p, sw.traj[a], float(t_ref[k])) for a in range(10)])
        rows.append(resultlog.format_row(t_ref[k], st[:, 0], st[:, 1], st[:, 2], ptime_ref[k]))
    assert "\n".join(rows) + "\n" == open(HEAD).read()


@pytest.mark.gpu
def test_gpu_first_replan_writes_the_reference_log(cuda_lib, oracle, tmp_path):
    """maze10_dense #1: the first replan on the GPU (PIBT's first waypoints: one lattice step along x, as in
    test_oracle_pinning), its trajectory sampled at the log's save times t = 0, 0.1, 0.2 and written through the log
    writer: the same text as the head of the reference's own log (planning_time column copied, it is wall time)."""
    cfg, m = _parity.load_case("maze10")
    sw = _parity.make_oracle(cfg, m, 9, n_threads=8)
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=cuda_lib)
    pl.set_edt(sw.edt.dist, sw.edt.obst, sw.edt.dims, sw.edt.min_key, sw.edt.res)
    t_ref, _, _, _, ptime_ref = resultlog.read(HEAD)
    wp = m.start.astype(np.float32).copy()
    wp[:5, 0] += 0.5
    wp[5:, 0] -= 0.5
    out = tmp_path / "gpu.csv"
    log = resultlog.ResultLog(str(out), m.n_agents)
    pos, vel, acc = pl.state()
    log.record(0.0, pos, vel, acc, ptime_ref[0])
    pl.set_agents(waypoint=wp)
    pl.plan()
    assert (pl.status() & capi.FAIL_MASK).max() == 0
    traj = pl.traj()
    for k in (1, 2):
        st = np.array([oracle.state_at(sw.p, traj[a], float(t_ref[k])) for a in range(m.n_agents)])     # [N][3][3]
        log.record(t_ref[k], st[:, 0], st[:, 1], st[:, 2], ptime_ref[k])
    log.close()
    pl.close()
    assert open(out).read() == open(HEAD).read()


def test_summary_row_reproduces_the_reference_summary_file(tmp_path):
    """log/summary_DLSCGC_10agents.csv (fixture tests/golden/summary_ref.csv): the columns, the formatting, and the two
    derived quantities that can be recomputed from the reference's own result log -- total flight distance and the minimum
    inter-agent safety ratio over all recorded times -- come out as the reference printed them."""
    ref = open(os.path.join(_parity.ROOT, "tests", "golden", "summary_ref.csv")).read().strip().split("\n")
    assert ref[0] == resultlog.SUMMARY_COLUMNS
    want = ref[1].split(",")
    z = np.load(os.path.join(_parity.ROOT, "tests", "golden", "golden_log_full.npz"))
    pos = z["state"][:, :, 0:3].astype(np.float32)
    dist = resultlog.total_flight_distance(pos)
    safety = resultlog.safety_ratio_agents(pos, np.full(10, 0.15), np.full(10, 2.0))
    assert resultlog._g(dist) == want[2] == "134.096"
    assert resultlog._g(safety) == want[3] == "1.00058"
    f = lambda lo, hi: [float(x) for x in want[lo:hi]]
    row = resultlog.format_summary(want[0], float(want[1]), dist, safety, 1e9, f(5, 8), f(8, 11), f(11, 17), want[17], want[18],
                                   communication_range=3, world_dimension=2, M=10, dt=0.2)
    assert row == ref[1]
    p = tmp_path / "summary.csv"
    resultlog.append_summary(str(p), row); resultlog.append_summary(str(p), row)
    assert open(p).read() == ref[0] + "\n" + ref[1] + "\n" + ref[1] + "\n"


def test_result_log_with_obstacle_columns(tmp_path):
    """Missions with dynamic obstacles: `obs_id,t,px,py,pz,size` per obstacle after the agents (reference
    src/multi_sync_simulator.cpp:755-763, 829-843), read back the way MultiSyncReplayer::readCSVFile counts the columns."""
    import numpy as np
    from dlsc_gc_planner_b200 import resultlog
    p = str(tmp_path / "r.csv")
    log = resultlog.ResultLog(p, 2, 2)
    pos = np.array([[1, 2, 3], [-0.5, 0.25, 1]], np.float32)
    z = np.zeros((2, 3), np.float32)
    log.record(0.1, pos, z, z, [0.001, 0.002], np.array([[2, 0, 1], [0, -2, 1]], np.float32), [0.3, 0.25])
    log.record(0.2, pos + 1, z, z, None, np.array([[1.9, 0.5, 1], [0.5, -1.9, 1]], np.float32), [0.3, 0.25])
    log.close()
    lines = open(p).read().strip().split("\n")
    assert lines[0] == ",".join(["id,t,px,py,pz,vx,vy,vz,ax,ay,az,planning_time"] * 2 + ["obs_id,t,px,py,pz,size"] * 2)
    assert lines[1].endswith(",0,0.1,2,0,1,0.3,1,0.1,0,-2,1,0.25")
    t, rp, rv, ra, pt, op, orad = resultlog.read(p, with_obstacles=True)
    assert np.allclose(t, [0.1, 0.2]) and np.array_equal(rp[1], pos + 1) and np.allclose(pt[0], [0.001, 0.002])
    assert np.allclose(op[1], [[1.9, 0.5, 1], [0.5, -1.9, 1]]) and np.allclose(orad, 0.3 * np.array([[1, 0.25 / 0.3]] * 2))
    assert resultlog.read(p)[1].shape == (2, 2, 3)
