"""Neighbour search (reference: the comm-range filter of MultiSyncSimulator::broadcastMsgs,
src/multi_sync_simulator.cpp:481-503): the uniform-grid path (k_nbr_bin + k_nbr_search), the all-pairs path
(k_neighbours) and a numpy restatement must give the same lists -- ascending index, Chebyshev distance on float
differences, first K kept on overflow, own mission only."""
import os
import subprocess
import sys

import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi, missions

pytestmark = pytest.mark.gpu


def reference_lists(pos, group, R, K):
    N = len(pos)
    idx = np.zeros((N, K), np.int32)
    cnt = np.zeros(N, np.int32)
    over = np.zeros(N, bool)
    p = pos.astype(np.float32)
    for a in range(N):
        d = np.abs(p - p[a]).astype(np.float32).max(axis=1).astype(np.float64)
        ok = (group == group[a]) & ~(d > R)
        ok[a] = False
        j = np.nonzero(ok)[0]
        over[a] = len(j) > K
        j = j[:K]
        idx[a, :len(j)] = j
        cnt[a] = len(j)
    return idx, cnt, over


def planner_lists(lib, cfg, m, K, pos, group=None):
    pl = capi.SwarmPlanner(cfg, m, max_nbr=K, lib=lib)
    if group is not None:
        pl.set_groups(group)
    rec = pl.get_records()
    o = cfg.M * (cfg.n + 1) * 3
    rec[:, o:o + 3] = pos
    pl.set_records(0, rec)
    pl.run_stages(capi.STAGE_NBR)
    idx, cnt = pl.neighbours()
    st = pl.status()
    pl.close()
    return idx, cnt, (st & capi.NBR_OVERFLOW) != 0


def check(lib, cfg, m, K, pos, group=None):
    g = np.zeros(len(pos), np.int32) if group is None else group
    ridx, rcnt, rover = reference_lists(pos, g, cfg.comm_range, K)
    idx, cnt, over = planner_lists(lib, cfg, m, K, pos, group)
    assert np.array_equal(cnt, rcnt)
    valid = np.arange(K)[None, :] < rcnt[:, None]
    assert np.array_equal(np.where(valid, idx, 0), np.where(valid, ridx, 0))
    assert np.array_equal(over, rover)


def test_grid_search_matches_reference_on_the_forest(cuda_lib):
    cfg = missions.PlannerConfig.forest3d()
    m = missions.synthetic_forest(n_agents=1024, half_extent=16.0, seed=5)
    rng = np.random.default_rng(0)
    pos = m.start.astype(np.float32) + rng.uniform(-0.4, 0.4, m.start.shape).astype(np.float32)
    check(cuda_lib, cfg, m, 96, pos)


def test_grid_search_edge_cases(cuda_lib):
    """Agents exactly one range apart (boundary of the test and of the cells), agents outside the world box, a
    crowd larger than the candidate buffer (falls back to the plain scan), a capacity overflow."""
    cfg = missions.PlannerConfig.forest3d()
    m = missions.synthetic_forest(n_agents=512, half_extent=12.0, seed=7)
    R = np.float32(cfg.comm_range)
    pos = m.start.astype(np.float32).copy()
    pos[0] = (0.0, 0.0, 1.0); pos[1] = (R, 0.0, 1.0); pos[2] = (np.nextafter(R, np.float32(10)), 0.0, 1.0)
    pos[3] = (-R, R, 1.0); pos[4] = (np.float32(3.0) * R, 0.0, 1.0); pos[5] = (np.float32(2.0) * R, 0.0, 1.0)
    pos[6] = (-12.7, 5.0, 1.0); pos[7] = (12.9, 5.0, 1.0); pos[8] = (40.0, 40.0, 1.0); pos[9] = (-40.0, 40.0, 3.5)
    pos[100:400] = np.float32((6.0, -6.0, 1.0)) + np.random.default_rng(1).uniform(-0.5, 0.5, (300, 3)).astype(np.float32)
    check(cuda_lib, cfg, m, 96, pos)          # the crowd of 300 overflows K = 96 and the 256-entry buffer
    check(cuda_lib, cfg, m, 8, pos)


def test_grid_search_with_mission_groups(cuda_lib):
    cfg = missions.PlannerConfig.forest3d()
    m = missions.synthetic_forest(n_agents=256, half_extent=6.0, seed=3)
    group = (np.arange(256) % 4).astype(np.int32)
    check(cuda_lib, cfg, m, 64, m.start.astype(np.float32), group)


def test_all_pairs_path_agrees(cuda_lib):
    """DLSC_NBR_GRID=0 (the all-pairs kernel) in a subprocess: same lists as the grid path."""
    code = ("import sys; sys.path[:0] = [%r, %r]\n"
            "import numpy as np\n"
            "from dlsc_gc_planner_b200 import capi, missions\n"
            "import test_gpu_neighbours as t\n"
            "cfg = missions.PlannerConfig.forest3d(); m = missions.synthetic_forest(n_agents=512, half_extent=10.0, seed=9)\n"
            "t.check(capi.load_library(), cfg, m, 96, m.start.astype(np.float32))\n"
            "g = (np.arange(512) // 64).astype(np.int32)\n"
            "t.check(capi.load_library(), cfg, m, 96, m.start.astype(np.float32), g)\n"
            "print('ok')\n") % (_parity.ROOT, os.path.join(_parity.ROOT, "tests"))
    env = dict(os.environ, DLSC_NBR_GRID="0")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
