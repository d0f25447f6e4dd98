"""The GJK core the LSC kernel runs (gjk::hull_origin, dlsc_math.cuh) against the REFERENCE'S OWN openGJK object
code (known-answer vectors generated from oracle/_ref = /root/reference/src/openGJK/openGJK.cpp compiled unmodified):

  tests/golden/gjk_ref.npz     4000 hulls (generic, near-collinear, origin-containing)
  tests/golden/gjk_leaves.npz  ~3000 hulls selected so that every reachable leaf of the S1D/S2D/S3D decision tree
                               (openGJK.cpp:243-631) is taken by >= 50 of them (tests/golden/make_gjk_leaves.py)

Two routes, both bit-exact (`==`):
  * dlsc_gjk_batch: the per-kernel entry point; compares the raw fp64 witness vector, the final simplex size and
    reports which leaves were taken,
  * the LSC stage itself (k_lsc through dlsc_run_stages): hull = one segment of the agent's initial trajectory,
    neighbour's predicted trajectory = 0, downwash 1 (exact), result read back with dlsc_get_lsc and compared with
    the normal / margins that follow from the reference witness vector (traj_planner.cpp:1118, 630-637).
CPU tier: the same source executed by tests/hostsim.  GPU tier (-m gpu): libdlsc_b200.so on the B200.
"""
import os

import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi, missions

GOLD = os.path.join(_parity.ROOT, "tests", "golden")
N_LEAVES = 57
# leaves no input can reach: 29 is logically dead in the reference too (n_edge == 2 implies ei or ej,
# openGJK.cpp:497-499); 31 / 33 / 49 need sign patterns of the half-space tests that no four points produce
# (none in 5 M hulls of ten geometric families, exact-tie lattices included); 53 follows from 49; 54 is the
# 25-iteration cap (never reached: at most 7 iterations observed).
UNREACHABLE = {29, 31, 33, 49, 53, 54}
MIN_HITS = 50


def _vectors():
    a = np.load(os.path.join(GOLD, "gjk_ref.npz"))
    b = np.load(os.path.join(GOLD, "gjk_leaves.npz"))
    pts = np.concatenate([a["pts"], b["pts"].astype(np.float64)])
    v = np.concatenate([a["v"], b["v"]])
    d = np.concatenate([a["d"], b["d"]])
    sn = np.concatenate([a["simplex"], b["simplex"]])
    return pts, v, d, sn, b["leaves"]


def _planner(lib, n_agents, M=10):
    cfg = missions.PlannerConfig.forest3d()
    assert cfg.M == M
    z = np.zeros((n_agents, 3), np.float32)
    one = np.ones(n_agents)
    m = missions.Mission(np.array([-8, -8, 0], np.float32), np.array([8, 8, 4], np.float32), z, z.copy(), 0.15 * one,
                         1.0 * one, one, 2 * one, one, np.zeros((0, 6), np.float32))
    return capi.SwarmPlanner(cfg, m, max_nbr=1, lib=lib)


def _f32(x):
    return np.asarray(x, np.float32)


def _expected_lsc(pts, v_ref, collision_dist):
    """normal / d of traj_planner.cpp:1118, 630-637 from the reference witness vector, float32 op by op
    (octomath::Vector3: float storage, dot / norm_sq evaluated in float then widened)."""
    cp = _f32(v_ref)                                             # geometry.hpp:302 double -> float
    nsq = (cp[:, 0] * cp[:, 0] + cp[:, 1] * cp[:, 1]) + cp[:, 2] * cp[:, 2]
    ln = np.sqrt(nsq.astype(np.float64))
    f = _f32(ln)
    nt = cp.copy()
    nz = ln > 0
    nt[nz] = cp[nz] / f[nz, None]
    rel = _f32(pts)                                              # [n][6][3]
    dot = (rel[:, :, 0] * nt[:, None, 0] + rel[:, :, 1] * nt[:, None, 1]) + rel[:, :, 2] * nt[:, None, 2]
    d = 0.5 * (collision_dist + dot.astype(np.float64))
    return nt, d


def _check_batch(pl, need_leaves=True):
    pts, v_ref, d_ref, sn_ref, leaves_fixture = _vectors()
    v, it, sn, lv = pl.gjk_batch(pts)
    assert int((v != v_ref).any(axis=1).sum()) == 0              # witness vector, all 64 bits of each component
    assert np.array_equal(np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]), d_ref)   # openGJK.cpp:779
    assert np.array_equal(sn, sn_ref)                            # final simplex size
    bits = ((lv[:, None] >> np.arange(N_LEAVES, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
    hist = bits.sum(axis=0)
    if need_leaves:
        for b in range(N_LEAVES):
            if b in UNREACHABLE:
                assert hist[b] == 0, (b, hist[b])
            else:
                assert hist[b] >= MIN_HITS, (b, int(hist[b]))
        # the leaves recorded when the fixture was generated (hostsim): same path on this device / build
        assert np.array_equal(lv[-len(leaves_fixture):], leaves_fixture)
    assert it.max() <= 8
    return hist


def _check_lsc_stage(lib):
    pts, v_ref, _, _, _ = _vectors()
    M = 10
    per = M - 1
    n_agents = (len(pts) + per - 1) // per
    pad = n_agents * per - len(pts)
    hull = np.concatenate([pts, pts[:pad]]).astype(np.float32).reshape(n_agents, per, 6, 3)
    vr = np.concatenate([v_ref, v_ref[:pad]]).reshape(n_agents, per, 3)
    pl = _planner(lib, n_agents, M)
    init = np.zeros((n_agents, M, 6, 3), np.float32)
    init[:, :per] = hull
    init[:, per] = hull[:, 0]
    pl.set_init_traj(init)
    pl.set_pred_traj(np.zeros((n_agents, M, 6, 3), np.float32))
    idx = ((np.arange(n_agents) + 1) % n_agents).astype(np.int32).reshape(n_agents, 1)
    pl.set_neighbours(idx, np.ones(n_agents, np.int32))
    pl.run_stages(capi.STAGE_LSC)
    normal, _, d = pl.lsc(with_anchor=False)
    cd = float(np.float32(0.15)) + 0.15                          # traj_planner.cpp:605, radius through float (agent_manager.cpp:256)
    nt, dd = _expected_lsc(hull.reshape(-1, 6, 3), vr.reshape(-1, 3), cd)
    got_n = normal[:, 0, :per].reshape(-1, 3)
    got_d = d[:, 0, :per].reshape(-1, 6)
    assert int((got_n != nt).any(axis=1).sum()) == 0
    assert int((got_d != dd).any(axis=1).sum()) == 0
    assert pl.counters()["gjk_iters"] > 2 * len(pts)
    pl.close()


def test_fixture_covers_every_reachable_leaf():
    z = np.load(os.path.join(GOLD, "gjk_leaves.npz"))
    bits = ((z["leaves"][:, None] >> np.arange(N_LEAVES, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
    hist = bits.sum(axis=0)
    assert all(hist[b] >= MIN_HITS for b in range(N_LEAVES) if b not in UNREACHABLE), hist.tolist()


def test_gjk_core_matches_reference_object_code_hostsim(hostsim):
    pl = _planner(hostsim, 2)
    _check_batch(pl)
    pl.close()


def test_lsc_stage_matches_reference_object_code_hostsim(hostsim):
    _check_lsc_stage(hostsim)


@pytest.mark.gpu
def test_gjk_core_matches_reference_object_code_gpu(cuda_lib):
    pl = _planner(cuda_lib, 2)
    hist = _check_batch(pl)
    print("leaf histogram on the device:", hist.tolist())
    pl.close()


@pytest.mark.gpu
def test_lsc_kernel_matches_reference_object_code_gpu(cuda_lib):
    _check_lsc_stage(cuda_lib)
