"""Pins the CPU oracle (test infrastructure) against everything the reference offers for this path:
  * the reference's own openGJK object code (known-answer vectors in tests/golden/gjk_ref.npz, and live
    against oracle/_ref when it is built),
  * the reference's only recorded run (log/result_...DLSCGC_10agents.csv): first replan of maze10_dense #1,
  * (the independent QP cross-check -- second formulation + HiGHS + KKT certificate -- lives in test_qp_crosscheck.py).
"""
import numpy as np
import pytest

import _parity


def test_gjk_matches_reference_vectors(oracle):
    z = np.load(_parity.os.path.join(_parity.ROOT, "tests", "golden", "gjk_ref.npz"))
    bad = 0
    for p, v, d in zip(z["pts"], z["v"], z["d"]):
        d2, v2, _, _ = oracle.gjk(p)
        bad += int(d2 != d or (v2 != v).any())
    assert bad == 0


def test_gjk_matches_reference_object_code_live(oracle):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(5)
    for _ in range(3000):
        p = (rng.normal(size=3) * rng.uniform(0, 2) + rng.normal(size=(6, 3)) * rng.uniform(0.01, 1)).astype(np.float32)
        d, v, _, _ = oracle.gjk(p.astype(np.float64))
        d2, v2, _ = oracle.ref_gjk(p.astype(np.float64))
        assert d == d2 and (v == v2).all()


def _sig6(x):
    return float("%.6g" % x)


def test_golden_log_first_replan(oracle):
    """LSC + SFC + goal + QP + state step composed: reproduces the logged CPLEX run to all printed digits."""
    cfg, m = _parity.load_case("maze10")
    sw = _parity.make_oracle(cfg, m, 9)
    wp = sw.pos.copy()
    wp[:5, 0] += 0.5          # PIBT's first waypoints: one lattice step along x (agents 0-4 fly +x, 5-9 fly -x)
    wp[5:, 0] -= 0.5
    sw.waypoint = wp
    st = sw.step()
    assert (st & ~16).max() == 0
    z = np.load(_parity.os.path.join(_parity.ROOT, "tests", "golden", "golden_log.npz"))
    for row, t in ((1, 0.1), (2, 0.2)):
        assert abs(z["t"][row] - t) < 1e-12
        for a in range(10):
            s = oracle.state_at(sw.p, sw.traj[a], t).reshape(9)
            ref = z["state"][row, a]
            for i in range(9):
                assert abs(_sig6(s[i]) - ref[i]) <= 1e-6 * max(1.0, abs(ref[i])), (row, a, i, s[i], ref[i])


def test_edt_cell_index_float_trap(oracle):
    """world_min.y = -0.3f = -0.30000001 -> floor(-3.0000001) = -4."""
    cfg, m = _parity.load_case("maze10")
    sw = _parity.make_oracle(cfg, m, 9)
    assert sw.edt.dims == (81, 48, 26) and sw.edt.min_key == (-20, -4, 0)
