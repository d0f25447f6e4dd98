"""Octomap .bt worlds (reference src/map_manager.cpp:66-73: OcTree::readBinary + expand; the maze_tro2022 worlds ship in
this format): host-side loader missions.load_world_bt -> occupancy grid -> dlsc_build_edt_occupancy.  octomap is absent
here, so the reader follows the published OcTree binary format and is checked (i) by writing trees in that format and
reading them back, (ii) by the node count every .bt header states (a whole-file check of the traversal), on all the
reference's files when the tree is present, (iii) by planning the reference's maze10_tro2022 mission in that world."""
import glob
import os

import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi, missions

GOLD = os.path.join(_parity.ROOT, "tests", "golden", "world_bt.npz")


def write_bt(path, occupied, free, res=0.1):
    """Minimal writer of the OcTree binary format: `occupied` / `free` are sets of finest-level cells (absolute keys);
    no pruning.  Returns the node count written to the header."""
    def build(x0, y0, z0, edge):
        if edge == 1:
            return "occ" if (x0, y0, z0) in occupied else ("free" if (x0, y0, z0) in free else None)
        h = edge // 2
        kids = [build(x0 + (i & 1) * h, y0 + ((i >> 1) & 1) * h, z0 + ((i >> 2) & 1) * h, h) if touched(x0 + (i & 1) * h, y0 + ((i >> 1) & 1) * h, z0 + ((i >> 2) & 1) * h, h) else None
                for i in range(8)]
        return kids if any(k is not None for k in kids) else None
    cells = occupied | free
    def touched(x0, y0, z0, edge):
        return any(x0 <= x < x0 + edge and y0 <= y < y0 + edge and z0 <= z < z0 + edge for x, y, z in cells)
    root = build(0, 0, 0, 65536)
    out, count = bytearray(), [1]
    def emit(node):
        bits = 0
        for i, k in enumerate(node):
            if k is None:
                continue
            count[0] += 1
            bits |= (3 if isinstance(k, list) else (2 if k == "occ" else 1)) << (2 * i)      # (b0,b1): free (1,0) occ (0,1) inner (1,1)
        out.extend(bytes([bits & 255, bits >> 8]))
        for k in node:
            if isinstance(k, list):
                emit(k)
    emit(root)
    with open(path, "wb") as f:
        f.write(b"# Octomap OcTree binary file\n# test\nid OcTree\nsize %d\nres %g\ndata\n" % (count[0], res))
        f.write(bytes(out))
    return count[0]


def test_bt_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    occ = {tuple(int(v) for v in rng.integers(32768 - 20, 32768 + 20, 3)) for _ in range(60)}
    free = {tuple(int(v) for v in rng.integers(32768 - 20, 32768 + 20, 3)) for _ in range(60)} - occ
    p = str(tmp_path / "t.bt")
    write_bt(p, occ, free, res=0.25)
    res, cubes = missions.load_world_bt(p)
    assert res == 0.25 and np.all(cubes[:, 3] == 1)
    assert {tuple(int(v) + 32768 for v in c[:3]) for c in cubes.tolist()} == occ
    grid = missions.occupancy_from_cubes(cubes, (40, 40, 40), (-20, -20, -20))
    assert grid.sum() == len(occ)
    for x, y, z in occ:
        assert grid[x - 32768 + 20, y - 32768 + 20, z - 32768 + 20] == 1
    # a wrong node count in the header is detected
    raw = open(p, "rb").read().replace(b"size ", b"size 1", 1)
    open(p, "wb").write(raw)
    with pytest.raises(ValueError):
        missions.load_world_bt(p)


@pytest.mark.skipif(not os.path.isdir("/root/reference/world/maze_tro2022"), reason="reference tree not present")
def test_every_reference_bt_file_parses():
    z = np.load(GOLD)
    files = sorted(glob.glob("/root/reference/world/maze_tro2022/*.bt"))
    assert len(files) >= 20
    for f in files:
        res, cubes = missions.load_world_bt(f)            # raises unless the traversal visits exactly `size` nodes
        assert res == 0.1 and len(cubes) > 1000
        lo, hi = cubes[:, :3].min(0), (cubes[:, :3] + cubes[:, 3:4]).max(0)
        assert np.all(lo >= 0) and np.all(hi <= [72, 72, 25])                       # the 7.2 x 7.2 x 2.5 m maze
        if f.endswith("maze9_1.bt"):
            assert np.array_equal(cubes, z["cubes"])


def bt_case():
    z = np.load(GOLD)
    cfg = missions.PlannerConfig.maze2d()
    m = missions.Mission(z["world_min"], z["world_max"], z["start"], z["goal"], z["radius"], z["downwash"], z["max_vel"],
                         z["max_acc"], z["nominal_vel"])
    return cfg, m, z["cubes"]


def run_bt_world(lib, steps):
    from oracle import oracle_py as O
    cfg, m, cubes = bt_case()
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=lib)
    dims, mk = pl.edt_dims()
    occ = missions.occupancy_from_cubes(cubes, dims, mk)
    assert occ.sum() > 50000
    pl.build_edt_occupancy(occ)
    dist, obst, dims2, mk2 = pl.get_edt()
    # every agent starts and ends in free space of this world, walls are where the file puts them
    for q in np.concatenate([m.start, m.goal]):
        c = np.floor(q.astype(np.float64) / cfg.world_res).astype(int) - np.array(mk)
        assert occ[c[0], c[1], c[2]] == 0
    # lock-step against the oracle on the very same grid arrays
    p = _parity.oracle_params(cfg, m)
    sw = O.Swarm(p, m.start, m.goal, m.radius, m.downwash, m.max_vel, m.max_acc, m.nominal_vel,
                 edt=O.Edt(p, dist, obst, dims2, mk2), max_nbr=9, n_threads=os.cpu_count() or 1)
    boxes = np.array([[(c[0] + c[3] / 2) * 0.1, (c[1] + c[3] / 2) * 0.1, (c[2] + c[3] / 2) * 0.1, c[3] * 0.1, c[3] * 0.1, c[3] * 0.1]
                      for c in cubes], np.float32)
    mm = missions.Mission(m.world_min, m.world_max, m.start, m.goal, m.radius, m.downwash, m.max_vel, m.max_acc, m.nominal_vel, boxes)
    w = _parity.run_lockstep(pl, sw, mm, steps, _parity.default_waypoints(cfg, mm))
    pl.close()
    return w


def test_bt_world_mission_hostsim(hostsim):
    from test_hostsim_parity import check_worst
    check_worst(run_bt_world(hostsim, 12))


@pytest.mark.gpu
def test_bt_world_mission_gpu(cuda_lib):
    from test_hostsim_parity import check_worst
    check_worst(run_bt_world(cuda_lib, 40))
