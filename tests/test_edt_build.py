"""Distance-grid construction (dlsc_build_edt, replacing MapManager::updateOctreeFromCSV +
DynamicEDTOctomap, reference src/map_manager.cpp:61-82, 264-316) against the oracle's orc_edt_build: bit-exact
distances and nearest-obstacle indices (ties -> lowest linear cell index).  The CPU tier runs the kernel cores
through the host simulator, the GPU tier the CUDA kernels through the C ABI."""
import numpy as np
import pytest

import _parity
from dlsc_gc_planner_b200 import capi, edt as edtmod, missions
from oracle import oracle_py as O
from test_hostsim_parity import check_worst


def _check_grid(lib, cfg, m, boxes=None):
    boxes = m.boxes if boxes is None else boxes
    p = _parity.oracle_params(cfg, m)
    ref = O.edt_build(p, boxes)
    pl = capi.SwarmPlanner(cfg, m, max_nbr=4, lib=lib)
    pl.build_edt(boxes)
    dist, obst, dims, mk = pl.get_edt()
    assert tuple(dims) == ref.dims and tuple(mk) == ref.min_key
    assert np.array_equal(dist.view(np.uint32), ref.dist.view(np.uint32))
    assert np.array_equal(obst, ref.obst)
    return pl, ref


@pytest.mark.parametrize("name", ["forest10", "maze10"])
def test_hostsim_grid_matches_oracle(hostsim, name):
    cfg, m = _parity.load_case(name)
    pl, ref = _check_grid(hostsim, cfg, m)
    assert (ref.obst[:, 0] >= 0).any()
    pl.close()


def test_hostsim_edge_cases(hostsim):
    cfg, m = _parity.load_case("forest10")
    # no obstacle at all: every cell at the cap, no nearest obstacle
    pl, ref = _check_grid(hostsim, cfg, m, np.zeros((0, 6), np.float32))
    assert (ref.obst == -1).all() and np.all(ref.dist == ref.dist[0])
    pl.close()
    # boxes sticking out of / entirely outside the world, a degenerate (zero-size) box, a one-voxel box
    wmin, wmax = np.asarray(m.world_min, np.float32), np.asarray(m.world_max, np.float32)
    boxes = np.array([[wmin[0], wmin[1], 0.3, 1.0, 1.0, 1.0],
                      [wmax[0] + 3.0, 0.0, 1.0, 0.5, 0.5, 0.5],
                      [0.0, 0.0, 1.0, 0.0, 0.5, 0.5],
                      [0.25, 0.35, 1.05, 0.1, 0.1, 0.1],
                      [wmax[0], wmax[1], wmax[2], 0.4, 0.4, 0.4]], np.float32)
    pl, ref = _check_grid(hostsim, cfg, m, boxes)
    pl.close()


def test_hostsim_occupancy_entry_point(hostsim):
    cfg, m = _parity.load_case("maze10")
    pl, ref = _check_grid(hostsim, cfg, m)
    occ, dims, mk = edtmod.occupancy(m.world_min, m.world_max, cfg.world_res, m.boxes)
    assert tuple(dims) == ref.dims
    pl.build_edt_occupancy(occ)
    dist, obst, _, _ = pl.get_edt()
    assert np.array_equal(dist, ref.dist) and np.array_equal(obst, ref.obst)
    pl.close()


def test_window_limit_is_an_error(hostsim):
    cfg, m = _parity.load_case("forest10")
    pl = capi.SwarmPlanner(cfg, m, max_nbr=4, lib=hostsim)
    with pytest.raises(capi.DlscError):
        pl.build_edt(m.boxes, maxdist=5.0)          # 51 cells > 16
    pl.close()


# ------------------------------------------------------------------------------------------------ GPU tier
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["forest10", "maze10"])
def test_gpu_grid_matches_oracle(cuda_lib, name):
    cfg, m = _parity.load_case(name)
    pl, _ = _check_grid(cuda_lib, cfg, m)
    assert pl.edt_build_ms() > 0.0
    pl.close()


@pytest.mark.gpu
def test_gpu_edge_cases_and_occupancy(cuda_lib):
    cfg, m = _parity.load_case("forest10")
    pl, _ = _check_grid(cuda_lib, cfg, m, np.zeros((0, 6), np.float32))
    pl.close()
    wmin, wmax = np.asarray(m.world_min, np.float32), np.asarray(m.world_max, np.float32)
    boxes = np.array([[wmin[0], wmin[1], 0.3, 1.0, 1.0, 1.0], [wmax[0] + 3.0, 0.0, 1.0, 0.5, 0.5, 0.5],
                      [0.0, 0.0, 1.0, 0.0, 0.5, 0.5], [0.25, 0.35, 1.05, 0.1, 0.1, 0.1],
                      [wmax[0], wmax[1], wmax[2], 0.4, 0.4, 0.4]], np.float32)
    pl, ref = _check_grid(cuda_lib, cfg, m, boxes)
    occ, _, _ = edtmod.occupancy(m.world_min, m.world_max, cfg.world_res, boxes)
    pl.build_edt_occupancy(occ)
    dist, obst, _, _ = pl.get_edt()
    assert np.array_equal(dist, ref.dist) and np.array_equal(obst, ref.obst)
    with pytest.raises(capi.DlscError):
        pl.build_edt(boxes, maxdist=5.0)
    pl.close()


@pytest.mark.gpu
def test_gpu_full_size_forest_grid(cuda_lib):
    """The 641 x 641 x 26 grid of the 4096-agent synthetic forest (BASELINE configs[3]) against the oracle."""
    cfg = missions.PlannerConfig.forest3d()
    m = missions.synthetic_forest(n_agents=4096, half_extent=32.0, seed=4096)
    pl, ref = _check_grid(cuda_lib, cfg, m)
    assert ref.dims == (641, 641, 26)
    pl.close()


@pytest.mark.gpu
def test_gpu_lockstep_on_device_built_grid(cuda_lib):
    """forest10 replans with the SFC stage reading the device-built grid: same boxes / trajectories as the oracle."""
    cfg, m = _parity.load_case("forest10")
    sw = _parity.make_oracle(cfg, m, 9, n_threads=8)
    pl = capi.SwarmPlanner(cfg, m, max_nbr=9, lib=cuda_lib)
    pl.build_edt(m.boxes)
    w = _parity.run_lockstep(pl, sw, m, 20, _parity.default_waypoints(cfg, m))
    check_worst(w)
    pl.close()
