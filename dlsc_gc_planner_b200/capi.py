"""ctypes binding of libdlsc_b200.so (include/dlsc_b200.h) and the `SwarmPlanner` host object.

`SwarmPlanner` mirrors, for a whole agent block at once, what the reference keeps per agent in
AgentManager + TrajPlanner (reference src/agent_manager.cpp:4-108, src/traj_planner.cpp:35-63):
`plan()` = TrajPlanner::plan for every agent, `advance()` = AgentManager::doStep for every agent.

There is no CPU path here: the loader opens the in-tree CUDA library and raises when it is missing
or when no CUDA device is usable.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DLSC_B200_LIB: another build of the same CUDA library (A/B variants of a kernel built with `make OUT=... OBJ=...`)
LIB_PATH = os.environ.get("DLSC_B200_LIB") or os.path.join(_HERE, "libdlsc_b200.so")

OK, QP_MAXITER, QP_NUMERIC, SFC_INIT_FAILED, GOAL_INFEASIBLE, SFC_REUSED, NBR_OVERFLOW, QP_IPM_USED = 0, 1, 2, 4, 8, 16, 32, 64
STAGE_PREDICT, STAGE_NBR, STAGE_LSC, STAGE_SFC, STAGE_GOAL, STAGE_QP, STAGE_ALL = 1, 2, 4, 8, 16, 32, 63
STAGE_NAMES = ("predict", "nbr", "lsc", "sfc", "goal", "qp")
# a replan with one of these bits must not be flown as planned: QP failures fall back to the initial trajectory inside the
# library, the others are errors of the step; NBR_OVERFLOW means collision constraints against the neighbours beyond
# max_nbr are missing (the reference has no cap, multi_sync_simulator.cpp:481-503)
FAIL_MASK = QP_MAXITER | QP_NUMERIC | SFC_INIT_FAILED | GOAL_INFEASIBLE | NBR_OVERFLOW


class DlscParams(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("n", C.c_int32), ("phi", C.c_int32), ("dim", C.c_int32),
        ("use_sfc", C.c_int32), ("max_nbr", C.c_int32),
        ("dt", C.c_double),
        ("world_min", C.c_double * 3), ("world_max", C.c_double * 3),
        ("world_res", C.c_double), ("grid_res", C.c_double), ("z_2d", C.c_double),
        ("comm_range", C.c_double), ("w_control", C.c_double), ("w_terminal", C.c_double),
        ("reset_threshold", C.c_double),
        ("qp_max_iter", C.c_int32), ("qp_solver", C.c_int32), ("qp_screen_slack", C.c_double),
        ("qp_active_max", C.c_int32), ("reserved0", C.c_int32),
    ]


class DlscAgents(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("vel", C.c_void_p), ("acc", C.c_void_p), ("waypoint", C.c_void_p),
                ("disturbed", C.c_void_p)]


class DlscAgentProps(C.Structure):
    _fields_ = [("radius", C.c_void_p), ("downwash", C.c_void_p), ("max_vel", C.c_void_p),
                ("max_acc", C.c_void_p), ("nominal_vel", C.c_void_p)]


class DlscObstacles(C.Structure):
    _fields_ = [("n", C.c_int32), ("pos", C.c_void_p), ("vel", C.c_void_p), ("radius", C.c_void_p),
                ("downwash", C.c_void_p), ("max_acc", C.c_void_p)]


class DlscObstacleParams(C.Structure):
    _fields_ = [("slack_collision_weight", C.c_double), ("uncertainty_horizon", C.c_double),
                ("size_prediction", C.c_int32), ("reserved0", C.c_int32)]


EXPORTS = (
    "dlsc_last_error dlsc_abi_version dlsc_device_count dlsc_create dlsc_destroy dlsc_set_stream dlsc_get_stream "
    "dlsc_set_edt dlsc_set_agent_props dlsc_reset dlsc_set_agents dlsc_records_device dlsc_record_floats "
    "dlsc_bind_records dlsc_set_records dlsc_get_records dlsc_step dlsc_run_stages dlsc_advance "
    "dlsc_publish_records dlsc_sync dlsc_get_seq dlsc_set_seq dlsc_get_traj dlsc_get_qp_x dlsc_get_cost "
    "dlsc_get_violation dlsc_get_qp_iters dlsc_get_status dlsc_get_goal dlsc_get_state dlsc_get_init_traj "
    "dlsc_get_pred_traj dlsc_get_neighbours dlsc_get_lsc dlsc_get_sfc dlsc_set_sfc dlsc_enable_timing "
    "dlsc_get_timings dlsc_launch_count dlsc_get_counters dlsc_waypoint_device dlsc_traj_device "
    "dlsc_set_waypoints_device dlsc_measure_fp64_peak dlsc_run_stages_subset dlsc_set_init_traj "
    "dlsc_set_pred_traj dlsc_set_neighbours dlsc_set_lsc dlsc_set_groups dlsc_edt_dims dlsc_build_edt "
    "dlsc_build_edt_occupancy dlsc_get_edt dlsc_edt_build_ms dlsc_p2p_export dlsc_p2p_connect dlsc_exchange_records "
    "dlsc_p2p_status dlsc_p2p_disconnect dlsc_gjk_batch dlsc_cuda_build dlsc_wp_last_error dlsc_wp_create dlsc_wp_destroy "
    "dlsc_wp_dims dlsc_wp_set_grid dlsc_wp_set_nodes dlsc_wp_get_nodes dlsc_wp_pibt dlsc_wp_step dlsc_wp_pibt_timesteps "
    "dlsc_set_obstacles dlsc_get_slack dlsc_get_trap dlsc_get_obstacle_pred dlsc_bind_traj_host dlsc_wp_set_warning dlsc_wp_pibt_obs dlsc_wp_set_obstacles dlsc_wp_set_alerts dlsc_wp_get_warning").split()


def build_library(force=False):
    """Compile libdlsc_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", src, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", src, "all"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def _declare(lib):
    lib.dlsc_last_error.restype = C.c_char_p
    if hasattr(lib, "dlsc_wp_last_error"):
        lib.dlsc_wp_last_error.restype = C.c_char_p
        lib.dlsc_wp_pibt_timesteps.restype = C.c_int64
        lib.dlsc_wp_pibt_timesteps.argtypes = [C.c_void_p]
        lib.dlsc_wp_destroy.argtypes = [C.c_void_p]
    for name in ("dlsc_records_device", "dlsc_get_stream", "dlsc_waypoint_device", "dlsc_traj_device"):
        if hasattr(lib, name):
            getattr(lib, name).restype = C.c_void_p
            getattr(lib, name).argtypes = [C.c_void_p]
    if hasattr(lib, "dlsc_edt_build_ms"):
        lib.dlsc_edt_build_ms.restype = C.c_double
        lib.dlsc_edt_build_ms.argtypes = [C.c_void_p]
    if hasattr(lib, "dlsc_launch_count"):
        lib.dlsc_launch_count.restype = C.c_int64
        lib.dlsc_launch_count.argtypes = [C.c_void_p]
    return lib


_LIB = None


def load_library(path=None):
    """Open the CUDA library.  Raises (no fallback) when it has not been built."""
    global _LIB
    if path is None:
        if _LIB is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError("libdlsc_b200.so is not built: run `python -c 'import __graft_entry__ as g; "
                                   "g.build()'` or `make -C dlsc_gc_planner_b200/csrc` (no CPU fallback exists)")
            lib = _declare(C.CDLL(LIB_PATH))
            if not hasattr(lib, "dlsc_cuda_build"):
                # DLSC_B200_LIB may only name another build of the CUDA library (A/B variants), never a CPU stand-in
                raise RuntimeError("%s is not a CUDA build of libdlsc_b200 (no dlsc_cuda_build symbol): the product has "
                                   "no CPU path" % LIB_PATH)
            _LIB = lib
        return _LIB
    return _declare(C.CDLL(path))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_params(cfg, world_min, world_max, max_nbr, qp_max_iter=0, qp_screen_slack=0.0, qp_solver=0, qp_active_max=0):
    """cfg: missions.PlannerConfig."""
    p = DlscParams()
    p.M, p.n, p.phi, p.dim, p.use_sfc, p.max_nbr = cfg.M, cfg.n, cfg.phi, cfg.dim, int(cfg.use_sfc), int(max_nbr)
    p.dt = cfg.dt
    for k in range(3):
        p.world_min[k] = float(np.float32(world_min[k]))
        p.world_max[k] = float(np.float32(world_max[k]))
    p.world_res, p.grid_res, p.z_2d = cfg.world_res, cfg.grid_res, cfg.z_2d
    p.comm_range, p.w_control, p.w_terminal = cfg.comm_range, cfg.w_control, cfg.w_terminal
    p.reset_threshold = cfg.reset_threshold
    p.qp_max_iter = qp_max_iter
    p.qp_screen_slack = qp_screen_slack
    p.qp_solver = qp_solver
    p.qp_active_max = qp_active_max
    return p


class DlscError(RuntimeError):
    pass


class SwarmPlanner:
    """One context = the agent block [begin, begin+n_local) of a swarm of n_agents on one GPU."""

    def __init__(self, cfg, mission, max_nbr=None, begin=0, n_local=None, device=0, lib=None, qp_max_iter=0,
                 qp_screen_slack=0.0, qp_solver=0, qp_active_max=0):
        self.lib = lib if lib is not None else load_library()
        self.cfg = cfg
        self.N = int(mission.n_agents)
        self.begin = int(begin)
        self.NL = int(n_local if n_local is not None else self.N - begin)
        self.M, self.P, self.D = cfg.M, cfg.n + 1, cfg.dim
        self.K = int(max_nbr if max_nbr is not None else max(self.N - 1, 1))
        self.params = make_params(cfg, mission.world_min, mission.world_max, self.K, qp_max_iter, qp_screen_slack, qp_solver, qp_active_max)
        self.ctx = C.c_void_p()
        self._ck(self.lib.dlsc_create(C.byref(self.params), self.N, self.begin, self.NL, int(device), C.byref(self.ctx)))
        sl = slice(self.begin, self.begin + self.NL)
        f64 = lambda x: np.ascontiguousarray(x[sl], np.float64)
        self._props = [f64(mission.radius), f64(mission.downwash), f64(mission.max_vel), f64(mission.max_acc),
                       f64(mission.nominal_vel)]
        pr = DlscAgentProps(*[a.ctypes.data for a in self._props])
        self._ck(self.lib.dlsc_set_agent_props(self.ctx, C.byref(pr)))
        start = np.ascontiguousarray(mission.start[sl], np.float32).copy()
        if cfg.dim == 2:
            start[:, 2] = np.float32(cfg.z_2d)
        self.start = start
        self._ck(self.lib.dlsc_reset(self.ctx, _p(start)))
        self.rec_floats = int(self.lib.dlsc_record_floats(self.ctx))

    # -- plumbing -------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise DlscError((self.lib.dlsc_last_error() or b"?").decode())

    def close(self):
        if self.ctx:
            self.lib.dlsc_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs ---------------------------------------------------------------------------------
    def set_edt(self, dist, obst, dims, min_key, res):
        dist = np.ascontiguousarray(dist, np.float32)
        obst = np.ascontiguousarray(obst, np.int32)
        d3 = (C.c_int32 * 3)(*[int(x) for x in dims])
        k3 = (C.c_int32 * 3)(*[int(x) for x in min_key])
        self._ck(self.lib.dlsc_set_edt(self.ctx, _p(dist), _p(obst), d3, k3, C.c_double(res)))

    def edt_dims(self):
        d3, k3 = (C.c_int32 * 3)(), (C.c_int32 * 3)()
        self._ck(self.lib.dlsc_edt_dims(self.ctx, d3, k3))
        return tuple(d3), tuple(k3)

    def build_edt(self, boxes, maxdist=1.0):
        """Distance grid from the mission's CSV boxes [nb][6] (cx, cy, cz, sx, sy, sz), built on the device
        (MapManager::updateOctreeFromCSV + setGlobalMap, src/map_manager.cpp:61-82, 264-316)."""
        b = np.ascontiguousarray(np.asarray(boxes, np.float32).reshape(-1, 6))
        self._ck(self.lib.dlsc_build_edt(self.ctx, _p(b) if len(b) else None, C.c_int(len(b)), C.c_double(maxdist)))

    def build_edt_occupancy(self, occ, maxdist=1.0):
        dims, _ = self.edt_dims()
        o = np.ascontiguousarray(occ, np.uint8)
        assert o.size == dims[0] * dims[1] * dims[2]
        self._ck(self.lib.dlsc_build_edt_occupancy(self.ctx, _p(o), C.c_double(maxdist)))

    def get_edt(self):
        """-> dist [ncell] f32, obst [ncell][3] i32, dims, min_key (dlsc_set_edt layout)"""
        dims, mk = self.edt_dims()
        nc = dims[0] * dims[1] * dims[2]
        dist = np.empty(nc, np.float32)
        obst = np.empty((nc, 3), np.int32)
        self._ck(self.lib.dlsc_get_edt(self.ctx, _p(dist), _p(obst)))
        return dist, obst, dims, mk

    def edt_build_ms(self):
        return float(self.lib.dlsc_edt_build_ms(self.ctx))

    # -- record exchange over peer memory (one process per GPU) -----------------------------------
    def p2p_export(self):
        h = (C.c_ubyte * 64)()
        self._ck(self.lib.dlsc_p2p_export(self.ctx, h))
        return bytes(h)

    def p2p_connect(self, world, rank, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._ck(self.lib.dlsc_p2p_connect(self.ctx, C.c_int(world), C.c_int(rank), buf))

    def exchange_records(self):
        self._ck(self.lib.dlsc_exchange_records(self.ctx))

    def p2p_status(self):
        self._ck(self.lib.dlsc_p2p_status(self.ctx))

    def p2p_disconnect(self):
        self._ck(self.lib.dlsc_p2p_disconnect(self.ctx))

    def set_groups(self, group):
        """Mission index per local agent (Monte-Carlo batches); call after construction / reset."""
        g = np.ascontiguousarray(group, np.int32)
        assert g.shape == (self.NL,)
        self._ck(self.lib.dlsc_set_groups(self.ctx, _p(g)))

    def set_obstacles(self, pos, vel=None, radius=0.15, downwash=1.0, max_acc=0.0, slack_weight=1.0,
                      size_prediction=True, uncertainty_horizon=1.0):
        """Dynamic (non-agent) obstacles of the next replans (dlsc_set_obstacles); pos=None removes them.  Scalars are
        broadcast.  Mirrors the Obstacle list TrajPlanner::setObstacles receives for non-agent entries plus the three
        Param fields the dynamic-obstacle path reads."""
        if pos is None:
            self._ck(self.lib.dlsc_set_obstacles(self.ctx, None, None))
            self.n_dyn = 0
            return
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        n = pos.shape[0]
        vel = np.ascontiguousarray(np.zeros((n, 3)) if vel is None else vel, np.float32).reshape(n, 3)
        full = lambda v: np.full(n, v, np.float64) if np.isscalar(v) else np.ascontiguousarray(v, np.float64)
        radius, downwash, max_acc = full(radius), full(downwash), full(max_acc)
        o = DlscObstacles(n, _p(pos), _p(vel), _p(radius), _p(downwash), _p(max_acc))
        op = DlscObstacleParams(float(slack_weight), float(uncertainty_horizon), int(bool(size_prediction)), 0)
        self._ck(self.lib.dlsc_set_obstacles(self.ctx, C.byref(o), C.byref(op)))
        self.n_dyn = n

    def slack(self):
        """[n_local][n_obstacles][M] slack variables of the last QP (<= 0)."""
        nd = getattr(self, "n_dyn", 0)
        out = np.zeros((self.NL, nd, self.M), np.float64)
        if nd:
            self._ck(self.lib.dlsc_get_slack(self.ctx, _p(out)))
        return out

    def trap(self):
        out = np.zeros(self.NL, np.uint8)
        self._ck(self.lib.dlsc_get_trap(self.ctx, _p(out)))
        return out

    def obstacle_pred(self):
        nd = getattr(self, "n_dyn", 0)
        out = np.zeros((nd, self.M, 6, 3), np.float32)
        if nd:
            self._ck(self.lib.dlsc_get_obstacle_pred(self.ctx, _p(out)))
        return out

    def set_agents(self, pos=None, vel=None, acc=None, waypoint=None, disturbed=None):
        f = lambda x: None if x is None else np.ascontiguousarray(x, np.float32)
        keep = [f(pos), f(vel), f(acc), f(waypoint),
                None if disturbed is None else np.ascontiguousarray(disturbed, np.uint8)]
        a = DlscAgents(*[None if x is None else x.ctypes.data for x in keep])
        self._ck(self.lib.dlsc_set_agents(self.ctx, C.byref(a)))
        self._ck(self.lib.dlsc_sync(self.ctx))     # host arrays may go away after return

    def set_agents_async(self, agents_struct):
        """Caller keeps the (pinned) arrays alive; no synchronisation."""
        self._ck(self.lib.dlsc_set_agents(self.ctx, C.byref(agents_struct)))

    def set_records(self, first, rec):
        rec = np.ascontiguousarray(rec, np.float32)
        self._ck(self.lib.dlsc_set_records(self.ctx, int(first), int(rec.shape[0]), _p(rec)))

    def get_records(self, first=0, count=None):
        count = self.N - first if count is None else count
        out = np.zeros((count, self.rec_floats), np.float32)
        self._ck(self.lib.dlsc_get_records(self.ctx, int(first), int(count), _p(out)))
        return out

    def set_sfc(self, sfc=None, init_flag=None):
        sfc = None if sfc is None else np.ascontiguousarray(sfc, np.float32)
        fl = None if init_flag is None else np.ascontiguousarray(init_flag, np.uint8)
        self._ck(self.lib.dlsc_set_sfc(self.ctx, _p(sfc), _p(fl)))

    # -- the path -------------------------------------------------------------------------------
    def plan(self):
        """One replan of every agent of the block (stream ordered, asynchronous)."""
        self._ck(self.lib.dlsc_step(self.ctx))

    def run_stages(self, mask):
        self._ck(self.lib.dlsc_run_stages(self.ctx, int(mask)))

    def advance(self):
        self._ck(self.lib.dlsc_advance(self.ctx))

    def bind_traj_host(self, array):
        """Every replan delivers the trajectories into `array` ([n_local][M][6][3] float32, ideally pinned): complete after
        sync().  None unbinds.  The caller keeps the array alive."""
        if array is None:
            self._ck(self.lib.dlsc_bind_traj_host(self.ctx, None))
            self._traj_host = None
            return
        assert array.dtype == np.float32 and array.flags["C_CONTIGUOUS"] and array.size == self.NL * self.M * self.P * 3
        self._ck(self.lib.dlsc_bind_traj_host(self.ctx, C.c_void_p(array.ctypes.data)))
        self._traj_host = array

    def publish_records(self):
        self._ck(self.lib.dlsc_publish_records(self.ctx))

    def sync(self):
        self._ck(self.lib.dlsc_sync(self.ctx))

    @property
    def seq(self):
        return int(self.lib.dlsc_get_seq(self.ctx))

    @seq.setter
    def seq(self, v):
        self._ck(self.lib.dlsc_set_seq(self.ctx, int(v)))

    # -- outputs --------------------------------------------------------------------------------
    def _get(self, fn, shape, dtype):
        out = np.zeros(shape, dtype)
        self._ck(fn(self.ctx, _p(out)))
        return out

    def traj(self):
        return self._get(self.lib.dlsc_get_traj, (self.NL, self.M, self.P, 3), np.float32)

    def qp_x(self):
        return self._get(self.lib.dlsc_get_qp_x, (self.NL, self.D, self.M, self.P), np.float64)

    def cost(self):
        return self._get(self.lib.dlsc_get_cost, (self.NL,), np.float64)

    def violation(self):
        return self._get(self.lib.dlsc_get_violation, (self.NL,), np.float64)

    def qp_iters(self):
        return self._get(self.lib.dlsc_get_qp_iters, (self.NL,), np.int32)

    def status(self):
        return self._get(self.lib.dlsc_get_status, (self.NL,), np.int32)

    def goal(self):
        return self._get(self.lib.dlsc_get_goal, (self.NL, 3), np.float32)

    def init_traj(self):
        return self._get(self.lib.dlsc_get_init_traj, (self.NL, self.M, self.P, 3), np.float32)

    def pred_traj(self):
        return self._get(self.lib.dlsc_get_pred_traj, (self.N, self.M, self.P, 3), np.float32)

    def sfc(self):
        return self._get(self.lib.dlsc_get_sfc, (self.NL, self.M, 6), np.float32)

    def state(self):
        pos = np.zeros((self.NL, 3), np.float32)
        vel = np.zeros((self.NL, 3), np.float32)
        acc = np.zeros((self.NL, 3), np.float32)
        self._ck(self.lib.dlsc_get_state(self.ctx, _p(pos), _p(vel), _p(acc)))
        return pos, vel, acc

    def neighbours(self):
        idx = np.zeros((self.NL, self.K), np.int32)
        cnt = np.zeros(self.NL, np.int32)
        self._ck(self.lib.dlsc_get_neighbours(self.ctx, _p(idx), _p(cnt)))
        return idx, cnt

    def lsc(self, with_anchor=True):
        normal = np.zeros((self.NL, self.K, self.M, 3), np.float32)
        d = np.zeros((self.NL, self.K, self.M, self.P), np.float64)
        anchor = np.zeros((self.NL, self.K, self.M, self.P, 3), np.float32) if with_anchor else None
        self._ck(self.lib.dlsc_get_lsc(self.ctx, _p(normal), _p(anchor), _p(d)))
        return normal, anchor, d

    def set_init_traj(self, traj):
        t = np.ascontiguousarray(traj, np.float32)
        assert t.shape == (self.NL, self.M, self.P, 3)
        self._ck(self.lib.dlsc_set_init_traj(self.ctx, _p(t)))

    def set_pred_traj(self, traj):
        t = np.ascontiguousarray(traj, np.float32)
        assert t.shape == (self.N, self.M, self.P, 3)
        self._ck(self.lib.dlsc_set_pred_traj(self.ctx, _p(t)))

    def set_neighbours(self, idx, cnt):
        i = np.ascontiguousarray(idx, np.int32)
        n = np.ascontiguousarray(cnt, np.int32)
        assert i.shape == (self.NL, self.K) and n.shape == (self.NL,)
        self._ck(self.lib.dlsc_set_neighbours(self.ctx, _p(i), _p(n)))

    def gjk_batch(self, pts):
        """Per-kernel parity entry: hulls [n][6][3] f64 -> witness v [n][3], iterations, simplex size, leaf bit set."""
        p = np.ascontiguousarray(pts, np.float64)
        n = p.shape[0]
        assert p.shape == (n, 6, 3)
        v = np.zeros((n, 3), np.float64)
        it = np.zeros(n, np.int32)
        sn = np.zeros(n, np.int32)
        lv = np.zeros(n, np.uint64)
        self._ck(self.lib.dlsc_gjk_batch(self.ctx, _p(p), C.c_int(n), _p(v), _p(it), _p(sn), _p(lv)))
        return v, it, sn, lv

    def counters(self):
        out = np.zeros(16, np.int64)
        self._ck(self.lib.dlsc_get_counters(self.ctx, _p(out)))
        return dict(pairs=int(out[0]), gjk_iters=int(out[1]), edt_lookups=int(out[2]), qp_iters=int(out[3]),
                    qp_rows=int(out[4]), sfc_tests_mask=int(out[5]), sfc_tests_records=int(out[6]), sfc_tests_sat=int(out[7]), sfc_vertices_alg=int(out[8]))

    def enable_timing(self, on=True):
        self._ck(self.lib.dlsc_enable_timing(self.ctx, int(on)))

    def timings(self):
        ms = (C.c_double * 6)()
        n = C.c_int(0)
        self._ck(self.lib.dlsc_get_timings(self.ctx, ms, C.byref(n)))
        return dict(zip(STAGE_NAMES, [float(x) for x in ms])), n.value

    def launch_count(self):
        return int(self.lib.dlsc_launch_count(self.ctx))

    def set_waypoints_device(self, device_ptr):
        self._ck(self.lib.dlsc_set_waypoints_device(self.ctx, C.c_void_p(int(device_ptr))))

    def set_stream(self, cuda_stream):
        self._ck(self.lib.dlsc_set_stream(self.ctx, C.c_void_p(int(cuda_stream))))

    def measure_fp64_peak(self):
        v = C.c_double(0)
        self._ck(self.lib.dlsc_measure_fp64_peak(self.ctx, C.byref(v)))
        return v.value

    def records_device_ptr(self):
        return int(self.lib.dlsc_records_device(self.ctx))

    def bind_records(self, device_ptr):
        self._ck(self.lib.dlsc_bind_records(self.ctx, C.c_void_p(int(device_ptr))))


class WaypointProvider:
    """Host object over dlsc_wp_*: the reference's waypoint layer (comm-range groups + PIBT + update rules,
    MultiSyncSimulator::decentralizedMAPP, reference src/multi_sync_simulator.cpp:308-466) for a whole swarm."""

    def __init__(self, cfg, mission, lib=None, edt=None):
        self.lib = lib if lib is not None else load_library()
        self.cfg = cfg
        self.N = int(mission.n_agents)
        self.M, self.P = cfg.M, cfg.n + 1
        self.params = make_params(cfg, mission.world_min, mission.world_max, 1)
        start = np.ascontiguousarray(mission.start, np.float32).copy()
        goal = np.ascontiguousarray(mission.goal, np.float32).copy()
        if cfg.dim == 2:
            start[:, 2] = np.float32(cfg.z_2d); goal[:, 2] = np.float32(cfg.z_2d)
        self.ctx = C.c_void_p()
        self._ck(self.lib.dlsc_wp_create(C.byref(self.params), self.N, _p(start), _p(goal), C.c_double(float(mission.radius[0])),
                                         C.c_double(float(mission.downwash[0])), C.byref(self.ctx)))
        if edt is not None:
            self.set_grid(*edt)

    def _ck(self, rc):
        if rc < 0:
            raise DlscError((self.lib.dlsc_wp_last_error() or b"?").decode())
        return rc

    def close(self):
        if self.ctx:
            self.lib.dlsc_wp_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dims(self):
        d3 = (C.c_int32 * 3)()
        self._ck(self.lib.dlsc_wp_dims(self.ctx, d3))
        return tuple(d3)

    def set_grid(self, dist, obst, dims, min_key, res):
        dist = np.ascontiguousarray(dist, np.float32); obst = np.ascontiguousarray(obst, np.int32)
        d3 = (C.c_int32 * 3)(*[int(x) for x in dims]); k3 = (C.c_int32 * 3)(*[int(x) for x in min_key])
        self._ck(self.lib.dlsc_wp_set_grid(self.ctx, _p(dist), _p(obst), d3, k3, C.c_double(res)))

    def set_nodes(self, dims, exists):
        d3 = (C.c_int32 * 3)(*[int(x) for x in dims])
        e = np.ascontiguousarray(exists, np.uint8)
        assert e.size == int(dims[0]) * int(dims[1]) * int(dims[2])
        self._ck(self.lib.dlsc_wp_set_nodes(self.ctx, d3, _p(e)))

    def nodes(self):
        w, d, h = self.dims()
        out = np.zeros(w * d * h, np.uint8)
        self._ck(self.lib.dlsc_wp_get_nodes(self.ctx, _p(out)))
        return out

    def pibt(self, start, current, goal, max_t=6000):
        s, c, g = (np.ascontiguousarray(x, np.int32) for x in (start, current, goal))
        plan = np.zeros((max_t, len(s)), np.int32)
        T = self._ck(self.lib.dlsc_wp_pibt(self.ctx, C.c_int(len(s)), _p(s), _p(c), _p(g), C.c_int(max_t), _p(plan)))
        return plan[:T].copy()

    def set_warning(self, warning):
        """Warning flags of the lattice nodes ([w*d*h] uint8; None clears them)."""
        if warning is None:
            self._ck(self.lib.dlsc_wp_set_warning(self.ctx, None))
        else:
            wr = np.ascontiguousarray(warning, np.uint8)
            self._ck(self.lib.dlsc_wp_set_warning(self.ctx, _p(wr)))

    def set_obstacles(self, pos, vel=None, radius=0.15, max_acc=0.0, uncertainty_horizon=1.0):
        """Dynamic obstacles of the next step() calls (None removes them)."""
        if pos is None:
            self._ck(self.lib.dlsc_wp_set_obstacles(self.ctx, 0, None, None, None, None, C.c_double(uncertainty_horizon)))
            return
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        n = pos.shape[0]
        vel = np.ascontiguousarray(np.zeros((n, 3)) if vel is None else vel, np.float32).reshape(n, 3)
        full = lambda v: np.full(n, v, np.float64) if np.isscalar(v) else np.ascontiguousarray(v, np.float64)
        r, ma = full(radius), full(max_acc)
        self._ck(self.lib.dlsc_wp_set_obstacles(self.ctx, n, _p(pos), _p(vel), _p(r), _p(ma), C.c_double(uncertainty_horizon)))

    def warning(self):
        w, d, h = self.dims()
        out = np.zeros(w * d * h, np.uint8)
        self._ck(self.lib.dlsc_wp_get_warning(self.ctx, _p(out)))
        return out

    def set_alerts(self, alerts):
        """alerts: per agent a list of obstacle ids (TrajOptResult::collision_alert), or None."""
        if alerts is None:
            self._ck(self.lib.dlsc_wp_set_alerts(self.ctx, None, None, 0))
            return
        stride = max(1, max(len(a) for a in alerts))
        cnt = np.array([len(a) for a in alerts], np.int32)
        ids = np.full((len(alerts), stride), -1, np.int32)
        for i, a in enumerate(alerts):
            ids[i, :len(a)] = a
        self._ck(self.lib.dlsc_wp_set_alerts(self.ctx, _p(cnt), _p(ids), stride))

    def pibt_obs(self, start, current, goal, obs_node, obs_dist, max_t=6000):
        s, c, g, o = (np.ascontiguousarray(x, np.int32) for x in (start, current, goal, obs_node))
        od = np.ascontiguousarray(obs_dist, np.float32)
        plan = np.zeros((max_t, len(s)), np.int32)
        T = self._ck(self.lib.dlsc_wp_pibt_obs(self.ctx, C.c_int(len(s)), _p(s), _p(c), _p(g), _p(o), _p(od), C.c_int(max_t), _p(plan)))
        return plan[:T].copy()

    def step(self, pos, goal_cur, traj, waypoint):
        """-> updated waypoints [N][3] (traj None before the first replan)"""
        pos = np.ascontiguousarray(pos, np.float32); gc = np.ascontiguousarray(goal_cur, np.float32)
        wp = np.ascontiguousarray(waypoint, np.float32).copy()
        tr = None if traj is None else np.ascontiguousarray(traj, np.float32)
        assert pos.shape == (self.N, 3) and gc.shape == (self.N, 3) and wp.shape == (self.N, 3)
        assert tr is None or tr.shape == (self.N, self.M, self.P, 3)
        self._ck(self.lib.dlsc_wp_step(self.ctx, _p(pos), _p(gc), _p(tr), _p(wp)))
        return wp

    def pibt_timesteps(self):
        return int(self.lib.dlsc_wp_pibt_timesteps(self.ctx))
