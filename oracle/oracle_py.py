"""ctypes binding of the CPU ORACLE (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (dlsc_gc_planner_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


class OrcParams(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("n", C.c_int32), ("phi", C.c_int32), ("dim", C.c_int32),
        ("use_sfc", C.c_int32), ("reserved0", C.c_int32),
        ("dt", C.c_double),
        ("world_min", C.c_double * 3), ("world_max", C.c_double * 3),
        ("world_res", C.c_double), ("grid_res", C.c_double), ("z_2d", C.c_double),
        ("comm_range", C.c_double), ("w_control", C.c_double), ("w_terminal", C.c_double),
        ("reset_threshold", C.c_double),
    ]


class OrcEdt(C.Structure):
    _fields_ = [
        ("dims", C.c_int32 * 3), ("min_key", C.c_int32 * 3), ("res", C.c_double),
        ("dist", C.c_void_p), ("obst", C.c_void_p),
    ]


class OrcStepIO(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("seq", C.c_int), ("max_nbr", C.c_int), ("n_threads", C.c_int),
        ("pos", C.c_void_p), ("vel", C.c_void_p), ("acc", C.c_void_p), ("waypoint", C.c_void_p),
        ("disturbed", C.c_void_p), ("radius", C.c_void_p), ("downwash", C.c_void_p),
        ("max_vel", C.c_void_p), ("max_acc", C.c_void_p), ("nominal_vel", C.c_void_p),
        ("edt", C.c_void_p),
        ("goal_cur", C.c_void_p), ("prev_traj", C.c_void_p), ("sfc", C.c_void_p),
        ("sfc_init_flag", C.c_void_p),
        ("init_traj", C.c_void_p), ("pred_traj", C.c_void_p), ("nbr_idx", C.c_void_p),
        ("nbr_cnt", C.c_void_p), ("lsc_normal", C.c_void_p), ("lsc_anchor", C.c_void_p),
        ("lsc_d", C.c_void_p), ("qp_x", C.c_void_p), ("cost", C.c_void_p),
        ("max_violation", C.c_void_p), ("qp_iters", C.c_void_p), ("status", C.c_void_p),
        ("stage_seconds", C.c_void_p),
        ("n_dyn", C.c_int), ("dyn_size_prediction", C.c_int), ("dyn_uncertainty_horizon", C.c_double),
        ("slack_collision_weight", C.c_double),
        ("dyn_pos", C.c_void_p), ("dyn_vel", C.c_void_p), ("dyn_radius", C.c_void_p),
        ("dyn_downwash", C.c_void_p), ("dyn_max_acc", C.c_void_p),
        ("comm_box", C.c_void_p), ("qp_slack", C.c_void_p), ("trap", C.c_void_p),
    ]


def build(force=False):
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "dlsc_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/src/openGJK/openGJK.cpp"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_gjk_hull_origin.restype = C.c_double
        _LIB.orc_sfc_expand.restype = C.c_int
        _LIB.orc_qp_solve.restype = C.c_int
        _LIB.orc_neighbours.restype = C.c_int
    return _LIB


def ref_lib():
    """The reference's own openGJK object code (oracle/_ref), or None when not built."""
    global _REF
    if _REF is None:
        so = os.path.join(_HERE, "_ref", "libopengjk_ref.so")
        if not os.path.exists(so):
            return None
        _REF = C.CDLL(so)
        _REF.ref_gjk_hull_origin.restype = C.c_double
    return _REF


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_params(M=5, n=5, phi=3, dim=3, use_sfc=False, dt=0.2, world_min=(-5, -5, 0),
                world_max=(5, 5, 2.5), world_res=0.1, grid_res=0.5, z_2d=1.0, comm_range=-1.0,
                w_control=0.01, w_terminal=1.0, reset_threshold=0.5):
    p = OrcParams()
    p.M, p.n, p.phi, p.dim, p.use_sfc = M, n, phi, dim, int(use_sfc)
    p.dt = dt
    for k in range(3):
        p.world_min[k] = float(np.float32(world_min[k]))
        p.world_max[k] = float(np.float32(world_max[k]))
    p.world_res, p.grid_res, p.z_2d = world_res, grid_res, z_2d
    p.comm_range, p.w_control, p.w_terminal = comm_range, w_control, w_terminal
    p.reset_threshold = reset_threshold
    return p


def gjk(points):
    pts = np.ascontiguousarray(points, dtype=np.float64)
    v = np.zeros(3)
    it = C.c_int(0)
    sn = C.c_int(0)
    d = lib().orc_gjk_hull_origin(_p(pts), C.c_int(pts.shape[0]), _p(v), C.byref(it), C.byref(sn))
    return d, v, it.value, sn.value


def ref_gjk(points):
    r = ref_lib()
    pts = np.ascontiguousarray(points, dtype=np.float64)
    v = np.zeros(3)
    sn = C.c_int(0)
    d = r.ref_gjk_hull_origin(_p(pts), C.c_int(pts.shape[0]), _p(v), C.byref(sn))
    return d, v, sn.value


def closest_segments(a0, a1, b0, b1):
    f = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    a0, a1, b0, b1 = f(a0), f(a1), f(b0), f(b1)
    p1 = np.zeros(3, np.float32)
    p2 = np.zeros(3, np.float32)
    d = C.c_double(0)
    lib().orc_closest_segments(_p(a0), _p(a1), _p(b0), _p(b1), _p(p1), _p(p2), C.byref(d))
    return p1, p2, d.value


class Edt:
    """EDT grid arrays + the C view."""

    def __init__(self, params, dist, obst, dims, min_key):
        self.dist = np.ascontiguousarray(dist, np.float32)
        self.obst = np.ascontiguousarray(obst, np.int32)
        self.dims = tuple(int(x) for x in dims)
        self.min_key = tuple(int(x) for x in min_key)
        self.res = params.world_res
        self.c = OrcEdt()
        for k in range(3):
            self.c.dims[k] = self.dims[k]
            self.c.min_key[k] = self.min_key[k]
        self.c.res = self.res
        self.c.dist = self.dist.ctypes.data
        self.c.obst = self.obst.ctypes.data


def edt_build(params, boxes):
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
    dims = (C.c_int32 * 3)()
    mk = (C.c_int32 * 3)()
    lib().orc_edt_dims(C.byref(params), dims, mk)
    nc = dims[0] * dims[1] * dims[2]
    dist = np.zeros(nc, np.float32)
    obst = np.zeros((nc, 3), np.int32)
    lib().orc_edt_build(C.byref(params), C.c_int(boxes.shape[0]), _p(boxes), _p(dist), _p(obst))
    return Edt(params, dist, obst, list(dims), list(mk))


def state_at(params, traj, t):
    traj = np.ascontiguousarray(traj, np.float32)
    st = np.zeros(9, np.float32)
    lib().orc_state_at(C.byref(params), _p(traj), C.c_double(t), _p(st))
    return st.reshape(3, 3)


def q_base(params):
    P = params.n + 1
    Q = np.zeros((P, P))
    lib().orc_q_base(C.byref(params), _p(Q))
    return Q


class Swarm:
    """Host-side state of a lock-step swarm driven through orc_step (mirrors the fields the
    reference keeps in AgentManager / TrajPlanner between replans)."""

    def __init__(self, params, start, goal, radius=0.15, downwash=2.0, max_vel=1.0, max_acc=2.0,
                 nominal_vel=1.0, edt=None, max_nbr=None, n_threads=1):
        self.p = params
        self.N = N = int(np.asarray(start).shape[0])
        M, P = params.M, params.n + 1
        f32 = lambda x: np.ascontiguousarray(x, np.float32)
        self.pos = f32(start).reshape(N, 3).copy()
        if params.dim == 2:
            self.pos[:, 2] = np.float32(params.z_2d)
        self.vel = np.zeros((N, 3), np.float32)
        self.acc = np.zeros((N, 3), np.float32)
        self.goal_des = f32(goal).reshape(N, 3).copy()
        if params.dim == 2:
            self.goal_des[:, 2] = np.float32(params.z_2d)
        self.goal_cur = self.pos.copy()        # agent_manager.cpp:9-10
        self.waypoint = self.pos.copy()
        self.disturbed = np.zeros(N, np.uint8)
        full = lambda v: np.full(N, v, np.float64) if np.isscalar(v) else np.ascontiguousarray(v, np.float64)
        self.radius, self.downwash = full(radius), full(downwash)
        self.max_vel, self.max_acc, self.nominal_vel = full(max_vel), full(max_acc), full(nominal_vel)
        self.edt = edt
        self.K = K = int(max_nbr if max_nbr is not None else max(N - 1, 1))
        self.n_threads = n_threads
        self.seq = 0
        self.traj = np.zeros((N, M, P, 3), np.float32)
        self.sfc = np.zeros((N, M, 6), np.float32)
        self.sfc_init = np.ones(N, np.uint8)
        self.init_traj = np.zeros((N, M, P, 3), np.float32)
        self.pred_traj = np.zeros((N, M, P, 3), np.float32)
        self.nbr_idx = np.zeros((N, K), np.int32)
        self.nbr_cnt = np.zeros(N, np.int32)
        self.lsc_normal = np.zeros((N, K, M, 3), np.float32)
        self.lsc_anchor = np.zeros((N, K, M, P, 3), np.float32)
        self.lsc_d = np.zeros((N, K, M, P), np.float64)
        self.qp_x = np.zeros((N, params.dim, M, P), np.float64)
        self.cost = np.zeros(N)
        self.max_violation = np.zeros(N)
        self.qp_iters = np.zeros(N, np.int32)
        self.status = np.zeros(N, np.int32)
        self.stage_seconds = np.zeros(5)
        self.comm_box = np.zeros((N, 6), np.float32)
        self.trap = np.zeros(N, np.uint8)
        self.set_obstacles(None)

    def set_obstacles(self, pos, vel=None, radius=None, downwash=None, max_acc=None, slack_weight=1.0,
                      size_prediction=True, uncertainty_horizon=1.0):
        """Dynamic (non-agent) obstacles of the next replans: Obstacle::position / velocity / radius / downwash /
        max_acc as the simulator's obstacle generator publishes them, plus opt/slack_collision_weight,
        obs/size_prediction, obs/uncertainty_horizon.  None removes them."""
        self.n_dyn = 0 if pos is None else int(np.asarray(pos).reshape(-1, 3).shape[0])
        nd = self.n_dyn
        self.dyn_pos = np.ascontiguousarray(np.zeros((0, 3)) if pos is None else pos, np.float32).reshape(nd, 3).copy()
        self.dyn_vel = np.ascontiguousarray(np.zeros((nd, 3)) if vel is None else vel, np.float32).reshape(nd, 3).copy()
        full = lambda v, dflt: np.full(nd, dflt if v is None else v, np.float64) if (v is None or np.isscalar(v)) \
            else np.ascontiguousarray(v, np.float64).copy()
        self.dyn_radius, self.dyn_downwash, self.dyn_max_acc = full(radius, 0.15), full(downwash, 1.0), full(max_acc, 0.0)
        self.slack_weight, self.size_prediction, self.uncertainty_horizon = float(slack_weight), bool(size_prediction), float(uncertainty_horizon)
        self.qp_slack = np.zeros((self.N, max(nd, 1), self.p.M), np.float64)

    def step(self, a_begin=None, a_end=None):
        """One replan of every agent (TrajPlanner::plan for all agents), or of agents [a_begin, a_end)."""
        self.seq += 1
        io = OrcStepIO()
        io.N, io.seq, io.max_nbr, io.n_threads = self.N, self.seq, self.K, self.n_threads
        for name in ("pos", "vel", "acc", "waypoint", "disturbed", "radius", "downwash", "max_vel",
                     "max_acc", "nominal_vel", "goal_cur", "sfc", "init_traj", "pred_traj", "nbr_idx",
                     "nbr_cnt", "lsc_normal", "lsc_anchor", "lsc_d", "qp_x", "cost", "max_violation",
                     "qp_iters", "status", "stage_seconds"):
            setattr(io, name, getattr(self, name).ctypes.data)
        io.prev_traj = self.traj.ctypes.data
        io.sfc_init_flag = self.sfc_init.ctypes.data
        io.edt = C.addressof(self.edt.c) if self.edt is not None else None
        io.n_dyn, io.dyn_size_prediction = self.n_dyn, int(self.size_prediction)
        io.dyn_uncertainty_horizon, io.slack_collision_weight = self.uncertainty_horizon, self.slack_weight
        for name in ("dyn_pos", "dyn_vel", "dyn_radius", "dyn_downwash", "dyn_max_acc", "comm_box", "qp_slack", "trap"):
            setattr(io, name, getattr(self, name).ctypes.data)
        if a_begin is None:
            lib().orc_step(C.byref(self.p), C.byref(io))
        else:
            lib().orc_step_range(C.byref(self.p), C.byref(io), C.c_int(a_begin), C.c_int(a_end))
        return self.status

    def advance(self):
        """AgentManager::doStep: move every agent to its trajectory state at t = dt."""
        for a in range(self.N):
            st = state_at(self.p, self.traj[a], self.p.dt)
            self.pos[a], self.vel[a], self.acc[a] = st[0], st[1], st[2]
            if self.p.dim == 2:
                self.pos[a, 2] = np.float32(self.p.z_2d)
