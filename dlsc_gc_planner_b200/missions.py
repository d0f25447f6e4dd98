"""Mission / world input formats of the reference and the synthetic swarm generators.

Host-side data loading only (numpy); mirrors what the reference's ``Mission::readMissionFile``
(reference src/mission.cpp:94-200) and ``MapManager::updateOctreeFromCSV`` (src/map_manager.cpp:264-316)
read.  Coordinates are parsed to float32 exactly like ``GetFloat()`` does (src/mission.cpp:111-112, 163-189).
"""
import json
import os
from dataclasses import dataclass, field

import numpy as np


@dataclass
class PlannerConfig:
    """Launch-file parameter sets (reference launch/*.launch)."""
    M: int = 5
    n: int = 5
    phi: int = 3
    dim: int = 3
    use_sfc: bool = False
    dt: float = 0.2
    world_res: float = 0.1
    grid_res: float = 0.5
    z_2d: float = 1.0
    comm_range: float = -1.0
    w_control: float = 0.01
    w_terminal: float = 1.0
    reset_threshold: float = 0.5

    @staticmethod
    def empty():      # launch/testall_DLSCGC_empty.launch
        return PlannerConfig(M=5, dim=3, use_sfc=False, comm_range=-1.0)

    @staticmethod
    def forest3d():   # launch/testall_DLSCGC_3D.launch
        return PlannerConfig(M=10, dim=3, use_sfc=True, comm_range=3.0)

    @staticmethod
    def maze2d():     # launch/simulation.launch
        return PlannerConfig(M=10, dim=2, use_sfc=True, comm_range=3.0)


@dataclass
class Mission:
    world_min: np.ndarray
    world_max: np.ndarray
    start: np.ndarray          # [N,3] float32
    goal: np.ndarray           # [N,3] float32
    radius: np.ndarray         # [N] float64
    downwash: np.ndarray
    max_vel: np.ndarray
    max_acc: np.ndarray
    nominal_vel: np.ndarray
    boxes: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))  # cx,cy,cz,sx,sy,sz
    obstacles: list = field(default_factory=list)   # mission JSON "obstacles" entries (dynamic obstacles), see obstacle_states

    @property
    def n_agents(self):
        return int(self.start.shape[0])


def load_mission(path, dim=3, z_2d=1.0):
    """Parse a reference mission JSON (missions/readme.txt)."""
    with open(path) as f:
        doc = json.load(f)
    d = doc["world"][0]["dimension"]
    wmin = np.array(d[:3], np.float32)
    wmax = np.array(d[3:], np.float32)
    quads = doc["quadrotors"]
    start, goal, rad, dw, mv, ma, nv = [], [], [], [], [], [], []
    for ag in doc["agents"]:
        q = quads[ag["type"]]
        s = [float(x) for x in ag["start"]]
        g = [float(x) for x in ag["goal"]]
        if dim == 2:
            s[2] = z_2d
            g[2] = z_2d
        start.append(s)
        goal.append(g)
        rad.append(float(ag.get("radius", q["radius"])) if "size" in ag else float(q["radius"]))
        dw.append(float(ag.get("downwash", q["downwash"])))
        mv.append(float(q["max_vel"][0]))          # first element only, src/mission.cpp:120-124
        ma.append(float(q["max_acc"][0]))
        nv.append(float(q["nominal_velocity"]))
    f64 = lambda x: np.array(x, np.float64)
    return Mission(wmin, wmax, np.array(start, np.float32), np.array(goal, np.float32), f64(rad), f64(dw),
                   f64(mv), f64(ma), f64(nv), obstacles=list(doc.get("obstacles", [])))


SPIN4 = [dict(type="spin", axis_position=[0.0, 0.0, 1.0], axis_ori=[0, 0, 1], start=s, speed=1.0, size=0.3, max_acc=2.0,
              downwash=1.0) for s in ([2.0, 0.0, 1.0], [0.0, 2.0, 1.0], [-2.0, 0.0, 1.0], [0.0, -2.0, 1.0])]
"""the "obstacles" block of the reference's missions/forest10_spin4_* and maze10_tro2022_spin4_* files"""


def obstacle_states(obstacles, t):
    """Scenario generation for the dynamic-obstacle path: the states ObstacleGenerator::update(t) publishes
    (reference include/obstacle_generator.hpp:66-91) for the mission JSON's "spin" and "straight" entries
    (SpinObstacle / StraightObstacle, include/obstacle.hpp:96-152, 154-245; parsed like src/mission.cpp:209-259).
    -> dict(pos [n,3] f32, vel [n,3] f32, radius, downwash, max_acc [n] f64): the arguments of
    SwarmPlanner.set_obstacles.  The other motion models (patrol: replanned by the simulator; chasing, gaussian, real)
    are experiment tooling and are not restated."""
    pos, vel, rad, dw, ma = [], [], [], [], []
    for o in obstacles:
        kind = o["type"]
        if kind == "spin":
            a = np.array(o["start"], np.float64) - np.array(o["axis_position"], np.float64)
            n = np.array(o["axis_ori"], np.float64)
            n = n / np.linalg.norm(n)
            r = a - a.dot(n) * n
            w = o["speed"] / np.linalg.norm(r)
            th = w * t
            rot = lambda v, ang: v * np.cos(ang) + np.cross(n, v) * np.sin(ang) + n * n.dot(v) * (1 - np.cos(ang))
            p = rot(a, th)                                      # q p q^-1
            pos.append(np.array(o["axis_position"], np.float64) + p)
            vel.append(w * rot(p, np.pi / 2))                   # the reference rotates the full offset, not its radial part
        elif kind == "straight":
            s0, g = np.array(o["start"], np.float32), np.array(o["goal"], np.float32)
            speed, amax = float(o["speed"]), float(o["max_acc"])
            dist = float(np.linalg.norm((g - s0).astype(np.float64)))
            nrm = ((g - s0) / np.float32(dist)).astype(np.float64)
            dacc = 0.5 * speed * speed / amax
            p, v = g.astype(np.float64), np.zeros(3)
            if dist > 2 * dacc:
                t1 = speed / amax; t2 = t1 + (dist - 2 * dacc) / speed; t3 = t1 + t2
                if t < t1: p, v = s0 + nrm * 0.5 * amax * t * t, nrm * amax * t
                elif t < t2: p, v = s0 + nrm * (0.5 * amax * t1 * t1 + speed * (t - t1)), nrm * speed
                elif t < t3: p, v = g - nrm * 0.5 * amax * (t3 - t) ** 2, nrm * (speed - amax * (t - t2))
            else:
                t1 = np.sqrt(dist / amax); t2 = 2 * t1
                if t < t1: p, v = s0 + nrm * 0.5 * amax * t * t, nrm * amax * t
                elif t < t2: p, v = s0 + nrm * (0.5 * dist + amax * t1 * (t - t1) - 0.5 * amax * (t - t1) ** 2), nrm * amax * (t2 - t)
            pos.append(p); vel.append(v)
        else:
            raise NotImplementedError("obstacle type %r: only 'spin' and 'straight' are restated" % kind)
        rad.append(float(o["size"])); ma.append(float(o["max_acc"]))
        dw.append(float(o["downwash"]) if float(o.get("downwash", 0)) != 0 else 1.0)
    f64 = lambda x: np.array(x, np.float64)
    return dict(pos=np.array(pos, np.float32).reshape(-1, 3), vel=np.array(vel, np.float32).reshape(-1, 3),
                radius=f64(rad), downwash=f64(dw), max_acc=f64(ma))


def add_goal_noise(mission, max_noise, dim=3, seed=0):
    """Mission::addNoise (reference src/mission.cpp:395-406): desired_goal(k) += (float)(U(0,1) * max_noise) for
    the first `dim` coordinates -- seeded here (the reference draws from std::random_device) so that Monte-Carlo
    batches are reproducible: seed = mission index."""
    rng = np.random.default_rng(seed)
    goal = mission.goal.astype(np.float32).copy()
    u = rng.random((mission.n_agents, dim)).astype(np.float32)
    goal[:, :dim] = goal[:, :dim] + (u.astype(np.float64) * max_noise).astype(np.float32)
    return Mission(mission.world_min, mission.world_max, mission.start, goal, mission.radius, mission.downwash,
                   mission.max_vel, mission.max_acc, mission.nominal_vel, mission.boxes, mission.obstacles)


def concat_missions(ms):
    """Independent missions of one world side by side in one agent array (Monte-Carlo batch); returns the batch
    and the mission index of every agent (for dlsc_set_groups)."""
    cat = lambda f: np.concatenate([getattr(m, f) for m in ms])
    batch = Mission(ms[0].world_min, ms[0].world_max, cat("start"), cat("goal"), cat("radius"), cat("downwash"),
                    cat("max_vel"), cat("max_acc"), cat("nominal_vel"), ms[0].boxes)
    group = np.concatenate([np.full(m.n_agents, i, np.int32) for i, m in enumerate(ms)])
    return batch, group


def load_world_csv(path):
    """Rows ``cx,cy,cz,sx,sy,sz`` (src/map_manager.cpp:267-283; tokens go through stod then float)."""
    rows = []
    with open(path) as f:
        for line in f:
            tok = [t for t in line.strip().split(",") if t != ""]
            if len(tok) < 6:
                continue
            rows.append([float(t) for t in tok[:6]])
    return np.array(rows, np.float32).reshape(-1, 6)


def load_world_bt(path):
    """Parse an octomap binary tree (.bt), the other world format MapManager::setGlobalMap accepts
    (reference src/map_manager.cpp:66-73: OcTree::readBinary, then expand()).  octomap itself is absent here; the
    format is the published one of octomap 1.9 (OcTreeBase::readBinary / OcTree::readBinaryNode): a text header
    ("id OcTree", "size <nodes>", "res <m>", "data"), then the tree depth-first, two bytes per inner node = 2 bits per
    child, bit pair (b[2i], b[2i+1]) = (1,0) free leaf, (0,1) occupied leaf, (1,1) inner node, (0,0) unknown; child i
    sits at offset (i & 1, i >> 1 & 1, i >> 2 & 1); depth 16, key 32768 = coordinate 0.
    -> (res, cubes) with cubes [n, 4] int64 = (kx, ky, kz, edge) per OCCUPIED leaf in cells relative to the origin
    (cell k covers [k res, (k+1) res)); a pruned leaf above the finest depth is one cube with edge > 1.
    Raises ValueError when the node count differs from the header (a cheap whole-file check of the reading)."""
    data = open(path, "rb").read()
    end = data.index(b"\ndata\n") + 6
    head = {}
    for line in data[:end].decode("ascii", "replace").splitlines():
        parts = line.split()
        if len(parts) == 2 and not line.startswith("#"):
            head[parts[0]] = parts[1]
    if head.get("id") != "OcTree":
        raise ValueError("not an OcTree binary file: id %r" % head.get("id"))
    res, n_nodes = float(head["res"]), int(head["size"])
    pos, count, cubes = end, 1, []
    stack = [(0, 0, 0, 65536)]                      # inner nodes still to read, depth-first in child order
    while stack:
        x0, y0, z0, edge = stack.pop()
        if pos + 2 > len(data):
            raise ValueError("truncated .bt file")
        bits = data[pos] | (data[pos + 1] << 8)
        pos += 2
        h = edge // 2
        inner = []
        for i in range(8):
            b0, b1 = (bits >> (2 * i)) & 1, (bits >> (2 * i + 1)) & 1
            if not (b0 or b1):
                continue
            count += 1
            cx, cy, cz = x0 + (i & 1) * h, y0 + ((i >> 1) & 1) * h, z0 + ((i >> 2) & 1) * h
            if b0 and b1:
                inner.append((cx, cy, cz, h))
            elif b1:
                cubes.append((cx - 32768, cy - 32768, cz - 32768, h))
        stack.extend(reversed(inner))               # readBinaryNode recurses into the inner children in index order
    if count != n_nodes:
        raise ValueError(".bt node count %d != header size %d" % (count, n_nodes))
    return res, np.array(cubes, np.int64).reshape(-1, 4)


def occupancy_from_cubes(cubes, dims, min_key):
    """Occupancy grid [dims0][dims1][dims2] uint8 of the planner's map window (dlsc_edt_dims: map index = cell key -
    min_key) from load_world_bt cubes -- what dlsc_build_edt_occupancy takes.  Cells outside the window are dropped, like
    the bounding-box iteration of DynamicEDTOctomap (reference src/map_manager.cpp:75-79)."""
    occ = np.zeros(tuple(int(d) for d in dims), np.uint8)
    for kx, ky, kz, e in np.asarray(cubes, np.int64):
        lo = np.array([kx, ky, kz]) - np.asarray(min_key, np.int64)
        hi = lo + e
        lo = np.maximum(lo, 0); hi = np.minimum(hi, occ.shape)
        if np.all(hi > lo):
            occ[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = 1
    return occ


def lattice_step_waypoints(pos, goal, grid_res=0.5, occupied=None):
    """Documented stand-in for the PIBT waypoint provider (reference src/grid_based_planner.cpp:64-94):
    one greedy step on the ``grid_res`` lattice from the current waypoint toward the goal, axis with the
    largest remaining distance first, skipping lattice nodes in ``occupied`` (a set of integer node keys).
    Identical waypoints are fed to the oracle and to the GPU path, so parity is independent of it."""
    pos = np.asarray(pos, np.float32)
    goal = np.asarray(goal, np.float32)
    out = pos.copy()
    for a in range(pos.shape[0]):
        delta = goal[a].astype(np.float64) - pos[a].astype(np.float64)
        order = np.argsort(-np.abs(delta))
        for k in order:
            if abs(delta[k]) < 0.5 * grid_res:
                continue
            cand = pos[a].astype(np.float64).copy()
            cand[k] += np.sign(delta[k]) * grid_res
            key = tuple(int(round(c / grid_res)) for c in cand)
            if occupied is not None and key in occupied:
                continue
            out[a] = cand.astype(np.float32)
            break
    return out


def synthetic_forest(n_agents=4096, half_extent=32.0, height=2.5, tree_density=0.1, seed=4096,
                     grid_res=0.5, z=1.0):
    """square world [-h,h]^2 x [0,height], trees 0.5x0.5xheight boxes at
    ``tree_density`` per m^2, distinct start / goal lattice nodes (0.5 m xy lattice at height z) that keep
    1.0 m (xy, Chebyshev) clearance from every tree centre.  Deviation from the survey text: numpy's PCG64
    replaces std::mt19937_64 (no bit-compatible generator in numpy); seeds are fixed so every run,
    oracle or GPU, sees the same world."""
    rng = np.random.default_rng(seed)
    h = float(half_extent)
    n_trees = int(round(tree_density * (2 * h) ** 2 / 10.0) * 10) if tree_density > 0 else 0
    n_trees = int(round(tree_density * (2 * h) ** 2))
    nodes_1d = np.arange(-h + 1.0, h - 1.0 + 1e-9, grid_res)
    gx, gy = np.meshgrid(nodes_1d, nodes_1d, indexing="ij")
    nodes = np.stack([gx.ravel(), gy.ravel()], 1)
    # trees on the 0.1 m grid so the box edges are cell aligned
    tc = np.round(rng.uniform(-h + 1.0, h - 1.0, size=(n_trees, 2)) * 10.0) / 10.0 + 0.05
    # nodes with clearance
    free = np.ones(nodes.shape[0], bool)
    for t in range(n_trees):
        free &= np.max(np.abs(nodes - tc[t]), axis=1) >= 1.0
    free_idx = np.nonzero(free)[0]
    if free_idx.size < n_agents:
        raise ValueError("world too small for %d agents" % n_agents)
    rng2 = np.random.default_rng(seed + 1)
    s_idx = rng2.choice(free_idx, n_agents, replace=False)
    g_idx = rng2.choice(free_idx, n_agents, replace=False)
    start = np.concatenate([nodes[s_idx], np.full((n_agents, 1), z)], 1).astype(np.float32)
    goal = np.concatenate([nodes[g_idx], np.full((n_agents, 1), z)], 1).astype(np.float32)
    boxes = np.zeros((n_trees, 6), np.float32)
    boxes[:, 0:2] = tc
    boxes[:, 2] = height / 2
    boxes[:, 3:5] = 0.5
    boxes[:, 5] = height
    full = lambda v: np.full(n_agents, v, np.float64)
    return Mission(np.array([-h, -h, 0.0], np.float32), np.array([h, h, height], np.float32), start, goal,
                   full(0.15), full(2.0), full(1.0), full(2.0), full(1.0), boxes)


def synthetic_empty(n_agents=70, half_extent=None, seed=7, grid_res=0.5):
    """Obstacle-free 3-D swarm shaped like missions/empty*/multi_random_*agents_*.json: distinct start and
    goal nodes on the 0.5/0.5/1.0 m lattice of a box world."""
    rng = np.random.default_rng(seed)
    if half_extent is None:
        half_extent = max(1.5, 0.5 * np.ceil(np.cbrt(n_agents * 4.0)))
    h = float(half_extent)
    xs = np.arange(-h + 0.5, h - 0.5 + 1e-9, grid_res)
    zs = np.arange(0.5, 2.0 + 1e-9, 1.0) if h <= 2.0 else np.arange(0.5, 2 * h - 0.5 + 1e-9, 1.0)
    gx, gy, gz = np.meshgrid(xs, xs, zs, indexing="ij")
    nodes = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1)
    if nodes.shape[0] < n_agents:
        raise ValueError("world too small")
    s = nodes[rng.choice(nodes.shape[0], n_agents, replace=False)]
    g = nodes[rng.choice(nodes.shape[0], n_agents, replace=False)]
    full = lambda v: np.full(n_agents, v, np.float64)
    zmax = 2.5 if h <= 2.0 else 2 * h
    return Mission(np.array([-h, -h, 0.0], np.float32), np.array([h, h, zmax], np.float32),
                   s.astype(np.float32), g.astype(np.float32), full(0.15), full(2.0), full(1.0), full(2.0),
                   full(1.0))


def occupied_nodes(boxes, grid_res=0.5, inflate=0.2):
    """Integer (ix, iy) keys of the ``grid_res`` lattice nodes that lie inside an obstacle box inflated by
    ``inflate`` metres in x/y (what the reference's GridBasedPlanner marks as blocked from the EDT,
    src/grid_based_planner.cpp:94-164, restated geometrically)."""
    occ = set()
    for b in np.asarray(boxes, np.float64).reshape(-1, 6):
        lo = b[0:2] - 0.5 * b[3:5] - inflate
        hi = b[0:2] + 0.5 * b[3:5] + inflate
        i0, i1 = int(np.ceil(lo[0] / grid_res - 1e-9)), int(np.floor(hi[0] / grid_res + 1e-9))
        j0, j1 = int(np.ceil(lo[1] / grid_res - 1e-9)), int(np.floor(hi[1] / grid_res + 1e-9))
        for i in range(i0, i1 + 1):
            for j in range(j0, j1 + 1):
                occ.add((i, j))
    return occ


class LatticeRouter:
    """Per-agent breadth-first distance-to-goal fields on the x/y lattice (4-connected, obstacle nodes
    removed).  Stand-in for the single-agent part of the reference's grid planner; used for small swarms
    in walled worlds (maze), where a greedy step toward the goal would stop at the first wall."""

    def __init__(self, world_min, world_max, grid_res, occupied, goals):
        g = float(grid_res)
        self.g = g
        self.i0 = int(np.ceil(float(world_min[0]) / g - 1e-9))
        self.j0 = int(np.ceil(float(world_min[1]) / g - 1e-9))
        self.ni = int(np.floor(float(world_max[0]) / g + 1e-9)) - self.i0 + 1
        self.nj = int(np.floor(float(world_max[1]) / g + 1e-9)) - self.j0 + 1
        free = np.ones((self.ni, self.nj), bool)
        for (i, j) in (occupied or ()):
            if 0 <= i - self.i0 < self.ni and 0 <= j - self.j0 < self.nj:
                free[i - self.i0, j - self.j0] = False
        self.free = free
        self.fields = []
        for q in np.asarray(goals, np.float64):
            self.fields.append(self._bfs(int(round(q[0] / g)) - self.i0, int(round(q[1] / g)) - self.j0))

    def _bfs(self, gi, gj):
        INF = 1 << 30
        d = np.full((self.ni, self.nj), INF, np.int64)
        if not (0 <= gi < self.ni and 0 <= gj < self.nj):
            return d
        d[gi, gj] = 0
        frontier = [(gi, gj)]
        while frontier:
            nxt = []
            for (i, j) in frontier:
                for di, dj in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                    a, b = i + di, j + dj
                    if 0 <= a < self.ni and 0 <= b < self.nj and self.free[a, b] and d[a, b] == INF:
                        d[a, b] = d[i, j] + 1
                        nxt.append((a, b))
            frontier = nxt
        return d

    def candidates(self, agent, cur):
        """Lattice neighbours of `cur` ordered by distance-to-goal (closer first), only improving ones."""
        g, d = self.g, self.fields[agent]
        i, j = int(round(cur[0] / g)) - self.i0, int(round(cur[1] / g)) - self.j0
        here = d[i, j] if (0 <= i < self.ni and 0 <= j < self.nj) else 1 << 30
        out = []
        for di, dj in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            a, b = i + di, j + dj
            if 0 <= a < self.ni and 0 <= b < self.nj and d[a, b] < here:
                c = np.array(cur, np.float64)
                c[0] += di * g
                c[1] += dj * g
                out.append((d[a, b], len(out), c))
        out.sort(key=lambda t: (t[0], t[1]))
        return [c for _, _, c in out]


def next_waypoints(waypoint, goal_cur, goal_des, traj, pos, cfg, occupied=None, router=None):
    """Documented stand-in for the reference's waypoint provider (PIBT on the lattice + the update rules of
    MultiSyncSimulator::decentralizedMAPP, src/multi_sync_simulator.cpp:308-466).  It keeps the update
    rules that matter to the hot path and replaces the multi-agent search by a single-agent step:
      * a waypoint only moves once the agent's current goal point has reached it (:407-413);
      * the new waypoint is one lattice step (grid_res in x/y, 2*grid_res in z) from the old one: along the
        breadth-first shortest path when a `router` is given, else toward the desired goal, largest
        remaining axis first; never onto an obstacle node or another agent's waypoint (:418-447);
      * with a communication range R it must stay within R/2 - 1e-5 (Chebyshev) of every segment start
        point and of the end point of the agent's current trajectory (:386-404).
    The same waypoints are fed to the oracle and to the GPU path, so parity does not depend on it."""
    wp = np.asarray(waypoint, np.float32).copy()
    goal_cur = np.asarray(goal_cur, np.float32)
    goal_des = np.asarray(goal_des, np.float32)
    N = wp.shape[0]
    g = float(cfg.grid_res)
    steps = np.array([g, g, 2.0 * g])
    key = lambda q: (int(round(q[0] / g)), int(round(q[1] / g)), int(round(q[2] / g)))
    taken = {key(wp[a].astype(np.float64)): a for a in range(N)}
    reached = np.max(np.abs(goal_cur - wp), axis=1) < 1e-4
    R = float(cfg.comm_range)
    for a in np.nonzero(reached)[0]:
        cur = wp[a].astype(np.float64)
        delta = goal_des[a].astype(np.float64) - cur
        cands = []
        if router is not None:
            cands = router.candidates(a, cur)
        else:
            naxes = 3 if cfg.dim == 3 else 2
            for k in np.argsort(-np.abs(delta[:naxes]), kind="stable"):
                if abs(delta[k]) < 0.5 * steps[k]:
                    continue
                c = cur.copy()
                c[k] += np.sign(delta[k]) * steps[k]
                cands.append(c)
        for cand in cands:
            if occupied is not None and (int(round(cand[0] / g)), int(round(cand[1] / g))) in occupied:
                continue
            kc = key(cand)
            if kc in taken and taken[kc] != a:
                continue
            if R > 0:
                if traj is None:
                    pts = np.asarray(pos[a], np.float64)[None, :]
                else:
                    pts = np.concatenate([traj[a][:, 0, :], traj[a][-1:, -1, :]], 0).astype(np.float64)
                if np.max(np.abs(pts - cand[None, :])) > 0.5 * R - 1e-5:
                    break
            taken.pop(key(cur), None)
            taken[kc] = a
            wp[a] = cand.astype(np.float32)
            break
    return wp
