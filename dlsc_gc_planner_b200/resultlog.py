"""Result log in the reference's on-disk format (MultiSyncSimulator::saveResultAsCSV, reference
src/multi_sync_simulator.cpp:735-851), the de-facto interchange with the reference's replayer
(src/multi_sync_replayer.cpp): one row per recorded time, per agent the twelve columns
``id,t,px,py,pz,vx,vy,vz,ax,ay,az,planning_time``, numbers as C++ ``ostream << double / float`` prints them
(6 significant digits, %g).  Agents only (obstacle columns ``obs_id,t,px,py,pz,size`` follow the agents in the
reference when a mission has dynamic obstacles: not on this path yet)."""
import numpy as np

AGENT_COLUMNS = "id,t,px,py,pz,vx,vy,vz,ax,ay,az,planning_time"


def _g(x):
    """C++ default stream formatting of a double / float: precision 6, %g (a negative zero prints as "-0", as
    `ostream << -0.0f` does)."""
    return "%g" % float(x)


def header(n_agents):
    return ",".join([AGENT_COLUMNS] * n_agents)


def format_row(t, pos, vel, acc, planning_time):
    """pos / vel / acc [N][3] (float32 as point3d), planning_time [N] seconds."""
    cells = []
    for a in range(len(pos)):
        cells.append(str(a))
        cells.append(_g(t))
        for v in (pos[a], vel[a], acc[a]):
            cells.extend(_g(np.float32(c)) for c in v)
        cells.append(_g(planning_time[a]))
    return ",".join(cells)


class ResultLog:
    """Append-only writer: `log.record(t, pos, vel, acc, planning_time)` once per recorded time."""

    def __init__(self, path, n_agents):
        self.n = int(n_agents)
        self.f = open(path, "w")
        self.f.write(header(self.n) + "\n")

    def record(self, t, pos, vel, acc, planning_time=None):
        pt = np.zeros(self.n) if planning_time is None else planning_time
        self.f.write(format_row(t, pos, vel, acc, pt) + "\n")

    def close(self):
        self.f.close()


def read(path):
    """-> t [T], pos / vel / acc [T][N][3] float32, planning_time [T][N] (what the replayer parses)."""
    with open(path) as f:
        lines = [l.strip() for l in f if l.strip()]
    n = lines[0].count("id,t,")
    rows = np.array([[float(c) for c in l.split(",")] for l in lines[1:]], np.float64).reshape(len(lines) - 1, n, 12)
    return (rows[:, 0, 1], rows[:, :, 2:5].astype(np.float32), rows[:, :, 5:8].astype(np.float32),
            rows[:, :, 8:11].astype(np.float32), rows[:, :, 11])
