"""Result log in the reference's on-disk format (MultiSyncSimulator::saveResultAsCSV, reference
src/multi_sync_simulator.cpp:735-851), the de-facto interchange with the reference's replayer
(src/multi_sync_replayer.cpp): one row per recorded time, per agent the twelve columns
``id,t,px,py,pz,vx,vy,vz,ax,ay,az,planning_time``, numbers as C++ ``ostream << double / float`` prints them
(6 significant digits, %g); when the mission has dynamic obstacles their six columns ``obs_id,t,px,py,pz,size`` follow
the agents (:755-763, 829-843)."""
import numpy as np

AGENT_COLUMNS = "id,t,px,py,pz,vx,vy,vz,ax,ay,az,planning_time"
OBSTACLE_COLUMNS = "obs_id,t,px,py,pz,size"


def _g(x):
    """C++ default stream formatting of a double / float: precision 6, %g (a negative zero prints as "-0", as
    `ostream << -0.0f` does)."""
    return "%g" % float(x)


def header(n_agents, n_obstacles=0):
    return ",".join([AGENT_COLUMNS] * n_agents + [OBSTACLE_COLUMNS] * n_obstacles)


def format_row(t, pos, vel, acc, planning_time, obs_pos=None, obs_radius=None):
    """pos / vel / acc [N][3] (float32 as point3d), planning_time [N] seconds; obs_pos [n_obs][3] (float32), obs_radius
    [n_obs] (double) of the dynamic obstacles, if any."""
    cells = []
    for a in range(len(pos)):
        cells.append(str(a))
        cells.append(_g(t))
        for v in (pos[a], vel[a], acc[a]):
            cells.extend(_g(np.float32(c)) for c in v)
        cells.append(_g(planning_time[a]))
    for o in range(0 if obs_pos is None else len(obs_pos)):
        cells.append(str(o))
        cells.append(_g(t))
        cells.extend(_g(np.float32(c)) for c in obs_pos[o])
        cells.append(_g(obs_radius[o]))
    return ",".join(cells)


class ResultLog:
    """Append-only writer: `log.record(t, pos, vel, acc, planning_time[, obs_pos, obs_radius])` once per recorded time."""

    def __init__(self, path, n_agents, n_obstacles=0):
        self.n, self.on = int(n_agents), int(n_obstacles)
        self.f = open(path, "w")
        self.f.write(header(self.n, self.on) + "\n")

    def record(self, t, pos, vel, acc, planning_time=None, obs_pos=None, obs_radius=None):
        pt = np.zeros(self.n) if planning_time is None else planning_time
        assert (0 if obs_pos is None else len(obs_pos)) == self.on
        self.f.write(format_row(t, pos, vel, acc, pt, obs_pos, obs_radius) + "\n")

    def close(self):
        self.f.close()


def read(path, with_obstacles=False):
    """-> t [T], pos / vel / acc [T][N][3] float32, planning_time [T][N] -- what MultiSyncReplayer::readCSVFile parses
    (reference src/multi_sync_replayer.cpp:53-112: agents counted by the "id" header cells, obstacles by "obs_id");
    with_obstacles adds obs_pos [T][n_obs][3] float32 and obs_radius [T][n_obs]."""
    with open(path) as f:
        lines = [l.strip() for l in f if l.strip()]
    cells = lines[0].split(",")
    n, on = cells.count("id"), cells.count("obs_id")
    data = np.array([[float(c) for c in l.split(",")] for l in lines[1:]], np.float64)
    rows = data[:, :12 * n].reshape(len(lines) - 1, n, 12)
    out = (rows[:, 0, 1], rows[:, :, 2:5].astype(np.float32), rows[:, :, 5:8].astype(np.float32),
           rows[:, :, 8:11].astype(np.float32), rows[:, :, 11])
    if not with_obstacles:
        return out
    obs = data[:, 12 * n:12 * n + 6 * on].reshape(len(lines) - 1, on, 6)
    return out + (obs[:, :, 2:5].astype(np.float32), obs[:, :, 5])


# ---- summary file (MultiSyncSimulator::saveSummarizedResultAsCSV, reference src/multi_sync_simulator.cpp:852-900) ----
SUMMARY_COLUMNS = ("start_time,total_flight_time,total_flight_distance,safety_ratio_agent,safety_ratio_obs,"
                   "mapf_time_average,mapf_time_min,mapf_time_max,planning_time_average,planning_time_min,planning_time_max,"
                   "initial_traj_planning_time,obstacle_prediction_time,goal_planning_time,lsc_generation_time,"
                   "sfc_generation_time,traj_optimization_time,mission_file_name,world_file_name,planner_mode,goal_mode,mapf_mode,"
                   "communication_range,world_dimension,M,dt")


def total_flight_distance(pos):
    """getTotalDistance (:902-911): sum over agents of the polyline length through the recorded positions [T][N][3]
    (point3d differences are float, every norm is sqrt of a float sum of squares)."""
    p = np.asarray(pos, np.float32)
    d = (p[1:] - p[:-1]).astype(np.float32)
    nsq = ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32) + d[..., 2] * d[..., 2]).astype(np.float32)
    return float(np.sqrt(nsq.astype(np.float64)).sum())


def safety_ratio_agents(pos, radius, downwash):
    """Minimum over recorded times and agent pairs of the downwash-scaled distance / (r_i + r_j)
    (:653-674, ellipsoidalDistance include/util.hpp:163-167)."""
    p = np.asarray(pos, np.float32)
    r = np.asarray(radius, np.float64); dw = np.asarray(downwash, np.float64)
    best = 1e9                                       # SP_INFINITY is the initial value the reference prints when nothing is closer
    for i in range(p.shape[1]):
        for j in range(p.shape[1]):
            if i == j:
                continue
            k = (dw[i] * r[i] + dw[j] * r[j]) / (r[i] + r[j])
            d = (p[:, i] - p[:, j]).astype(np.float32)
            d[:, 2] = (d[:, 2].astype(np.float64) / k).astype(np.float32)
            nsq = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2]).astype(np.float32)
            best = min(best, float(np.sqrt(nsq.astype(np.float64)).min() / (r[i] + r[j])))
    return best


def format_summary(start_time, total_flight_time, total_distance, safety_agent, safety_obs, mapf, planning, stage_means,
                   mission_file, world_file, planner_mode="DLSCGC", goal_mode="grid_based_planner", mapf_mode="pibt",
                   communication_range=3, world_dimension=2, M=10, dt=0.2):
    """One row.  start_time is written as the reference's string (ros::Time as "sec.usec"); mapf / planning = (average, min,
    max) seconds; stage_means = averages of (initial_traj, obstacle_prediction, goal, lsc, sfc, traj_optimization) seconds."""
    cells = [str(start_time), _g(total_flight_time), _g(total_distance), _g(safety_agent), _g(safety_obs)]
    cells += [_g(x) for x in mapf] + [_g(x) for x in planning] + [_g(x) for x in stage_means]
    cells += [mission_file, world_file, planner_mode, goal_mode, mapf_mode, _g(communication_range), str(int(world_dimension)),
              str(int(M)), _g(dt)]
    return ",".join(cells)


def append_summary(path, row):
    """Appends a row, writing the column line first when the file is new or empty (as the reference does)."""
    import os
    new = not os.path.exists(path) or os.path.getsize(path) == 0
    with open(path, "a") as f:
        if new:
            f.write(SUMMARY_COLUMNS + "\n")
        f.write(row + "\n")
