// dlsc_compat.hpp -- the reference's per-agent C++ classes, source compatible, backed by libdlsc_b200.so.
//
// Drop-in for the three hot-path classes of dlsc_gc_planner (namespace MATP):
//   TrajPlanner            reference include/traj_planner.hpp:40-60, src/traj_planner.cpp:4-63
//   TrajOptimizer          reference include/traj_optimizer.hpp:18-33, src/traj_optimizer.cpp:18-165
//   CollisionConstraints   reference include/collision_constraints.hpp:116-171
// with the same constructor / method signatures, value-type results and error behaviour
// (std::invalid_argument for bad configuration, `throw PlanningReport::QPFAILED` for solver failures that
// the reference lets escape, the initial-trajectory failsafe inside TrajPlanner::plan).
//
// How the unchanged per-agent call pattern is batched (AgentManager / MultiSyncSimulator need no edits):
//   * every TrajPlanner of a mission shares one device context (SwarmBatch, keyed by the mission size);
//   * MultiSyncSimulator::broadcastMsgs calls setObstacles() on every agent before the first plan() of a
//     step (src/multi_sync_simulator.cpp:468-536).  setObstacles() marks a new step;
//   * plan(agent, ...) uploads that agent's state and replans it on the GPU.  If the driver announces all
//     agents of the step up front with the optional  TrajPlanner::stageAgent(agent, disturbed)  (one added
//     loop in MultiSyncSimulator::plan), the first plan() of the step launches ONE
//     batched dlsc_step() for the whole swarm and the remaining plan() calls return their cached slice.
//
// When ROS / octomap / dynamicEDT3D / Eigen headers are present the real types are used; otherwise
// (this repository's tests) minimal stand-ins with the same names are provided below
// (#define DLSC_COMPAT_STANDALONE, default when <ros/ros.h> is not found).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "dlsc_b200.h"

#if !defined(DLSC_COMPAT_STANDALONE) && defined(__has_include)
#if !__has_include(<ros/ros.h>)
#define DLSC_COMPAT_STANDALONE 1
#endif
#endif

#ifdef DLSC_COMPAT_STANDALONE
// ---------------------------------------------------------------------------------------------------
// stand-ins for the third-party types that appear in the reference signatures
// ---------------------------------------------------------------------------------------------------
namespace ros {
struct NodeHandle {};
struct Time {
    double t = 0;
    static Time now() { return Time(); }
    double toSec() const { return t; }
};
}  // namespace ros
namespace octomap {
struct point3d {          // octomath::Vector3: float storage
    float v[3] = {0.f, 0.f, 0.f};
    point3d() = default;
    point3d(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
    float& x() { return v[0]; } float& y() { return v[1]; } float& z() { return v[2]; }
    float x() const { return v[0]; } float y() const { return v[1]; } float z() const { return v[2]; }
    float& operator()(unsigned i) { return v[i]; }
    float operator()(unsigned i) const { return v[i]; }
    point3d operator-(const point3d& o) const { return point3d(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    point3d operator+(const point3d& o) const { return point3d(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    point3d operator*(float s) const { return point3d(v[0] * s, v[1] * s, v[2] * s); }
    double norm() const { return std::sqrt((double)(float)(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])); }
};
struct OcTree {};
}  // namespace octomap
// DynamicEDTOctomap stand-in: owns the distance grid arrays the accessor getDistanceAndClosestObstacle serves
class DynamicEDTOctomap {
public:
    std::vector<float> dist;          // [ncell] metres
    std::vector<int32_t> obst;        // [ncell][3]
    int32_t dims[3] = {0, 0, 0}, min_key[3] = {0, 0, 0};
    double res = 0.1;
    // alternative: let the library build the grid on the GPU from the mission's obstacle CSV rows
    // (MapManager::updateOctreeFromCSV + setGlobalMap, src/map_manager.cpp:61-82, 264-316 -> dlsc_build_edt)
    std::vector<float> boxes;         // [nb][6] cx, cy, cz, sx, sy, sz; used when from_boxes
    bool from_boxes = false;
    double maxdist = 1.0;             // DynamicEDTOctomap(maxdist = 1.0, ...)
};
namespace Eigen { struct MatrixXd {}; }
#else
#include <ros/ros.h>
#include <octomap/OcTree.h>
#include <dynamicEDT3D/dynamicEDTOctomap.h>
#include <Eigen/Dense>
#endif

namespace MATP {

typedef octomap::point3d point3d;
typedef std::vector<point3d> point3ds;

// ---- include/sp_const.hpp:56-160 ----
enum class PlannerMode { DLSCGC };
enum PlanningReport { Initialized, INITTRAJGENERATIONFAILED, CONSTRAINTGENERATIONFAILED, QPFAILED, WAITFORROSMSG, SUCCESS };
enum ObstacleType { DEFAULT, AGENT, DYN_SPIN, DYN_STRAIGHT, DYN_PATROL, DYN_CHASING, DYN_GAUSSIAN, DYN_REAL };
struct PlanningTime {
    void update(double time) {
        current = time;
        if (time < min) min = time;
        if (time > max) max = time;
        N_sample++;
        average = (average * (N_sample - 1) + time) / N_sample;
    }
    double current = 0, min = 1e9, max = 0, average = 0;
    int N_sample = 0;
};
struct PlanningTimeStatistics {
    PlanningTime mapf_time, initial_traj_planning_time, obstacle_prediction_time, goal_planning_time,
        lsc_generation_time, sfc_generation_time, traj_optimization_time, total_planning_time;
};
struct PlanningStatistics { int planning_seq = 0; PlanningTimeStatistics planning_time; };
struct State { point3d position, velocity, acceleration; };
struct Agent {
    int id = 0, cid = 0;
    State current_state;
    point3d start_point, desired_goal_point, current_goal_point, next_waypoint;
    double max_vel = 1, max_acc = 2, radius = 0.15, downwash = 2, nominal_velocity = 1;
};
typedef std::vector<Agent> Agents;

// ---- include/trajectory.hpp:9-70 (control-point container part) ----
template <typename T>
class Segment {
public:
    std::vector<T> control_points;
    double segment_time = 0;
    T startPoint() const { return control_points.front(); }
    T lastPoint() const { return control_points.back(); }
    T operator[](int idx) const { return control_points[idx]; }
    T& operator[](int idx) { return control_points[idx]; }
};
template <typename T>
class Trajectory {
public:
    Trajectory() : M(0), n(0) {}
    Trajectory(size_t M_, size_t n_, double dt) : M(M_), n(n_) {
        segments.resize(M);
        for (auto& s : segments) { s.control_points.resize(n + 1); s.segment_time = dt; }
    }
    int size() const { return (int)segments.size(); }
    bool empty() const { return segments.empty(); }
    void clear() { segments.clear(); M = 0; }
    T startPoint() const { return segments.front().startPoint(); }
    T lastPoint() const { return segments.back().lastPoint(); }
    Segment<T> operator[](int idx) const { return segments[idx]; }
    Segment<T>& operator[](int idx) { return segments[idx]; }
private:
    size_t M, n;
    std::vector<Segment<T>> segments;
};
typedef Trajectory<point3d> traj_t;

// ---- include/obstacle.hpp:13-44 ----
struct Obstacle {
    ros::Time update_time;
    ObstacleType type = ObstacleType::DEFAULT;
    int id = -1;
    double radius = 0, downwash = 0, max_acc = 0;
    point3d position, velocity, goal_point, observed_position;
    Trajectory<point3d> prev_traj;
};
typedef std::vector<Obstacle> Obstacles;
struct CollisionAlert {
    point3d agent_position;
    Obstacles obstacles;
    void initialize() { obstacles.clear(); }
    bool activated() const { return !obstacles.empty(); }
};

// ---- the fields of include/param.hpp / include/mission.hpp the hot path reads ----
class Param {
public:
    int world_dimension = 3;
    bool world_use_octomap = false;
    double world_resolution = 0.1, world_z_2d = 1.0;
    double multisim_time_step = 0.2;
    PlannerMode planner_mode = PlannerMode::DLSCGC;
    double dt = 0.2;
    int M = 5, n = 5, phi = 3, phi_n = 1;
    double control_input_weight = 0.01, terminal_weight = 1.0, slack_collision_weight = 1.0;
    bool obs_size_prediction = true;
    double obs_uncertainty_horizon = 1.0;
    double grid_resolution = 0.5;
    double goal_threshold = 0.1, reset_threshold = 0.5, slack_threshold = 0.1;
    double communication_range = -1.0;
    int max_neighbours = 0;        // 0 -> qn - 1 (capacity of the device neighbour list; not in the reference)
};
class Mission {
public:
    size_t qn = 0, on = 0;
    Agents agents;
    point3d world_min, world_max;
};

// ---- include/collision_constraints.hpp:15-111 ----
class LSC {
public:
    LSC() = default;
    LSC(const point3d& obs_control_point_, const point3d& normal_vector_, double d_)
        : obs_control_point(obs_control_point_), normal_vector(normal_vector_), d(d_) {}
    bool isPointInLSC(const point3d& point) const {
        const point3d r = point - obs_control_point;
        return (double)(float)(r.x() * normal_vector.x() + r.y() * normal_vector.y() + r.z() * normal_vector.z()) - d > 0;
    }
    point3d obs_control_point, normal_vector;
    double d = 0;
};
typedef std::vector<LSC> LSCs;
class Box {
public:
    point3d box_min, box_max;
    Box() = default;
    Box(const point3d& mn, const point3d& mx) : box_min(mn), box_max(mx) {}
    bool isPointInBox(const point3d& p) const {
        for (unsigned k = 0; k < 3; k++)
            if (!(p(k) > box_min(k) - 1e-5 && p(k) < box_max(k) + 1e-5)) return false;
        return true;
    }
};
typedef std::vector<Box> SFCs;

struct TrajOptResult {
    traj_t desired_traj;
    double total_qp_cost = 0;
    CollisionAlert collision_alert;
};

// ---------------------------------------------------------------------------------------------------
namespace detail {

inline void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + dlsc_last_error());
}
inline dlsc_params make_params(const Param& p, const Mission& m, int max_nbr) {
    if (p.planner_mode != PlannerMode::DLSCGC) throw std::invalid_argument("[TrajPlanner] Invalid planner mode");   // traj_planner.cpp:237
    if (p.n != 5 || p.phi != 3) throw std::invalid_argument("[TrajPlanner] only n = 5, phi = 3 are supported");
    dlsc_params q{};
    q.M = p.M; q.n = p.n; q.phi = p.phi; q.dim = p.world_dimension; q.use_sfc = p.world_use_octomap ? 1 : 0;
    q.max_nbr = max_nbr;
    q.dt = p.dt;
    for (unsigned k = 0; k < 3; k++) { q.world_min[k] = m.world_min(k); q.world_max[k] = m.world_max(k); }
    q.world_res = p.world_resolution; q.grid_res = p.grid_resolution; q.z_2d = p.world_z_2d;
    q.comm_range = p.communication_range; q.w_control = p.control_input_weight; q.w_terminal = p.terminal_weight;
    q.reset_threshold = p.reset_threshold;
    return q;
}
inline traj_t to_traj(const float* t, int M, int n, double dt) {
    traj_t out(M, n, dt);
    for (int m = 0; m < M; m++)
        for (int i = 0; i <= n; i++) out[m][i] = point3d(t[(m * (n + 1) + i) * 3], t[(m * (n + 1) + i) * 3 + 1], t[(m * (n + 1) + i) * 3 + 2]);
    return out;
}
inline void from_traj(const traj_t& tr, int M, int n, float* t) {
    for (int m = 0; m < M; m++)
        for (int i = 0; i <= n; i++)
            for (unsigned k = 0; k < 3; k++) t[(m * (n + 1) + i) * 3 + k] = tr[m][i](k);
}

// One device context shared by all TrajPlanner objects of a mission.
class SwarmBatch {
public:
    // Registry: the planners of ONE mission share a context.  The key is the content the context depends on -- every
    // planner parameter the device reads and the whole mission (world box, every agent's start, goal and properties) --
    // so two missions that merely have the same agent count never alias, and a changed mission (the reference's
    // mission_changed path builds new planners) gets a fresh context.
    static std::string key_of(const Param& p, const Mission& m) {
        std::string k;
        auto put = [&k](const void* q, size_t n) { k.append(static_cast<const char*>(q), n); };
        const double pd[] = {p.world_resolution, p.world_z_2d, p.dt, p.control_input_weight, p.terminal_weight, p.grid_resolution,
                             p.reset_threshold, p.communication_range, p.slack_collision_weight, p.obs_uncertainty_horizon};
        const int pi[] = {p.world_dimension, p.world_use_octomap ? 1 : 0, p.M, p.n, p.phi, p.max_neighbours, (int)m.qn,
                          p.obs_size_prediction ? 1 : 0};
        put(pd, sizeof(pd)); put(pi, sizeof(pi));
        for (unsigned c = 0; c < 3; c++) { const float w[2] = {m.world_min(c), m.world_max(c)}; put(w, sizeof(w)); }
        for (const Agent& a : m.agents) {
            const float f[6] = {a.start_point(0), a.start_point(1), a.start_point(2), a.desired_goal_point(0), a.desired_goal_point(1),
                                a.desired_goal_point(2)};
            const double d[5] = {a.max_vel, a.max_acc, a.radius, a.downwash, a.nominal_velocity};
            put(f, sizeof(f)); put(d, sizeof(d));
        }
        return k;
    }
    static std::shared_ptr<SwarmBatch> get(const Param& p, const Mission& m) {
        static std::map<std::string, std::weak_ptr<SwarmBatch>> reg;
        const std::string key = key_of(p, m);
        for (auto it = reg.begin(); it != reg.end();) it = it->second.expired() ? reg.erase(it) : std::next(it);
        auto it = reg.find(key);
        if (it != reg.end())
            if (auto sp = it->second.lock()) return sp;
        auto sp = std::shared_ptr<SwarmBatch>(new SwarmBatch(p, m));
        reg[key] = sp;
        return sp;
    }
    ~SwarmBatch() { if (ctx) dlsc_destroy(ctx); }

    void new_step() {                        // called from every setObstacles(); idempotent within a step
        if (!step_open) return;
        step_open = false; batch_done = false;
        std::fill(staged.begin(), staged.end(), 0);
        std::fill(planned.begin(), planned.end(), 0);
        std::fill(have_list.begin(), have_list.end(), 0);
    }
    // The obstacle list the caller handed to agent i defines that agent's neighbours (the reference plans against exactly
    // the list broadcastMsgs built, src/multi_sync_simulator.cpp:481-503): dynamic obstacles first (entries N + o), then
    // the agents in ascending id.  With the reference's own filter this equals what the device search finds.
    void set_list(int i, const Obstacles& all) {
        if (i < 0 || i >= N) return;
        int32_t* row = nbr_idx.data() + (size_t)i * Kcap;
        int c = 0, nd = 0;
        for (const Obstacle& o : all) if (o.type != ObstacleType::AGENT) nd++;
        for (int o = 0; o < nd && c < Kcap; o++) row[c++] = N + o;
        std::vector<int> ids;
        for (const Obstacle& o : all) if (o.type == ObstacleType::AGENT && o.id >= 0 && o.id < N && o.id != i) ids.push_back(o.id);
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        if (nd + (int)ids.size() > Kcap) throw std::length_error("[TrajPlanner] more obstacles than the device list holds (Param::max_neighbours)");
        for (int id : ids) row[c++] = id;
        nbr_cnt[i] = c;
        have_list[i] = 1;
    }
    void stage(const Agent& a, bool disturbed) {
        const int i = a.id;
        for (unsigned k = 0; k < 3; k++) {
            pos[i * 3 + k] = a.current_state.position(k); vel[i * 3 + k] = a.current_state.velocity(k);
            acc[i * 3 + k] = a.current_state.acceleration(k); wp[i * 3 + k] = a.next_waypoint(k);
        }
        dist[i] = disturbed ? 1 : 0;
        staged[i] = 1;
    }
    // what broadcastMsgs tells an agent about the others (position / velocity of this step)
    void observe(const Obstacle& o) {
        if (o.type != ObstacleType::AGENT || o.id < 0 || o.id >= N) return;
        for (unsigned k = 0; k < 3; k++) { pos[o.id * 3 + k] = o.position(k); vel[o.id * 3 + k] = o.velocity(k); }
    }
    // the non-agent entries of the obstacle list broadcastMsgs hands to every agent (the same for all of them): position,
    // velocity, radius, downwash, max_acc -> dlsc_set_obstacles, re-sent only when something changed
    void set_dynamic(const Obstacles& all) {
        std::vector<float> p, v;
        std::vector<double> r, dw, ma;
        for (const Obstacle& o : all) {
            if (o.type == ObstacleType::AGENT) continue;
            for (unsigned k = 0; k < 3; k++) { p.push_back(o.position(k)); v.push_back(o.velocity(k)); }
            r.push_back(o.radius); dw.push_back(o.downwash); ma.push_back(o.max_acc);
        }
        if (r.size() > (size_t)DLSC_MAX_OBSTACLES)
            throw std::length_error("[TrajPlanner] more than " + std::to_string(DLSC_MAX_OBSTACLES) + " dynamic obstacles");
        if (p == dyn_pos && v == dyn_vel && r == dyn_r && dw == dyn_dw && ma == dyn_ma) return;
        dyn_pos = p; dyn_vel = v; dyn_r = r; dyn_dw = dw; dyn_ma = ma;
        if (r.empty()) { check(dlsc_set_obstacles(ctx, nullptr, nullptr), "dlsc_set_obstacles"); return; }
        dlsc_obstacles o{(int32_t)r.size(), dyn_pos.data(), dyn_vel.data(), dyn_r.data(), dyn_dw.data(), dyn_ma.data()};
        dlsc_obstacle_params op{slack_w, obs_horizon, obs_size_pred ? 1 : 0, 0};
        check(dlsc_set_obstacles(ctx, &o, &op), "dlsc_set_obstacles");
    }
    int n_dynamic() const { return (int)dyn_r.size(); }
    void set_distmap(const std::shared_ptr<DynamicEDTOctomap>& d) {
#ifdef DLSC_COMPAT_STANDALONE
        if (!d || d.get() == distmap_seen) return;
        if (d->from_boxes)
            check(dlsc_build_edt(ctx, d->boxes.data(), (int)(d->boxes.size() / 6), d->maxdist), "dlsc_build_edt");
        else
            check(dlsc_set_edt(ctx, d->dist.data(), d->obst.data(), d->dims, d->min_key, d->res), "dlsc_set_edt");
        distmap_seen = d.get();
#else
        // real dynamicEDT3D: export the live grid once per map -- one accessor call per cell centre, the very call the
        // reference's SFC code makes per lattice vertex (collision_constraints.cpp:880) -- into the device context
        if (!d || d.get() == distmap_seen) return;
        int32_t dims[3], mk[3];
        check(dlsc_edt_dims(ctx, dims, mk), "dlsc_edt_dims");
        const size_t nc = (size_t)dims[0] * dims[1] * dims[2];
        std::vector<float> dist_v(nc);
        std::vector<int32_t> obst_v(nc * 3);
        const double inv = 1.0 / res;
        size_t i = 0;
        for (int x = 0; x < dims[0]; x++)
            for (int y = 0; y < dims[1]; y++)
                for (int z = 0; z < dims[2]; z++, i++) {
                    const point3d centre((float)(((double)(x + mk[0]) + 0.5) * res), (float)(((double)(y + mk[1]) + 0.5) * res),
                                         (float)(((double)(z + mk[2]) + 0.5) * res));
                    float dd = -1.f;
                    point3d cl;
                    d->getDistanceAndClosestObstacle(centre, dd, cl);
                    dist_v[i] = dd;
                    const bool has = dd >= 0.f && dd < 1.f;            // beyond that the vertex test never looks at the obstacle
                    for (unsigned k = 0; k < 3; k++) obst_v[3 * i + k] = has ? (int32_t)std::floor(inv * (double)cl(k)) - mk[k] : -1;
                }
        check(dlsc_set_edt(ctx, dist_v.data(), obst_v.data(), dims, mk, res), "dlsc_set_edt");
        distmap_seen = d.get();
#endif
    }
    // replan agent i (or everybody when all agents were staged); fills the result cache
    void plan(int i) {
        step_open = true;
        // one batched launch only when every agent was announced before the first plan() of the step
        const bool none_planned = std::none_of(planned.begin(), planned.end(), [](uint8_t s) { return s != 0; });
        const bool all = (batch_done || none_planned) && std::all_of(staged.begin(), staged.end(), [](uint8_t s) { return s != 0; });
        batch_last = all;
        if (all && !batch_done) {
            upload();
            if (std::all_of(have_list.begin(), have_list.end(), [](uint8_t s) { return s != 0; })) {
                // the callers' obstacle lists are the neighbour lists: every stage but the device neighbour search
                check(dlsc_set_neighbours(ctx, nbr_idx.data(), nbr_cnt.data()), "dlsc_set_neighbours");
                check(dlsc_run_stages(ctx, DLSC_STAGE_ALL & ~DLSC_STAGE_NBR), "dlsc_run_stages");
                check(dlsc_set_seq(ctx, dlsc_get_seq(ctx) + 1), "dlsc_set_seq");
            } else {
                check(dlsc_step(ctx), "dlsc_step");
            }
            check(dlsc_publish_records(ctx), "dlsc_publish_records");
            fetch();
            batch_done = true;
        } else if (!all) {
            upload();
            if (have_list[i]) {
                check(dlsc_set_neighbours(ctx, nbr_idx.data(), nbr_cnt.data()), "dlsc_set_neighbours");
                check(dlsc_run_stages_subset(ctx, DLSC_STAGE_ALL & ~DLSC_STAGE_NBR, i, 1), "dlsc_run_stages_subset");
            } else {
                check(dlsc_run_stages_subset(ctx, DLSC_STAGE_ALL, i, 1), "dlsc_run_stages_subset");
            }
            pending_publish = true;
            fetch();
        }
        planned[i] = 1;
        if (!batch_done && std::all_of(planned.begin(), planned.end(), [](uint8_t s) { return s != 0; })) finish_serial_step();
    }
    void finish_serial_step() {
        // every agent was replanned one at a time from the same previous-step records: publish them together
        check(dlsc_publish_records(ctx), "dlsc_publish_records");
        check(dlsc_set_seq(ctx, dlsc_get_seq(ctx) + 1), "dlsc_set_seq");
        pending_publish = false;
        batch_done = true;
    }
    int N = 0, M = 0, n = 5;
    double dt = 0.2, res = 0.1;
    dlsc_ctx* ctx = nullptr;
    // the device context, for callers that bind more than the three classes do (e.g. MapManager handing over an
    // occupancy grid through dlsc_build_edt_occupancy)
    dlsc_ctx* context() const { return ctx; }
    // per-stage device time of the last step, amortised per agent [s]: predict, neighbours, LSC, SFC, goal, QP
    double stage_s[DLSC_N_STAGES] = {0, 0, 0, 0, 0, 0};
    std::vector<float> traj, goal;
    std::vector<double> cost;
    std::vector<double> slack;                     // [N][n_dynamic][M] slack variables of the last step
    std::vector<int32_t> status;
    std::vector<uint8_t> staged, planned, have_list;
    std::vector<int32_t> nbr_idx, nbr_cnt;         // per-agent obstacle lists as the callers passed them (set_list)
    int Kcap = 0;

private:
    SwarmBatch(const Param& p, const Mission& m) {
        N = (int)m.qn; M = p.M; n = p.n; dt = p.dt; res = p.world_resolution;
        slack_w = p.slack_collision_weight; obs_horizon = p.obs_uncertainty_horizon; obs_size_pred = p.obs_size_prediction;
        if (N < 1 || (size_t)N != m.agents.size()) throw std::invalid_argument("[TrajPlanner] mission has no agents");
        // capacity of the per-agent obstacle list: the agents in range plus room for the mission's dynamic obstacles
        const int K = (p.max_neighbours > 0 ? p.max_neighbours : std::max(N - 1, 1)) + DLSC_MAX_OBSTACLES;
        dlsc_params q = make_params(p, m, K);
        check(dlsc_create(&q, N, 0, N, 0, &ctx), "dlsc_create");
        std::vector<double> r(N), dw(N), mv(N), ma(N), nv(N);
        std::vector<float> start(3 * N);
        for (int i = 0; i < N; i++) {
            const Agent& a = m.agents[i];
            r[i] = a.radius; dw[i] = a.downwash; mv[i] = a.max_vel; ma[i] = a.max_acc; nv[i] = a.nominal_velocity;
            for (unsigned k = 0; k < 3; k++) start[3 * i + k] = a.start_point(k);
        }
        dlsc_agent_props pr{r.data(), dw.data(), mv.data(), ma.data(), nv.data()};
        check(dlsc_set_agent_props(ctx, &pr), "dlsc_set_agent_props");
        check(dlsc_reset(ctx, start.data()), "dlsc_reset");
        check(dlsc_enable_timing(ctx, 1), "dlsc_enable_timing");      // feeds PlanningStatistics (reference-scale missions)
        pos.assign(3 * N, 0.f); vel = pos; acc = pos; wp = pos; goal = pos;
        dist.assign(N, 0); staged.assign(N, 0); planned.assign(N, 0);
        Kcap = K; have_list.assign(N, 0); nbr_idx.assign((size_t)N * K, 0); nbr_cnt.assign(N, 0);
        traj.assign((size_t)N * M * (n + 1) * 3, 0.f); cost.assign(N, 0.0); status.assign(N, 0);
    }
    void upload() {
        dlsc_agents a{pos.data(), vel.data(), acc.data(), wp.data(), dist.data()};
        check(dlsc_set_agents(ctx, &a), "dlsc_set_agents");
        check(dlsc_sync(ctx), "dlsc_sync");
    }
    void fetch() {
        check(dlsc_get_traj(ctx, traj.data()), "dlsc_get_traj");
        check(dlsc_get_cost(ctx, cost.data()), "dlsc_get_cost");
        check(dlsc_get_status(ctx, status.data()), "dlsc_get_status");
        check(dlsc_get_goal(ctx, goal.data()), "dlsc_get_goal");
        if (n_dynamic() > 0) {
            slack.assign((size_t)N * n_dynamic() * M, 0.0);
            check(dlsc_get_slack(ctx, slack.data()), "dlsc_get_slack");
        }
        double ms[DLSC_N_STAGES];
        int n_steps = 0;
        check(dlsc_get_timings(ctx, ms, &n_steps), "dlsc_get_timings");
        if (n_steps > 0)
            for (int s = 0; s < DLSC_N_STAGES; s++) stage_s[s] = ms[s] * 1e-3 / (batch_last ? N : 1);
    }
    std::vector<float> pos, vel, acc, wp;
    std::vector<float> dyn_pos, dyn_vel;           // dynamic obstacles last sent to the device
    std::vector<double> dyn_r, dyn_dw, dyn_ma;
    double slack_w = 1.0, obs_horizon = 1.0;
    bool obs_size_pred = true;
    std::vector<uint8_t> dist;
    bool step_open = true, batch_done = false, pending_publish = false, batch_last = false;
    const void* distmap_seen = nullptr;
};
}  // namespace detail

// ---------------------------------------------------------------------------------------------------
// CollisionConstraints: value store of the LSCs / SFCs (what TrajOptimizer::solve reads)
// ---------------------------------------------------------------------------------------------------
class CollisionConstraints {
public:
    CollisionConstraints(const Param& param_, const Mission& mission_, double radius, double max_vel)
        : param(param_), mission(mission_), agent_radius(radius), agent_max_vel(max_vel) {
        sfcs.resize(param.M);
    }
    void initializeLSC(const Obstacles& obstacles) {                       // collision_constraints.cpp:454-461
        N_obs = (int)obstacles.size();
        types.resize(N_obs); obs_positions.resize(N_obs);
        for (int oi = 0; oi < N_obs; oi++) { types[oi] = obstacles[oi].type; obs_positions[oi] = obstacles[oi].position; }
        lscs.assign(N_obs, std::vector<LSCs>(param.M, LSCs(param.n + 1)));
    }
    void initializeSFC(const point3d&) { throw std::logic_error("CollisionConstraints::initializeSFC runs inside TrajPlanner::plan on the device"); }
    void constructSFCFromPoint(const point3d&, const point3d&) { throw std::logic_error("not on the DLSCGC path"); }
    void constructSFCFromConvexHull(const point3ds&, const point3d&) { throw std::logic_error("not on the DLSCGC path"); }
    void constructSFCFromInitialTraj(const traj_t&, const point3d&, const point3d&) {
        throw std::logic_error("CollisionConstraints::constructSFCFromInitialTraj runs inside TrajPlanner::plan on the device");
    }
    void constructCommunicationRange(const point3d& next_waypoint) {       // :538-546
        if (param.communication_range > 0) {
            const float h = (float)(0.5 * param.communication_range);
            communication_range = Box(next_waypoint - point3d(h, h, h), next_waypoint + point3d(h, h, h));
        }
    }
    bool isDynamicObstacle(int oi) const { return types[oi] != ObstacleType::AGENT; }
    bool isPointInFeasibleRegion(const point3d& point, int m, int i) const {   // :586-598
        for (int oi = 0; oi < N_obs; oi++)
            if (!lscs[oi][m][i].isPointInLSC(point)) return false;
        if (param.world_use_octomap && !sfcs[m].isPointInBox(point)) return false;
        return communication_range.isPointInBox(point);
    }
    LSC getLSC(int oi, int m, int i) const { return lscs[oi][m][i]; }
    Box getSFC(int m) const { return sfcs[m]; }
    size_t getObsSize() const { return (size_t)N_obs; }
    point3d getObsPosition(int oi) const { return obs_positions[oi]; }
    void setDistmap(std::shared_ptr<DynamicEDTOctomap> d) { distmap_ptr = std::move(d); }
    void setOctomap(std::shared_ptr<octomap::OcTree> o) { octree_ptr = std::move(o); }
    void setLSC(int oi, int m, int i, const LSC& lsc) { lscs[oi][m][i] = lsc; }
    void setSFC(int m, const Box& sfc) { sfcs[m] = sfc; }

private:
    friend class TrajOptimizer;
    Param param;
    Mission mission;
    double agent_radius, agent_max_vel;
    std::shared_ptr<DynamicEDTOctomap> distmap_ptr;
    std::shared_ptr<octomap::OcTree> octree_ptr;
    int N_obs = 0;
    std::vector<ObstacleType> types;
    point3ds obs_positions;
    std::vector<std::vector<LSCs>> lscs;    // [obs][segment][control point]
    SFCs sfcs;
    Box communication_range;
};

// ---------------------------------------------------------------------------------------------------
// TrajOptimizer: one QP on the device from explicitly given constraints (batch of one)
// ---------------------------------------------------------------------------------------------------
class TrajOptimizer {
public:
    TrajOptimizer(const Param& param_, const Mission& mission_, const Eigen::MatrixXd&) : param(param_), mission(mission_) {}
    void updateParam(const Param& param_) { param = param_; }

    TrajOptResult solve(const Agent& agent, const CollisionConstraints& constraints, const traj_t& initial_traj,
                        bool /*use_primal_algorithm*/) {
        const int M = param.M, n = param.n, P = n + 1;
        const int K = std::max((int)constraints.getObsSize(), 1), N = K + 1;
        Mission mm = mission;
        mm.qn = N;
        dlsc_params q = detail::make_params(param, mm, K);
        dlsc_ctx* ctx = nullptr;
        detail::check(dlsc_create(&q, N, 0, 1, 0, &ctx), "dlsc_create");
        struct Guard { dlsc_ctx* c; ~Guard() { dlsc_destroy(c); } } guard{ctx};
        double r = agent.radius, dw = agent.downwash, mv = agent.max_vel, ma = agent.max_acc, nv = agent.nominal_velocity;
        dlsc_agent_props pr{&r, &dw, &mv, &ma, &nv};
        detail::check(dlsc_set_agent_props(ctx, &pr), "dlsc_set_agent_props");
        float start[3] = {agent.current_state.position(0), agent.current_state.position(1), agent.current_state.position(2)};
        detail::check(dlsc_reset(ctx, start), "dlsc_reset");
        // record of the agent: state + current goal
        const int rec_n = dlsc_record_floats(ctx);
        std::vector<float> rec(rec_n, 0.f);
        detail::check(dlsc_get_records(ctx, 0, 1, rec.data()), "dlsc_get_records");
        const int o = M * P * 3;
        for (unsigned k = 0; k < 3; k++) {
            rec[o + k] = agent.current_state.position(k); rec[o + 3 + k] = agent.current_state.velocity(k);
            rec[o + 6 + k] = agent.current_goal_point(k);
        }
        detail::check(dlsc_set_records(ctx, 0, 1, rec.data()), "dlsc_set_records");
        float acc[3] = {agent.current_state.acceleration(0), agent.current_state.acceleration(1), agent.current_state.acceleration(2)};
        float wp[3] = {agent.next_waypoint(0), agent.next_waypoint(1), agent.next_waypoint(2)};
        dlsc_agents ag{nullptr, nullptr, acc, wp, nullptr};
        detail::check(dlsc_set_agents(ctx, &ag), "dlsc_set_agents");
        // constraints: obstacle oi -> fake neighbour oi + 1 whose "predicted trajectory" carries the anchors
        std::vector<float> init(M * P * 3), pred((size_t)N * M * P * 3, 0.f), normal((size_t)K * M * 3, 0.f), alast((size_t)K * 3, 0.f);
        std::vector<double> d((size_t)K * M * P, 0.0);
        std::vector<int32_t> idx(K, 0), cnt(1, (int32_t)constraints.getObsSize());
        detail::from_traj(initial_traj, M, n, init.data());
        for (int oi = 0; oi < (int)constraints.getObsSize(); oi++) {
            idx[oi] = oi + 1;
            for (int m = 0; m < M; m++) {
                const LSC l0 = constraints.getLSC(oi, m, 0);
                for (unsigned k = 0; k < 3; k++) normal[((size_t)oi * M + m) * 3 + k] = l0.normal_vector(k);
                for (int i = 0; i < P; i++) {
                    const LSC l = constraints.getLSC(oi, m, i);
                    d[((size_t)oi * M + m) * P + i] = l.d;
                    for (unsigned k = 0; k < 3; k++) pred[(((size_t)(oi + 1) * M + m) * P + i) * 3 + k] = l.obs_control_point(k);
                    if (m == M - 1 && i == n)
                        for (unsigned k = 0; k < 3; k++) alast[(size_t)oi * 3 + k] = l.obs_control_point(k);
                }
            }
        }
        std::vector<float> sfc(M * 6, 0.f);
        for (int m = 0; m < M; m++)
            for (unsigned k = 0; k < 3; k++) { sfc[m * 6 + k] = constraints.getSFC(m).box_min(k); sfc[m * 6 + 3 + k] = constraints.getSFC(m).box_max(k); }
        detail::check(dlsc_set_init_traj(ctx, init.data()), "dlsc_set_init_traj");
        detail::check(dlsc_set_pred_traj(ctx, pred.data()), "dlsc_set_pred_traj");
        detail::check(dlsc_set_neighbours(ctx, idx.data(), cnt.data()), "dlsc_set_neighbours");
        detail::check(dlsc_set_lsc(ctx, normal.data(), alast.data(), d.data()), "dlsc_set_lsc");
        detail::check(dlsc_set_sfc(ctx, sfc.data(), nullptr), "dlsc_set_sfc");
        detail::check(dlsc_run_stages(ctx, DLSC_STAGE_QP), "dlsc_run_stages");
        std::vector<float> out(M * P * 3);
        double cost = 0; int32_t status = 0;
        detail::check(dlsc_get_traj(ctx, out.data()), "dlsc_get_traj");
        detail::check(dlsc_get_cost(ctx, &cost), "dlsc_get_cost");
        detail::check(dlsc_get_status(ctx, &status), "dlsc_get_status");
        if (status & (DLSC_QP_MAXITER | DLSC_QP_NUMERIC)) throw PlanningReport::QPFAILED;     // traj_optimizer.cpp:152,161
        TrajOptResult res;
        res.desired_traj = detail::to_traj(out.data(), M, n, param.dt);
        res.total_qp_cost = cost;
        return res;
    }

private:
    Param param;
    Mission mission;
};

// ---------------------------------------------------------------------------------------------------
// TrajPlanner
// ---------------------------------------------------------------------------------------------------
class TrajPlanner {
public:
    TrajPlanner(const ros::NodeHandle& nh_, const Param& param_, const Mission& mission_, const Agent& agent_)
        : param(param_), mission(mission_), nh(nh_), agent(agent_), planner_seq(0) {
        batch = detail::SwarmBatch::get(param, mission);
    }

    // Optional: announce every agent of the step before the plan() loop -> one batched launch per step.
    static void stageAgent(const Param& param, const Mission& mission, const Agent& agent, bool is_disturbed) {
        detail::SwarmBatch::get(param, mission)->stage(agent, is_disturbed);
    }

    TrajOptResult plan(const Agent& agent_, const std::shared_ptr<octomap::OcTree>& /*octree_ptr*/,
                       const std::shared_ptr<DynamicEDTOctomap>& distmap_ptr, ros::Time /*sim_current_time*/,
                       bool is_disturbed) {
        agent = agent_;
        planner_seq++;
        statistics.planning_seq = planner_seq;
        if (param.world_use_octomap) batch->set_distmap(distmap_ptr);
        if (!batch->staged[agent.id]) batch->stage(agent, is_disturbed);
        batch->plan(agent.id);
        const int st = batch->status[agent.id];
        {   // PlanningStatistics as TrajPlanner::plan fills it (traj_planner.cpp:42-62), from the device stage timers;
            // a batched step reports every stage amortised per agent
            PlanningTimeStatistics& pt = statistics.planning_time;
            const double* s = batch->stage_s;
            pt.obstacle_prediction_time.update(s[0]); pt.initial_traj_planning_time.update(0.0);
            pt.lsc_generation_time.update(s[1] + s[2]); pt.sfc_generation_time.update(s[3]);
            pt.goal_planning_time.update(s[4]); pt.traj_optimization_time.update(s[5]);
            pt.total_planning_time.update(s[0] + s[1] + s[2] + s[3] + s[4] + s[5]);
        }
        if (st & DLSC_NBR_OVERFLOW)     // the reference has no neighbour cap: planning without some neighbours' constraints is an error
            throw std::length_error("[TrajPlanner] more agents in communication range than Param::max_neighbours");
        if (st & DLSC_SFC_INIT_FAILED) throw std::invalid_argument("[CollisionConstraints] Invalid initial SFC");  // collision_constraints.cpp:445-447
        if (st & DLSC_GOAL_INFEASIBLE) throw PlanningReport::QPFAILED;                                             // goal_optimizer.cpp:122,132
        TrajOptResult res;   // QP failure: the library already substituted initial_traj (traj_planner.cpp:749-777)
        const size_t L = (size_t)batch->M * (batch->n + 1) * 3;
        res.desired_traj = detail::to_traj(batch->traj.data() + L * agent.id, batch->M, batch->n, batch->dt);
        res.total_qp_cost = batch->cost[agent.id];
        // collision alert (traj_optimizer.cpp:84-105): dynamic obstacles whose slack variables sum above plan/slack_threshold
        if (!(st & (DLSC_QP_MAXITER | DLSC_QP_NUMERIC)) && batch->n_dynamic() > 0) {
            const int nd = batch->n_dynamic();
            int o = 0;
            for (size_t oi = 0; oi < obstacles.size() && o < nd; oi++) {
                if (obstacles[oi].type == ObstacleType::AGENT) continue;
                double slack_cost = 0;
                for (int m = 0; m < batch->M; m++) slack_cost += std::abs(batch->slack[((size_t)agent.id * nd + o) * batch->M + m]);
                if (slack_cost > param.slack_threshold) {
                    res.collision_alert.agent_position = agent.current_state.position;
                    Obstacle alert;
                    alert.id = (int)oi;
                    alert.position = obstacles[oi].position;
                    res.collision_alert.obstacles.emplace_back(alert);
                }
                o++;
            }
        }
        for (unsigned k = 0; k < 3; k++) agent.current_goal_point(k) = batch->goal[3 * agent.id + k];
        return res;
    }
    void publish() {}
    void setObstacles(const Obstacles& obstacles_) {
        obstacles = obstacles_;
        batch->new_step();
        batch->set_dynamic(obstacles);             // non-agent entries: dynamic obstacles (size prediction, slack QP) on the device
        batch->set_list(agent.id, obstacles);      // agent entries: this planner's neighbours
        for (const auto& o : obstacles) batch->observe(o);
    }
    // the mission's shared device context (NULL never): MapManager-side bindings hand maps over through it
    dlsc_ctx* deviceContext() const { return batch->context(); }
    int getPlannerSeq() const { return planner_seq; }
    point3d getCurrentGoalPosition() const { return agent.current_goal_point; }
    PlanningStatistics getPlanningStatistics() const { return statistics; }

private:
    Param param;
    Mission mission;
    ros::NodeHandle nh;
    Agent agent;
    int planner_seq;
    PlanningStatistics statistics;
    Obstacles obstacles;
    std::shared_ptr<detail::SwarmBatch> batch;
};

}  // namespace MATP
