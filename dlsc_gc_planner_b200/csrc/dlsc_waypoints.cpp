// dlsc_waypoints.cpp -- the waypoint provider behind the C ABI (dlsc_wp_*): where every agent's next_waypoint comes from.
//
// Replaces, for agent-only missions (no dynamic obstacles), the reference's
//   MultiSyncSimulator::decentralizedMAPP      src/multi_sync_simulator.cpp:308-466   comm-range groups, update rules
//   GridBasedPlanner::planMAPF / runMAPF / updateGridMap / updatePlanResult / planInitialPath
//                                               src/grid_based_planner.cpp:64-164, 292-453
//   MAPF::PIBT (Okumura et al. 2019)            src/mapf/pibt.cpp:13-219, src/mapf/solver.cpp:270-290 (distance tables)
//   Grid                                        third_party/grid-pathfinding/graph/src/graph.cpp:371-431 (neighbour order)
// Host code: the search is a serial priority-inheritance recursion over a few dozen lattice nodes per agent and runs
// once per replan step for the whole swarm -- control plane next to the batched kernels, not a kernel itself.  Written
// from scratch on flat arrays (node ids, no pointer graph); PIBT's candidate shuffle uses std::shuffle on std::mt19937
// seeded 0 per call exactly like the reference (src/mapf/problem.cpp:85), so plans are identical to the reference's
// object code built with the same standard library (tests/test_waypoints.py against oracle/_ref/libmapf_ref.so).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <random>
#include <set>
#include <string>
#include <vector>

#include "../../include/dlsc_b200.h"

namespace {

constexpr double kEps = 1e-9, kEpsF = 1e-5;      // SP_EPSILON, SP_EPSILON_FLOAT (include/sp_const.hpp:3-4)
constexpr int kMaxTimestep = 5000;               // DEFAULT_MAX_TIMESTEP (include/mapf/default_params.hpp)

struct P3 { float x, y, z; };                    // octomap::point3d: float storage
inline P3 p3(float a, float b, float c) { P3 r; r.x = a; r.y = b; r.z = c; return r; }
inline P3 p3_load(const float* p) { return p3(p[0], p[1], p[2]); }
inline float comp(const P3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
inline P3 sub(const P3& a, const P3& b) { return p3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline double norm(const P3& a) { const float r = a.x * a.x + a.y * a.y + a.z * a.z; return std::sqrt((double)r); }   // octomath::Vector3::norm
inline double distance(const P3& a, const P3& b) {     // octomath::Vector3::distance: float differences, double accumulation
    const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}
inline double linf(const P3& a, const P3& b) {         // include/util.hpp:131-140
    double d = 0;
    double c = (double)std::fabs(a.x - b.x); if (d < c) d = c;
    c = (double)std::fabs(a.y - b.y); if (d < c) d = c;
    c = (double)std::fabs(a.z - b.z); if (d < c) d = c;
    return d;
}

struct PlanResult {                              // include/grid_based_planner.hpp:16-21
    std::set<size_t> agent_ids;
    std::vector<std::vector<P3>> paths;
    int makespan() const {                       // grid_based_planner.cpp:11-26 (as written there: the last longer path wins)
        if (paths.empty()) return -1;
        int max_size = 0;
        for (const auto& p : paths) {
            if (p.empty()) return -1;
            if (max_size < (int)p.size()) max_size = (int)p.size() - 1;
        }
        return max_size;
    }
};

}  // namespace

struct dlsc_wp {
    int N = 0, dim = 3, M = 5;
    double grid_res = 0.5, z_2d = 1.0, comm_range = -1.0, downwash0 = 2.0, radius0 = 0.15;
    std::array<double, 3> gmin{}, gmax{};
    std::array<int, 3> gdim{1, 1, 1};
    std::vector<P3> start, desired_goal;
    std::vector<uint8_t> exists;                 // [w * d * h], id = w * d * z + w * y + x
    std::vector<uint8_t> warning;                // node inside a dynamic obstacle's reachable region (updateGridMap :140-150); empty = none
    std::vector<std::array<int, 6>> nbr;         // neighbour ids in the reference's order (left, right, up, down, top, bottom), -1 = none
    std::vector<int8_t> nbr_n;
    PlanResult plan_result;                      // one member shared by all groups, like GridBasedPlanner::plan_result
    // BFS distance tables by goal node: the reference rebuilds them in every call (solver.cpp:270-290), but they depend on
    // the lattice and the goal node only, so they are kept until the lattice changes
    std::vector<std::vector<uint16_t>> dist_cache;   // [agent][node]
    std::vector<int> dist_goal;                      // goal node each cached table was built for (-1: none)
    // dynamic obstacles of the next dlsc_wp_step (dlsc_wp_set_obstacles) and the agents' collision alerts (dlsc_wp_set_alerts)
    double dt = 0.2, obs_horizon = 1.0;
    std::vector<P3> obs_pos, obs_vel;
    std::vector<double> obs_radius, obs_max_acc;
    std::vector<std::vector<int>> alerts;        // [agent] obstacle ids (CollisionAlert::obstacles after updateCollisionAlert)
    std::string err;
    int64_t pibt_timesteps = 0;                  // work counter: PIBT timesteps of the last dlsc_wp_step

    double gres(int i) const { return i < 2 ? grid_res : grid_res * downwash0; }     // getGridResolution :597-603
    P3 node_point(int id) const {                                                   // posToPoint3D :537-548
        const int w = gdim[0], d = gdim[1];
        const int z = id / (w * d), y = (id - z * w * d) / w, x = id - z * w * d - y * w;
        P3 q;
        q.x = (float)(gmin[0] + x * gres(0));
        q.y = (float)(gmin[1] + y * gres(1));
        q.z = (dim == 2) ? (float)z_2d : (float)(gmin[2] + z * gres(2));
        return q;
    }
    bool valid_point(const P3& q) const {                                           // isValid :366-374
        for (int i = 0; i < dim; i++)
            if (comp(q, i) < gmin[i] || comp(q, i) > gmax[i]) return false;
        return true;
    }
    int point_id(const P3& q) const {                                               // point3DToPos / point3DToID :554-577
        int g[3] = {0, 0, 0};
        for (int i = 0; i < dim; i++) {
            g[i] = (int)std::round((comp(q, i) - gmin[i]) / gres(i));
            if (g[i] < 0) g[i] = 0; else if (g[i] >= gdim[i]) g[i] = gdim[i] - 1;
        }
        return gdim[0] * gdim[1] * g[2] + gdim[0] * g[1] + g[0];
    }
    bool node_exists(int x, int y, int z) const {
        return x >= 0 && x < gdim[0] && y >= 0 && y < gdim[1] && z >= 0 && z < gdim[2] && exists[(size_t)gdim[0] * gdim[1] * z + gdim[0] * y + x];
    }
    bool occupied(const P3& q) const { return !valid_point(q) || !exists[point_id(q)]; }   // isOccupied :376-378
    int closest_node(const P3& q) const {                                           // point3DToClosestNode :482-535
        const int id = point_id(q);
        if (exists[id]) return id;
        const int w = gdim[0], d = gdim[1];
        const int z = id / (w * d), y = (id - z * w * d) / w, x = id - z * w * d - y * w;
        const int cand[6][3] = {{x + 1, y, z}, {x - 1, y, z}, {x, y + 1, z}, {x, y - 1, z}, {x, y, z + 1}, {x, y, z - 1}};
        int best = -1;
        double best_d = 1e300;                     // SP_INFINITY stands for "larger than any distance"
        for (const auto& c : cand)
            if (node_exists(c[0], c[1], c[2])) {
                const int cid = w * d * c[2] + w * c[1] + c[0];
                const double dd = distance(q, node_point(cid));
                if (dd < best_d) { best_d = dd; best = cid; }
            }
        if (best >= 0) return best;
        for (int cid = 0; cid < (int)exists.size(); cid++)
            if (exists[cid]) {
                const double dd = distance(q, node_point(cid));
                if (dd < best_d) { best_d = dd; best = cid; }
            }
        return best;
    }
    bool obstacle_reaches(int o, const P3& q) const {                               // Obstacle::isCollided, obstacle.hpp:26-36
        const double horizon = M * dt;
        const double step = std::min(0.1 * horizon, 0.1);
        for (double t = 0; t <= horizon; t += step) {
            P3 c;
            c.x = obs_pos[o].x + obs_vel[o].x * (float)t; c.y = obs_pos[o].y + obs_vel[o].y * (float)t; c.z = obs_pos[o].z + obs_vel[o].z * (float)t;
            const double tm = std::min(t, obs_horizon);
            if (distance(c, q) < radius0 + obs_radius[o] + 0.5 * obs_max_acc[o] * tm * tm) return true;
        }
        return false;
    }
    std::vector<int> bfs_table(int from) const {                                    // createDistanceTable :646-670 (directed edges)
        std::vector<int> tab(exists.size(), 1000000000);                           // (int)SP_INFINITY
        std::queue<int> open;
        open.push(from); tab[from] = 0;
        while (!open.empty()) {
            const int v = open.front(); open.pop();
            for (int k = 0; k < nbr_n[v]; k++) {
                const int m = nbr[v][k];
                if (tab[v] + 1 >= tab[m]) continue;
                tab[m] = tab[v] + 1;
                open.push(m);
            }
        }
        return tab;
    }
    void build_edges() {                                                            // Grid::Grid, graph.cpp:371-431
        const int w = gdim[0], d = gdim[1], h = gdim[2];
        dist_goal.assign(dist_goal.size(), -1);      // the lattice changed: cached distance tables are stale
        nbr.assign(exists.size(), {-1, -1, -1, -1, -1, -1});
        nbr_n.assign(exists.size(), 0);
        for (int z = 0; z < h; z++)
            for (int y = 0; y < d; y++)
                for (int x = 0; x < w; x++) {
                    if (!node_exists(x, y, z)) continue;
                    const int id = w * d * z + w * y + x;
                    const int c[6][3] = {{x - 1, y, z}, {x + 1, y, z}, {x, y - 1, z}, {x, y + 1, z}, {x, y, z - 1}, {x, y, z + 1}};
                    int n = 0;
                    const bool v_warn = !warning.empty() && warning[id];
                    for (const auto& q : c)
                        if (node_exists(q[0], q[1], q[2])) {
                            const int wid = w * d * q[2] + w * q[1] + q[0];
                            // no edge from a clear node into a warning node (graph.cpp:389-424: v->warning || !w->warning)
                            if (v_warn || warning.empty() || !warning[wid]) nbr[id][n++] = wid;
                        }
                    nbr_n[id] = (int8_t)n;
                }
    }
};

namespace {

// ---- PIBT (src/mapf/pibt.cpp) on node ids ------------------------------------------------------------------------
struct PibtAgent { int id, v_now, v_next, g; int elapsed, init_d; float tie_breaker; int o; float obs_d; };   // o / obs_d: closest obstacle of interest

struct Pibt {
    dlsc_wp& G;
    int n;
    std::vector<const uint16_t*> dist;             // [agent][node] BFS distance to the agent's goal (solver.cpp:270-290)
    std::vector<std::vector<uint16_t>> own;        // tables of agents without a cache slot
    std::vector<int> occupied_now, occupied_next;  // node -> agent index or -1
    std::vector<PibtAgent> A;
    std::mt19937 mt;
    std::vector<std::vector<int>> plan;            // configurations [t][agent]

    // slot[i]: cache slot (global agent index) of problem agent i, or -1
    Pibt(dlsc_wp& g, const std::vector<int>& start, const std::vector<int>& cur, const std::vector<int>& goal,
         const std::vector<int>& slot = {}, const std::vector<int>& obs_node = {}, const std::vector<float>& obs_dist = {})
        : G(g), n((int)cur.size()), mt(0) {        // DEFAULT_SEED = 0, a fresh generator per problem (problem.cpp:85)
        const int nn = (int)G.exists.size();
        dist.assign(n, nullptr);
        own.resize(n);
        for (int i = 0; i < n; i++) {
            const int sl = slot.empty() ? -1 : slot[i];
            std::vector<uint16_t>& tab = (sl >= 0) ? G.dist_cache[sl] : own[i];
            if (!(sl >= 0 && G.dist_goal[sl] == goal[i] && (int)tab.size() == nn)) {
                tab.assign(nn, (uint16_t)kMaxTimestep);
                std::queue<int> open;
                open.push(goal[i]);
                tab[goal[i]] = 0;
                while (!open.empty()) {
                    const int v = open.front(); open.pop();
                    const int dv = tab[v];
                    for (int k = 0; k < G.nbr_n[v]; k++) {
                        const int m = G.nbr[v][k];
                        if (dv + 1 >= tab[m]) continue;
                        tab[m] = (uint16_t)(dv + 1);
                        open.push(m);
                    }
                }
                if (sl >= 0) G.dist_goal[sl] = goal[i];
            }
            dist[i] = tab.data();
        }
        occupied_now.assign(nn, -1); occupied_next.assign(nn, -1);
        A.resize(n);
        for (int i = 0; i < n; i++) {
            const bool has_o = !obs_node.empty() && obs_node[i] >= 0;
            A[i] = PibtAgent{i, cur[i], -1, goal[i], 0, (int)dist[i][start[i]], (float)i / (float)n, has_o ? obs_node[i] : -1,
                             has_o ? obs_dist[i] : 1e9f};                              // SP_INFINITY when there is none
            occupied_now[cur[i]] = i;
        }
        plan.push_back(cur);
    }
    float node_dist(int u, int v) const {                                           // Pos::euclideanDist, pos.cpp:27-32
        const int w = G.gdim[0], d = G.gdim[1];
        auto xyz = [&](int id, int* o) { o[2] = id / (w * d); o[1] = (id - o[2] * w * d) / w; o[0] = id - o[2] * w * d - o[1] * w; };
        int p[3], q[3];
        xyz(u, p); xyz(v, q);
        const float dx = (float)(p[0] - q[0]), dy = (float)(p[1] - q[1]), dz = (float)(p[2] - q[2]);
        return std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    float goal_dist(const PibtAgent& a, int v) const { return node_dist(a.g, v); }
    float obs_dist_of(const PibtAgent& a, int v) const {                            // PIBT::obsDist :230-236
        const float infinity = 10000;
        if (a.obs_d > infinity) return infinity;
        return node_dist(a.o, v);
    }
    int choose_node(PibtAgent& a) {                                                 // pibt.cpp:142-189
        std::vector<int> C(G.nbr[a.v_now].begin(), G.nbr[a.v_now].begin() + G.nbr_n[a.v_now]);
        C.push_back(a.v_now);
        std::shuffle(C.begin(), C.end(), mt);
        int v = -1;
        for (int u : C) {
            if (occupied_next[u] != -1) continue;                                   // vertex conflict
            const int aj = occupied_now[u];
            if (aj != -1 && A[aj].v_next == a.v_now) continue;                      // swap conflict
            if (u == a.g) return u;
            if (v == -1) { v = u; continue; }
            const int c_v = (int)dist[a.id][v], c_u = (int)dist[a.id][u];
            const float o_v = obs_dist_of(a, v), o_u = obs_dist_of(a, u);           // away from the obstacle of interest
            const float d_v = goal_dist(a, v), d_u = goal_dist(a, u);
            if ((c_u < c_v) || (c_u == c_v && occupied_now[v] != -1 && occupied_now[u] == -1) || (c_u == c_v && o_u > o_v) ||
                (c_u == c_v && occupied_now[v] == -1 && occupied_now[u] == -1 && d_u < d_v))
                v = u;
        }
        return v;
    }
    int plan_one_step(PibtAgent& a) {                                               // :131-139
        const int v = choose_node(a);
        if (v != -1) { occupied_next[v] = a.id; a.v_next = v; }
        return v;
    }
    bool func_pibt(PibtAgent& ai) {                                                 // :105-126
        int v = plan_one_step(ai);
        while (v != -1) {
            const int aj = occupied_now[v];
            if (aj != -1 && aj != ai.id && A[aj].v_next == -1) {
                if (!func_pibt(A[aj])) { v = plan_one_step(ai); continue; }
            }
            return true;
        }
        occupied_next[ai.v_now] = ai.id;
        ai.v_next = ai.v_now;
        return false;
    }
    int run() {                                                                     // :13-103; returns the timesteps made
        auto lower = [this](int x, int y) {                                         // true: x has lower priority than y
            const PibtAgent &a = A[x], &b = A[y];
            if (a.obs_d != b.obs_d) return a.obs_d > b.obs_d;                       // closer to its obstacle of interest first :16
            if (a.elapsed != b.elapsed) return a.elapsed < b.elapsed;
            if (a.init_d != b.init_d) return a.init_d < b.init_d;
            return a.tie_breaker < b.tie_breaker;
        };
        std::priority_queue<int, std::vector<int>, decltype(lower)> undecided(lower);
        std::vector<int> decided;
        for (int i = 0; i < n; i++) undecided.push(i);
        int timestep = 0;
        for (;;) {
            while (!undecided.empty()) {
                const int i = undecided.top(); undecided.pop();
                if (A[i].v_next == -1) func_pibt(A[i]);
                decided.push_back(i);
            }
            bool all_at_goal = true;
            std::vector<int> config(n, -1);
            for (int i : decided) {
                PibtAgent& a = A[i];
                if (occupied_now[a.v_now] == i) occupied_now[a.v_now] = -1;
                occupied_next[a.v_next] = -1;
                config[i] = a.v_next;
                occupied_now[a.v_next] = i;
                all_at_goal = all_at_goal && (a.v_next == a.g);
                a.elapsed = (a.v_next == a.g) ? 0 : a.elapsed + 1;
                a.v_now = a.v_next;
                a.v_next = -1;
                undecided.push(i);
            }
            decided.clear();
            plan.push_back(config);
            ++timestep;
            if (all_at_goal || timestep >= kMaxTimestep) break;
        }
        return timestep;
    }
};

}  // namespace

static thread_local std::string g_wp_err;
static int wp_fail(const std::string& m) { g_wp_err = m; return -1; }

extern "C" {

const char* dlsc_wp_last_error(void) { return g_wp_err.c_str(); }

int dlsc_wp_create(const dlsc_params* p, int n_agents, const float* start, const float* desired_goal, double agent_radius,
                   double agent_downwash, dlsc_wp** out) {
    if (!p || !start || !desired_goal || !out || n_agents < 1) return wp_fail("dlsc_wp_create: bad argument");
    dlsc_wp* w = new dlsc_wp();
    w->N = n_agents; w->dim = p->dim; w->M = p->M;
    w->grid_res = p->grid_res; w->z_2d = p->z_2d; w->comm_range = p->comm_range; w->dt = p->dt;
    w->downwash0 = agent_downwash; w->radius0 = agent_radius;
    for (int i = 0; i < 3; i++) {                                                   // GridBasedPlanner ctor :28-49
        const double gr = w->gres(i);
        const double wmin = (double)(float)p->world_min[i], wmax = (double)(float)p->world_max[i];
        w->gmin[i] = -std::floor((-wmin + kEps) / gr) * gr;
        w->gmax[i] = std::floor((wmax + kEps) / gr) * gr;
    }
    if (w->dim == 2) { w->gmin[2] = w->z_2d; w->gmax[2] = w->z_2d; }
    for (int i = 0; i < w->dim; i++) w->gdim[i] = (int)std::round((w->gmax[i] - w->gmin[i]) / w->gres(i)) + 1;
    if (w->dim == 2) w->gdim[2] = 1;
    if (w->gdim[0] < 1 || w->gdim[1] < 1 || w->gdim[2] < 1) { delete w; return wp_fail("dlsc_wp_create: empty lattice"); }
    w->start.resize(n_agents); w->desired_goal.resize(n_agents);
    for (int a = 0; a < n_agents; a++) { w->start[a] = p3_load(start + 3 * a); w->desired_goal[a] = p3_load(desired_goal + 3 * a); }
    w->exists.assign((size_t)w->gdim[0] * w->gdim[1] * w->gdim[2], 1);
    w->dist_cache.resize(n_agents); w->dist_goal.assign(n_agents, -1);
    w->build_edges();
    *out = w;
    return 0;
}

void dlsc_wp_destroy(dlsc_wp* w) { delete w; }

int dlsc_wp_dims(const dlsc_wp* w, int32_t dims[3]) {
    if (!w || !dims) return wp_fail("dlsc_wp_dims: null argument");
    for (int i = 0; i < 3; i++) dims[i] = w->gdim[i];
    return 0;
}

// Lattice nodes from the distance grid (GridBasedPlanner::updateGridMap, grid_based_planner.cpp:99-159): a node is dropped
// when the L-infinity distance from its point to the cell of the nearest obstacle is below radius - 1e-5.
int dlsc_wp_set_grid(dlsc_wp* w, const float* dist, const int32_t* obst, const int32_t dims[3], const int32_t min_key[3], double res) {
    if (!w) return wp_fail("dlsc_wp_set_grid: null context");
    const size_t nn = w->exists.size();
    if (!dist || !obst) { std::fill(w->exists.begin(), w->exists.end(), 1); w->build_edges(); return 0; }      // no map: every node free
    const float half = (1.0f * (float)0.5) * (float)res;                            // point3d(1, 1, 1) * 0.5 * world_resolution
    const double inv = 1.0 / res;
    for (size_t id = 0; id < nn; id++) {
        const P3 q = w->node_point((int)id);
        const int cx = (int)std::floor(inv * (double)q.x) - min_key[0], cy = (int)std::floor(inv * (double)q.y) - min_key[1],
                  cz = (int)std::floor(inv * (double)q.z) - min_key[2];
        bool keep = true;
        if (cx >= 0 && cx < dims[0] && cy >= 0 && cy < dims[1] && cz >= 0 && cz < dims[2]) {
            const size_t c = ((size_t)cx * dims[1] + cy) * dims[2] + cz;
            if (obst[3 * c] >= 0) {              // an obstacle within the transform's reach (beyond it the accessor's "closest
                                                 // obstacle" is unspecified in dynamicEDT3D: treated as far away)
                const P3 cl = p3((float)(((double)(obst[3 * c] + min_key[0]) + 0.5) * res), (float)(((double)(obst[3 * c + 1] + min_key[1]) + 0.5) * res),
                                 (float)(((double)(obst[3 * c + 2] + min_key[2]) + 0.5) * res));
                const P3 lo = p3(cl.x - half, cl.y - half, cl.z - half), hi = p3(cl.x + half, cl.y + half, cl.z + half);
                P3 cq = q;                        // Box::closestPoint (collision_constraints.cpp:226-237)
                if (q.x < lo.x) cq.x = lo.x; else if (q.x > hi.x) cq.x = hi.x;
                if (q.y < lo.y) cq.y = lo.y; else if (q.y > hi.y) cq.y = hi.y;
                if (q.z < lo.z) cq.z = lo.z; else if (q.z > hi.z) cq.z = hi.z;
                if (linf(q, cq) < w->radius0 - kEpsF) keep = false;
            }
        }
        w->exists[id] = keep ? 1 : 0;
    }
    w->build_edges();
    return 0;
}

// The lattice given directly (occupancy from another source, per-kernel parity tests): dims and one byte per node.
int dlsc_wp_set_nodes(dlsc_wp* w, const int32_t dims[3], const uint8_t* exists) {
    if (!w || !dims || !exists || dims[0] < 1 || dims[1] < 1 || dims[2] < 1) return wp_fail("dlsc_wp_set_nodes: bad argument");
    for (int i = 0; i < 3; i++) w->gdim[i] = dims[i];
    w->exists.assign(exists, exists + (size_t)dims[0] * dims[1] * dims[2]);
    w->build_edges();
    return 0;
}

int dlsc_wp_set_warning(dlsc_wp* w, const uint8_t* warning) {
    if (!w) return wp_fail("dlsc_wp_set_warning: null argument");
    if (warning) w->warning.assign(warning, warning + w->exists.size()); else w->warning.clear();
    w->build_edges();
    return 0;
}

int dlsc_wp_set_obstacles(dlsc_wp* w, int n, const float* pos, const float* vel, const double* radius, const double* max_acc,
                          double uncertainty_horizon) {
    if (!w || n < 0 || (n > 0 && (!pos || !vel || !radius || !max_acc))) return wp_fail("dlsc_wp_set_obstacles: bad argument");
    w->obs_pos.resize(n); w->obs_vel.resize(n); w->obs_radius.assign(radius, radius + n); w->obs_max_acc.assign(max_acc, max_acc + n);
    for (int o = 0; o < n; o++) { w->obs_pos[o] = p3_load(pos + 3 * o); w->obs_vel[o] = p3_load(vel + 3 * o); }
    w->obs_horizon = uncertainty_horizon;
    return 0;
}
int dlsc_wp_set_alerts(dlsc_wp* w, const int32_t* count, const int32_t* ids, int stride) {
    if (!w) return wp_fail("dlsc_wp_set_alerts: null argument");
    w->alerts.assign(w->N, {});
    if (!count || !ids) return 0;
    for (int a = 0; a < w->N; a++)
        for (int k = 0; k < count[a] && k < stride; k++) w->alerts[a].push_back(ids[(size_t)a * stride + k]);
    return 0;
}

int dlsc_wp_get_warning(const dlsc_wp* w, uint8_t* warning) {
    if (!w || !warning) return wp_fail("dlsc_wp_get_warning: null argument");
    for (size_t i = 0; i < w->exists.size(); i++) warning[i] = w->warning.empty() ? 0 : w->warning[i];
    return 0;
}

int dlsc_wp_get_nodes(const dlsc_wp* w, uint8_t* exists) {
    if (!w || !exists) return wp_fail("dlsc_wp_get_nodes: null argument");
    memcpy(exists, w->exists.data(), w->exists.size());
    return 0;
}

// PIBT alone (per-kernel parity entry, mirrors GridBasedPlanner::runMAPF :424-453): node ids in, plan [t][n] out.
int dlsc_wp_pibt(dlsc_wp* w, int n, const int32_t* start, const int32_t* current, const int32_t* goal, int max_t, int32_t* plan_out) {
    if (!w || n < 1 || !start || !current || !goal || !plan_out) return wp_fail("dlsc_wp_pibt: bad argument");
    std::vector<int> s(start, start + n), c(current, current + n), g(goal, goal + n);
    for (int i = 0; i < n; i++)
        for (int v : {s[i], c[i], g[i]})
            if (v < 0 || v >= (int)w->exists.size() || !w->exists[v]) return wp_fail("dlsc_wp_pibt: start / current / goal on a missing node");
    Pibt solver(*w, s, c, g);
    solver.run();
    const int T = (int)solver.plan.size();
    if (T > max_t) return wp_fail("dlsc_wp_pibt: plan longer than the output buffer");
    for (int t = 0; t < T; t++)
        for (int i = 0; i < n; i++) plan_out[t * n + i] = solver.plan[t][i];
    return T;
}

// The same with the closest dynamic obstacle of interest per agent (ProblemAgent's obs_node / obs_dist; obs_node < 0: none).
int dlsc_wp_pibt_obs(dlsc_wp* w, int n, const int32_t* start, const int32_t* current, const int32_t* goal, const int32_t* obs_node,
                     const float* obs_dist, int max_t, int32_t* plan_out) {
    if (!w || n < 1 || !start || !current || !goal || !obs_node || !obs_dist || !plan_out) return wp_fail("dlsc_wp_pibt_obs: bad argument");
    std::vector<int> s(start, start + n), c(current, current + n), g(goal, goal + n), o(obs_node, obs_node + n);
    std::vector<float> od(obs_dist, obs_dist + n);
    for (int i = 0; i < n; i++)
        for (int v : {s[i], c[i], g[i], o[i] < 0 ? s[i] : o[i]})
            if (v < 0 || v >= (int)w->exists.size() || !w->exists[v]) return wp_fail("dlsc_wp_pibt_obs: a node id names a missing node");
    Pibt solver(*w, s, c, g, {}, o, od);
    solver.run();
    const int T = (int)solver.plan.size();
    if (T > max_t) return wp_fail("dlsc_wp_pibt_obs: plan longer than the output buffer");
    for (int t = 0; t < T; t++)
        for (int i = 0; i < n; i++) plan_out[t * n + i] = solver.plan[t][i];
    return T;
}

// One call of MultiSyncSimulator::decentralizedMAPP for the whole swarm.
//   pos [N][3] current positions, goal_cur [N][3] current goal points, traj [N][M][6][3] current desired trajectories or
//   NULL before the first replan (traj.empty()), waypoint [N][3] in: next_waypoint of every agent, out: updated.
int dlsc_wp_step(dlsc_wp* w, const float* pos, const float* goal_cur, const float* traj, float* waypoint) {
    if (!w || !pos || !goal_cur || !waypoint) return wp_fail("dlsc_wp_step: null argument");
    const int N = w->N, M = w->M;
    w->pibt_timesteps = 0;
    // ---- ad-hoc network groups (:310-340) ----
    std::vector<std::set<size_t>> groups;
    groups.push_back({0});
    for (size_t qi = 1; qi < (size_t)N; qi++) {
        int cand = -1;
        int gi = 0;
        while (gi < (int)groups.size()) {
            for (const auto& qj : groups[gi]) {
                const double d = linf(p3_load(pos + 3 * qi), p3_load(pos + 3 * qj));
                if (w->comm_range < 0 || d < w->comm_range) {
                    if (cand == -1) { groups[gi].insert(qi); cand = gi; }
                    else { groups[cand].insert(groups[gi].begin(), groups[gi].end()); groups.erase(groups.begin() + gi); gi--; }
                    break;
                }
            }
            gi++;
        }
        if (cand == -1) groups.push_back({qi});
    }
    // ---- dynamic obstacles (GridBasedPlanner::planMAPF :64-94): warning nodes = lattice points inside an obstacle's reachable
    //      region (updateGridMap :140-150; non-"real" obstacle types, which never remove nodes), BFS tables from every
    //      obstacle's node (updateDistanceTables :165-174).  The same for every group of this step. ----
    const int n_obs = (int)w->obs_pos.size();
    std::vector<std::vector<int>> obs_tab;
    if (n_obs > 0) {
        std::vector<uint8_t> warn(w->exists.size(), 0);
        for (size_t id = 0; id < warn.size(); id++) {
            if (!w->exists[id]) continue;
            const P3 q = w->node_point((int)id);
            for (int o = 0; o < n_obs && !warn[id]; o++) warn[id] = w->obstacle_reaches(o, q) ? 1 : 0;
        }
        if (warn != w->warning) { w->warning = warn; w->build_edges(); }
        for (int o = 0; o < n_obs; o++) obs_tab.push_back(w->bfs_table(w->closest_node(w->obs_pos[o])));
    } else if (!w->warning.empty()) {
        w->warning.clear(); w->build_edges();
    }
    auto obs_cost = [&](const std::set<int>& ids, int node) {                          // getObsCost :685-697
        double cost = 0;
        for (int o : ids) {
            const int d = obs_tab[o][node];
            cost += d == 0 ? 1e9 : 1.0 / ((double)d * d);
        }
        return cost;
    };
    for (const auto& group : groups) {
        const std::vector<size_t> gv(group.begin(), group.end());
        const int n = (int)gv.size();
        // ---- GridBasedPlanner::planMAPF / runMAPF ----
        std::vector<int> s(n), c(n), g(n);
        std::vector<P3> cur_wp(n), goal_pt(n);
        for (int k = 0; k < n; k++) {
            const size_t qi = gv[k];
            cur_wp[k] = p3_load(waypoint + 3 * qi); goal_pt[k] = w->desired_goal[qi];
            s[k] = w->point_id(w->start[qi]); c[k] = w->point_id(cur_wp[k]); g[k] = w->point_id(goal_pt[k]);
            if (!w->exists[s[k]] || !w->exists[c[k]] || !w->exists[g[k]])           // the reference dereferences a null node here
                return wp_fail("dlsc_wp_step: agent " + std::to_string(qi) + ": start, waypoint or goal lies on an occupied lattice node");
        }
        // ---- updateDOI (:192-247) and updateGoal (:250-299): the dynamic obstacles of interest of every agent, and for the
        //      agents that have one an escape goal found by descending the obstacle cost from the agent's node ----
        std::vector<int> obs_node(n, -1);
        std::vector<float> obs_dist(n, 1e9f);
        bool doi_exist = false;
        for (int k = 0; k < n && n_obs > 0; k++) {
            const size_t qi = gv[k];
            const P3 agent_pos = p3_load(pos + 3 * qi);
            std::vector<int> cands;
            if (qi >= w->alerts.size() || w->alerts[qi].empty()) {
                for (int o = 0; o < n_obs; o++) if (w->obstacle_reaches(o, cur_wp[k])) cands.push_back(o);
            } else {
                cands = w->alerts[qi];
            }
            std::set<int> doi;
            double min_dist = 1e9;
            int closest = -1;
            for (int o : cands) {
                if (o < 0 || o >= n_obs) continue;
                doi.insert(o);
                const double d = distance(w->obs_pos[o], agent_pos);
                if (d < min_dist) { min_dist = d; closest = o; }
            }
            if (doi.empty()) continue;
            doi_exist = true;
            obs_node[k] = w->closest_node(w->obs_pos[closest]);
            obs_dist[k] = (float)min_dist;
            // updateGoal
            int nd = w->closest_node(agent_pos);
            const int gnode = w->closest_node(cur_wp[k]);
            P3 new_goal = w->node_point(nd);
            double min_cost = 1e9;
            bool restarted = false;
            std::queue<int> open;
            open.push(nd);
            while (!open.empty()) {
                nd = open.front(); open.pop();
                if (!restarted && nd == gnode) {                                        // restart the search at the waypoint
                    std::queue<int>().swap(open);
                    open.push(gnode);
                    min_cost = 1e9;
                    new_goal = w->node_point(gnode);
                    restarted = true;
                    continue;
                }
                const double c_n = obs_cost(doi, nd);
                for (int e = 0; e < w->nbr_n[nd]; e++) {
                    const int mnode = w->nbr[nd][e];
                    const double c_m = obs_cost(doi, mnode);
                    if (c_n < c_m + kEpsF) continue;
                    if (c_m < min_cost) { min_cost = c_m; new_goal = w->node_point(mnode); }
                    open.push(mnode);
                }
                if (min_cost < 0.01) break;
            }
            goal_pt[k] = new_goal;
            g[k] = w->point_id(new_goal);
        }
        std::vector<int> slot(gv.begin(), gv.end());
        if (doi_exist) slot.assign(n, -1);                  // escape goals change from step to step: no cached distance tables
        Pibt solver(*w, s, c, g, slot, obs_node, obs_dist);
        w->pibt_timesteps += solver.run();
        const auto& plan = solver.plan;
        // ---- updatePlanResult (:292-364) ----
        PlanResult prev = w->plan_result;                                           // planInitialPath (:398-436)
        {
            std::set<size_t> ids(gv.begin(), gv.end());
            const PlanResult& pr = w->plan_result;
            if (!(pr.agent_ids.empty() || pr.agent_ids.size() != (size_t)n || ids != pr.agent_ids)) {
                std::set<int> updated;
                for (int k = 0; k < n; k++)
                    if (pr.paths[k].size() < 2 || norm(sub(pr.paths[k][1], cur_wp[k])) < kEpsF) updated.insert(k);
                if ((int)updated.size() == n) {
                    for (auto& path : prev.paths) if (path.size() > 1) path.erase(path.begin());
                } else {
                    for (int k = 0; k < n; k++)
                        if (pr.paths[k].size() > 1 && updated.count(k)) prev.paths[k][0] = pr.paths[k][1];
                }
            }
        }
        PlanResult now;
        int repeat_start = 0;
        for (int t = 1; t < (int)plan.size(); t++) {
            bool repeat = true;
            for (int k = 0; k < n; k++) if (plan[0][k] != plan[t][k]) { repeat = false; break; }
            if (repeat) repeat_start = t;
        }
        now.paths.resize(n);
        for (int k = 0; k < n; k++) {
            now.agent_ids.insert(gv[k]);
            for (int t = repeat_start; t < (int)plan.size(); t++) now.paths[k].push_back(w->node_point(plan[t][k]));
        }
        auto solution_valid = [&](const PlanResult& r) {                            // isSolutionValid (:380-392)
            if (r.paths.empty()) return false;
            for (int k = 0; k < n; k++) {
                if (k >= (int)r.paths.size() || r.paths[k].empty()) return false;
                if (norm(sub(r.paths[k].back(), goal_pt[k])) > kEpsF) return false;
            }
            return true;
        };
        const bool valid_now = solution_valid(now), valid_prev = solution_valid(prev);
        const bool new_agent_added = now.agent_ids != prev.agent_ids;
        const bool better = now.makespan() < prev.makespan();
        if (!doi_exist && !new_agent_added && (!valid_now || (!better && valid_prev))) now = prev;   // :354-357
        w->plan_result = now;
        // ---- waypoint update rules (:378-452) ----
        std::vector<P3> desired(n);
        for (int k = 0; k < n; k++) {
            const auto& path = w->plan_result.paths[k];
            desired[k] = path[std::min(1, (int)path.size() - 1)];
        }
        std::set<size_t> update_cand;
        for (int k = 0; k < n; k++) {
            const size_t qi = gv[k];
            bool in_range = true;
            if (w->comm_range > 0) {
                for (int m = 0; m < M + 1; m++) {
                    double d;
                    if (!traj) d = linf(desired[k], p3_load(pos + 3 * qi));
                    else if (m < M) d = linf(desired[k], p3_load(traj + ((size_t)qi * M + m) * 18));
                    else d = linf(desired[k], p3_load(traj + ((size_t)qi * M + (M - 1)) * 18 + 15));
                    if (d > 0.5 * w->comm_range - kEpsF) { in_range = false; break; }
                }
            }
            const P3 nw = p3_load(waypoint + 3 * qi), gc = p3_load(goal_cur + 3 * qi);
            // Line(next_waypoint, desired).includePoint(current_goal_point)  (include/geometry.hpp:28-30)
            const bool on_line = std::fabs(distance(gc, nw) + distance(gc, desired[k]) - distance(nw, desired[k])) < kEpsF;
            if (in_range && norm(sub(desired[k], nw)) > kEpsF && on_line) update_cand.insert(qi);
        }
        auto index_of = [&](size_t q) { return (int)(std::lower_bound(gv.begin(), gv.end(), q) - gv.begin()); };
        // "Find valid update" (:418-447): repeatedly drop the first candidate (ascending id) whose desired waypoint
        // coincides with where another agent of the group will be (its desired waypoint if it is still a candidate, else
        // its current one), until none is left.  Same fixed point and same removal order as the reference's triple loop,
        // with the "who will be at this lattice node" question answered from a per-node list instead of a scan.
        if (n > 1 && !update_cand.empty()) {
            std::vector<std::vector<int>> at(w->exists.size());                     // node -> members whose target lies there
            auto target = [&](int kj) { return update_cand.count(gv[kj]) ? desired[kj] : p3_load(waypoint + 3 * gv[kj]); };
            for (int kj = 0; kj < n; kj++) at[w->point_id(target(kj))].push_back(kj);
            auto conflicts = [&](int k) {
                const int node = w->point_id(desired[k]);
                for (int kj : at[node])
                    if (kj != k && distance(desired[k], target(kj)) < kEpsF) return true;
                return false;
            };
            for (bool again = true; again;) {
                again = false;
                for (const auto& qi : update_cand) {
                    const int k = index_of(qi);
                    if (conflicts(k)) {
                        auto& from = at[w->point_id(desired[k])];
                        from.erase(std::find(from.begin(), from.end(), k));
                        update_cand.erase(qi);
                        at[w->point_id(p3_load(waypoint + 3 * qi))].push_back(k);    // it stays where it is
                        again = true;
                        break;
                    }
                }
            }
        }
        for (const auto& qi : update_cand) {
            const int k = index_of(qi);
            waypoint[3 * qi] = desired[k].x; waypoint[3 * qi + 1] = desired[k].y; waypoint[3 * qi + 2] = desired[k].z;
        }
    }
    return 0;
}

int64_t dlsc_wp_pibt_timesteps(const dlsc_wp* w) { return w ? w->pibt_timesteps : 0; }

}  // extern "C"
