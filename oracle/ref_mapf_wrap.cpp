/*
 * ref_mapf_wrap.cpp -- C-ABI shim around the REFERENCE's own PIBT solver (compiled unmodified from
 * /root/reference/src/mapf/{pibt,solver,problem,plan,paths}.cpp and third_party/grid-pathfinding/graph by
 * oracle/Makefile into oracle/_ref/).  Test infrastructure only: pins the waypoint layer's PIBT restatement
 * (dlsc_gc_planner_b200/csrc/dlsc_waypoints.cpp).  Mirrors GridBasedPlanner::runMAPF (src/grid_based_planner.cpp:424-453):
 * a Grid of the existing nodes, one ProblemAgent per agent, MAPF::PIBT(&P).solve(), the whole plan back.
 */
#include <cstdint>
#include <cstring>
#include <mapf/pibt.hpp>

using namespace MAPF;

// exists [w*d*h], id = w*d*z + w*y + x; start / current / goal: node ids [n].
// plan_out [max_t][n] node ids; returns the number of configurations of the plan (0 = failed), -1 when it does not fit.
extern "C" int ref_pibt_solve(int w, int d, int h, const uint8_t* exists, int n, const int* start, const int* current,
                              const int* goal, int max_t, int* plan_out) {
    Nodes V(w * d * h, nullptr);
    for (int z = 0; z < h; z++)
        for (int y = 0; y < d; y++)
            for (int x = 0; x < w; x++) {
                const int id = w * d * z + w * y + x;
                if (exists[id]) V[id] = new Node(id, x, y, z, false);
            }
    Grid* grid = new Grid(V, w, d, h);
    ProblemAgents agents;
    Node* any = nullptr;
    for (auto v : V) if (v) { any = v; break; }
    for (int i = 0; i < n; i++)      // no dynamic obstacle of interest: obstacle node unused, distance "infinite" (sp_const.hpp SP_INFINITY)
        agents.emplace_back(grid->getNode(start[i]), grid->getNode(current[i]), grid->getNode(goal[i]), any, 1000000007.0f);
    Problem P(grid, n, agents);
    PIBT solver(&P);
    solver.solve();
    Plan plan = solver.getSolution();
    const int T = plan.size();
    int rc = T;
    if (T > max_t) rc = -1;
    else
        for (int t = 0; t < T; t++)
            for (int i = 0; i < n; i++) plan_out[t * n + i] = plan.get(t, i)->id;
    return rc;      // Problem's destructor frees the grid and its nodes
}

// The same with dynamic obstacles: `warning` [w*d*h] marks the nodes inside an obstacle's reachable region (Grid then drops
// the edges from a clear node into a warning node, graph.cpp:371-431); obs_node / obs_dist per agent = the closest
// obstacle of interest and its distance (ProblemAgent, problem.hpp:9-19; obs_node < 0: none, distance SP_INFINITY).
extern "C" int ref_pibt_solve_obs(int w, int d, int h, const uint8_t* exists, const uint8_t* warning, int n, const int* start,
                                  const int* current, const int* goal, const int* obs_node, const float* obs_dist, int max_t,
                                  int* plan_out) {
    Nodes V(w * d * h, nullptr);
    for (int z = 0; z < h; z++)
        for (int y = 0; y < d; y++)
            for (int x = 0; x < w; x++) {
                const int id = w * d * z + w * y + x;
                if (exists[id]) V[id] = new Node(id, x, y, z, warning && warning[id]);
            }
    Grid* grid = new Grid(V, w, d, h);
    ProblemAgents agents;
    Node* any = nullptr;
    for (auto v : V) if (v) { any = v; break; }
    for (int i = 0; i < n; i++)
        agents.emplace_back(grid->getNode(start[i]), grid->getNode(current[i]), grid->getNode(goal[i]),
                            obs_node[i] >= 0 ? grid->getNode(obs_node[i]) : any, obs_node[i] >= 0 ? obs_dist[i] : 1e9f);
    Problem P(grid, n, agents);
    PIBT solver(&P);
    solver.solve();
    Plan plan = solver.getSolution();
    const int T = plan.size();
    int rc = T;
    if (T > max_t) rc = -1;
    else
        for (int t = 0; t < T; t++)
            for (int i = 0; i < n; i++) plan_out[t * n + i] = plan.get(t, i)->id;
    return rc;
}
