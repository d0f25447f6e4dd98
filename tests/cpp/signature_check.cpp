// signature_check.cpp -- compile-time check of the drop-in boundary: every public member SURVEY.md s8(b) lists is taken by a
// pointer of EXACTLY the reference's type (include/traj_planner.hpp:42-60, include/traj_optimizer.hpp:18-33,
// include/collision_constraints.hpp:118-171).  Built twice by tests/test_compat_cpp.py: against the stand-in third-party
// types (DLSC_COMPAT_STANDALONE) and against tests/cpp/realtree/ (ros / octomap / dynamicEDT3D / Eigen headers with the real
// APIs' spelling), which also compiles the live-DynamicEDTOctomap export of SwarmBatch::set_distmap.
#include <memory>
#include <type_traits>

#include "dlsc_compat.hpp"

using namespace MATP;

#define SAME(expr, ...) static_assert(std::is_same<decltype(expr), __VA_ARGS__>::value, #expr)

// TrajPlanner
static_assert(std::is_constructible<TrajPlanner, const ros::NodeHandle&, const Param&, const Mission&, const Agent&>::value, "TrajPlanner ctor");
SAME(&TrajPlanner::plan, TrajOptResult (TrajPlanner::*)(const Agent&, const std::shared_ptr<octomap::OcTree>&,
                                                          const std::shared_ptr<DynamicEDTOctomap>&, ros::Time, bool));
SAME(&TrajPlanner::publish, void (TrajPlanner::*)());
SAME(&TrajPlanner::setObstacles, void (TrajPlanner::*)(const Obstacles&));
SAME(&TrajPlanner::getPlannerSeq, int (TrajPlanner::*)() const);
SAME(&TrajPlanner::getCurrentGoalPosition, point3d (TrajPlanner::*)() const);
SAME(&TrajPlanner::getPlanningStatistics, PlanningStatistics (TrajPlanner::*)() const);
// TrajOptimizer
static_assert(std::is_constructible<TrajOptimizer, const Param&, const Mission&, const Eigen::MatrixXd&>::value, "TrajOptimizer ctor");
SAME(&TrajOptimizer::solve, TrajOptResult (TrajOptimizer::*)(const Agent&, const CollisionConstraints&, const traj_t&, bool));
SAME(&TrajOptimizer::updateParam, void (TrajOptimizer::*)(const Param&));
SAME(&TrajOptResult::desired_traj, traj_t TrajOptResult::*);
SAME(&TrajOptResult::total_qp_cost, double TrajOptResult::*);
SAME(&TrajOptResult::collision_alert, CollisionAlert TrajOptResult::*);
// CollisionConstraints
static_assert(std::is_constructible<CollisionConstraints, const Param&, const Mission&, double, double>::value, "CollisionConstraints ctor");
SAME(&CollisionConstraints::initializeSFC, void (CollisionConstraints::*)(const point3d&));
SAME(&CollisionConstraints::initializeLSC, void (CollisionConstraints::*)(const Obstacles&));
SAME(&CollisionConstraints::constructSFCFromPoint, void (CollisionConstraints::*)(const point3d&, const point3d&));
SAME(&CollisionConstraints::constructSFCFromConvexHull, void (CollisionConstraints::*)(const point3ds&, const point3d&));
SAME(&CollisionConstraints::constructSFCFromInitialTraj, void (CollisionConstraints::*)(const traj_t&, const point3d&, const point3d&));
SAME(&CollisionConstraints::constructCommunicationRange, void (CollisionConstraints::*)(const point3d&));
SAME(&CollisionConstraints::isDynamicObstacle, bool (CollisionConstraints::*)(int) const);
SAME(&CollisionConstraints::isPointInFeasibleRegion, bool (CollisionConstraints::*)(const point3d&, int, int) const);
SAME(&CollisionConstraints::getLSC, LSC (CollisionConstraints::*)(int, int, int) const);
SAME(&CollisionConstraints::getSFC, Box (CollisionConstraints::*)(int) const);
SAME(&CollisionConstraints::getObsSize, size_t (CollisionConstraints::*)() const);
SAME(&CollisionConstraints::getObsPosition, point3d (CollisionConstraints::*)(int) const);
SAME(&CollisionConstraints::setDistmap, void (CollisionConstraints::*)(std::shared_ptr<DynamicEDTOctomap>));
SAME(&CollisionConstraints::setOctomap, void (CollisionConstraints::*)(std::shared_ptr<octomap::OcTree>));
SAME(&CollisionConstraints::setLSC, void (CollisionConstraints::*)(int, int, int, const LSC&));
SAME(&CollisionConstraints::setSFC, void (CollisionConstraints::*)(int, const Box&));
// value types
SAME(&LSC::obs_control_point, point3d LSC::*);
SAME(&LSC::normal_vector, point3d LSC::*);
SAME(&LSC::d, double LSC::*);
SAME(&Box::box_min, point3d Box::*);
SAME(&Box::box_max, point3d Box::*);

#ifndef DLSC_COMPAT_STANDALONE
// real-tree mode: one map mission through the unchanged call pattern, the grid coming from a live DynamicEDTOctomap.
// argv: dist.bin obst.bin dims[3] min_key[3]  (written by the python test from the oracle's grid); prints the SFC-bound
// trajectory checksum of 3 replans, which the test compares with the same mission driven through dlsc_set_edt directly.
#include <cstdio>
#include <cstdlib>
#include <vector>
int main(int argc, char** argv) {
    if (argc < 9) return 2;
    int32_t dims[3] = {atoi(argv[3]), atoi(argv[4]), atoi(argv[5])}, mk[3] = {atoi(argv[6]), atoi(argv[7]), atoi(argv[8])};
    const size_t nc = (size_t)dims[0] * dims[1] * dims[2];
    std::vector<float> dist(nc);
    std::vector<int32_t> obst(nc * 3);
    FILE* f = fopen(argv[1], "rb"); if (!f || fread(dist.data(), 4, nc, f) != nc) return 3; fclose(f);
    f = fopen(argv[2], "rb"); if (!f || fread(obst.data(), 4, nc * 3, f) != nc * 3) return 3; fclose(f);
    auto distmap = std::make_shared<DynamicEDTOctomap>(dist, obst, dims, mk, 0.1);
    auto octree = std::make_shared<octomap::OcTree>(0.1);
    Param param;                     // launch/simulation.launch (maze, 2-D)
    param.M = 10; param.world_dimension = 2; param.world_use_octomap = true; param.communication_range = 3.0;
    Mission mission;
    const int N = argc > 9 ? atoi(argv[9]) : 2;
    mission.qn = N;
    mission.world_min = point3d((float)atof(argv[10]), (float)atof(argv[11]), (float)atof(argv[12]));
    mission.world_max = point3d((float)atof(argv[13]), (float)atof(argv[14]), (float)atof(argv[15]));
    for (int i = 0; i < N; i++) {
        Agent a;
        a.id = i;
        a.start_point = point3d((float)atof(argv[16 + 6 * i]), (float)atof(argv[17 + 6 * i]), (float)atof(argv[18 + 6 * i]));
        a.desired_goal_point = point3d((float)atof(argv[19 + 6 * i]), (float)atof(argv[20 + 6 * i]), (float)atof(argv[21 + 6 * i]));
        a.current_state.position = a.start_point; a.current_goal_point = a.start_point; a.next_waypoint = a.start_point;
        mission.agents.push_back(a);
    }
    ros::NodeHandle nh;
    std::vector<std::unique_ptr<TrajPlanner>> planners;
    for (int i = 0; i < N; i++) planners.emplace_back(new TrajPlanner(nh, param, mission, mission.agents[i]));
    std::vector<Agent> agents = mission.agents;
    std::vector<traj_t> trajs(N);
    for (int step = 0; step < 3; step++) {
        for (int i = 0; i < N; i++) {
            Obstacles obs;
            for (int j = 0; j < N; j++) {
                if (j == i) continue;
                Obstacle o;
                o.type = ObstacleType::AGENT; o.id = j; o.position = agents[j].current_state.position;
                o.velocity = agents[j].current_state.velocity; o.goal_point = agents[j].current_goal_point;
                o.radius = (float)agents[j].radius; o.downwash = (float)agents[j].downwash; o.prev_traj = trajs[j];
                obs.push_back(o);
            }
            planners[i]->setObstacles(obs);
        }
        double checksum = 0;
        for (int i = 0; i < N; i++) {
            TrajOptResult r = planners[i]->plan(agents[i], octree, distmap, ros::Time(), false);
            trajs[i] = r.desired_traj;
            agents[i].current_goal_point = planners[i]->getCurrentGoalPosition();
            for (int m = 0; m < param.M; m++)
                for (int k = 0; k <= param.n; k++) checksum += (m + 1) * (k + 1) * ((double)r.desired_traj[m][k].x() + 2.0 * r.desired_traj[m][k].y());
        }
        for (int i = 0; i < N; i++) {
            agents[i].current_state.position = trajs[i][1][0];
            agents[i].next_waypoint = agents[i].start_point;      // hover mission: the test is about the map binding
        }
        printf("step %d checksum %.9f\n", step, checksum);
    }
    return 0;
}
#else
int main() { return 0; }
#endif
