// compat_driver.cpp -- drives the compatibility classes exactly the way the reference's
// MultiSyncSimulator does (src/multi_sync_simulator.cpp:468-536): broadcast obstacles to every agent
// (setObstacles), then call plan() agent by agent; AgentManager::doStep is restated with the trajectory
// evaluation at t = dt.  Usage: compat_driver <mode> <steps>   mode: serial | staged
// Prints one line per step: max |traj_serial - traj_staged| is checked by the python test by running
// both modes and comparing the printed checksums.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "dlsc_compat.hpp"

using namespace MATP;

static State state_at_dt(const traj_t& tr, double dt, int n) {
    // Trajectory::getStateAt(dt): lands on segment 1 at normalised time 0 (src/trajectory.cpp:111-170)
    State s;
    const Segment<point3d> seg = tr[1];
    s.position = seg[0];
    s.velocity = (seg[1] - seg[0]) * (float)(n / dt);
    const point3d v1 = (seg[2] - seg[1]) * (float)(n / dt);
    s.acceleration = (v1 - s.velocity) * (float)((n - 1) / dt);
    return s;
}

int main(int argc, char** argv) {
    const bool staged = argc > 1 && !strcmp(argv[1], "staged");
    const int steps = argc > 2 ? atoi(argv[2]) : 10;
    Param param;                     // launch/testall_DLSCGC_empty.launch values
    param.M = 5; param.world_dimension = 3; param.world_use_octomap = false; param.communication_range = -1;
    Mission mission;
    const int N = 8;
    mission.qn = N;
    mission.world_min = point3d(-3.f, -3.f, 0.f); mission.world_max = point3d(3.f, 3.f, 2.5f);
    for (int i = 0; i < N; i++) {
        Agent a;
        a.id = i; a.cid = i + 1;
        const float ang = 6.2831853f * i / N;
        a.start_point = point3d(2.f * std::cos(ang), 2.f * std::sin(ang), 1.f);
        a.desired_goal_point = point3d(-2.f * std::cos(ang), -2.f * std::sin(ang), 1.f);   // antipodal swap
        a.current_state.position = a.start_point;
        a.current_goal_point = a.start_point; a.next_waypoint = a.start_point;
        mission.agents.push_back(a);
    }
    ros::NodeHandle nh;
    std::vector<std::unique_ptr<TrajPlanner>> planners;
    std::vector<Agent> agents = mission.agents;
    std::vector<traj_t> trajs(N);
    for (int i = 0; i < N; i++) planners.emplace_back(new TrajPlanner(nh, param, mission, agents[i]));
    std::shared_ptr<octomap::OcTree> octree;
    std::shared_ptr<DynamicEDTOctomap> distmap;
    for (int step = 0; step < steps; step++) {
        // waypoint stand-in: straight toward the goal, at most 0.5 m ahead of the current goal point
        for (int i = 0; i < N; i++) {
            point3d d = agents[i].desired_goal_point - agents[i].current_goal_point;
            const double len = d.norm();
            agents[i].next_waypoint = len > 0.5 ? agents[i].current_goal_point + d * (float)(0.5 / len) : agents[i].desired_goal_point;
        }
        // broadcastMsgs: every agent receives the others as obstacles (getAgent(), agent_manager.cpp:249-263)
        for (int i = 0; i < N; i++) {
            Obstacles obs;
            for (int j = 0; j < N; j++) {
                if (j == i) continue;
                Obstacle o;
                o.type = ObstacleType::AGENT; o.id = j;
                o.position = agents[j].current_state.position; o.velocity = agents[j].current_state.velocity;
                o.goal_point = agents[j].current_goal_point;
                o.radius = (float)agents[j].radius; o.downwash = (float)agents[j].downwash; o.max_acc = (float)agents[j].max_acc;
                o.prev_traj = trajs[j];
                obs.push_back(o);
            }
            planners[i]->setObstacles(obs);
        }
        if (staged)
            for (int i = 0; i < N; i++) TrajPlanner::stageAgent(param, mission, agents[i], false);
        double checksum = 0, cost = 0;
        for (int i = 0; i < N; i++) {
            TrajOptResult r = planners[i]->plan(agents[i], octree, distmap, ros::Time(), false);
            trajs[i] = r.desired_traj;
            agents[i].current_goal_point = planners[i]->getCurrentGoalPosition();
            cost += r.total_qp_cost;
            for (int m = 0; m < param.M; m++)
                for (int k = 0; k <= param.n; k++) checksum += (m + 1) * (k + 1) * ((double)r.desired_traj[m][k].x() + 2.0 * r.desired_traj[m][k].y() + 3.0 * r.desired_traj[m][k].z());
        }
        for (int i = 0; i < N; i++) agents[i].current_state = state_at_dt(trajs[i], param.dt, param.n);   // doStep
        double dmin = 1e9;
        for (int i = 0; i < N; i++)
            for (int j = i + 1; j < N; j++) {
                point3d d = agents[i].current_state.position - agents[j].current_state.position;
                d.z() = d.z() / 2.0f;
                dmin = std::min(dmin, d.norm());
            }
        printf("step %d checksum %.9f cost %.9f min_dist %.6f seq %d\n", step, checksum, cost, dmin, planners[0]->getPlannerSeq());
    }
    {   // PlanningStatistics is filled from the device stage timers; a second mission with the same agent count gets its
        // own context; a dynamic (non-agent) obstacle changes the plan
        const PlanningStatistics ps = planners[0]->getPlanningStatistics();
        printf("stats seq %d total %.3e qp %.3e samples %d\n", ps.planning_seq, ps.planning_time.total_planning_time.average,
               ps.planning_time.traj_optimization_time.average, ps.planning_time.total_planning_time.N_sample);
        Mission other = mission;
        other.agents[0].desired_goal_point = point3d(0.5f, 0.5f, 1.f);
        TrajPlanner p2(nh, param, other, other.agents[0]);
        printf("contexts distinct %d same %d\n", p2.deviceContext() != planners[0]->deviceContext() ? 1 : 0,
               planners[1]->deviceContext() == planners[0]->deviceContext() ? 1 : 0);
        // a dynamic (non-agent) obstacle in the list goes to the device path: same agent planned with and without an obstacle
        // that flies through its start position (not near its waypoint: with comm range <= 0 checkWaypointTrap drops those)
        Mission third = mission;
        third.agents[0].desired_goal_point = point3d(0.5f, 0.6f, 1.f);
        param.slack_threshold = 0.001;                      // src/param.cpp:104 default
        TrajPlanner p3(nh, param, third, third.agents[0]);
        double cost[2], endy[2];
        int alerts[2] = {0, 0}, alert_id[2] = {-1, -1};
        TrajPlanner* pp[2] = {&p2, &p3};
        Mission* mm[2] = {&other, &third};
        for (int w = 0; w < 2; w++) {
            Agent a0 = mm[w]->agents[0];
            a0.next_waypoint = point3d(1.53f, 0.16f, 1.f);
            Obstacles obs;
            for (int j = 1; j < N; j++) {
                Obstacle o;
                o.type = ObstacleType::AGENT; o.id = j; o.position = mm[w]->agents[j].start_point; o.goal_point = o.position;
                o.radius = (float)mm[w]->agents[j].radius; o.downwash = (float)mm[w]->agents[j].downwash;
                obs.push_back(o);
            }
            if (w == 0) {
                Obstacle d;
                d.type = ObstacleType::DYN_SPIN; d.id = 0; d.position = point3d(2.f, -0.6f, 1.f); d.velocity = point3d(0.f, 0.5f, 0.f);
                d.radius = 0.3; d.downwash = 1.0; d.max_acc = 0.0;
                obs.insert(obs.begin(), d);                       // broadcastMsgs lists the dynamic obstacles first
            }
            pp[w]->setObstacles(obs);
            TrajOptResult r = pp[w]->plan(a0, octree, distmap, ros::Time(), false);
            cost[w] = r.total_qp_cost; endy[w] = r.desired_traj.lastPoint().y();
            alerts[w] = (int)r.collision_alert.obstacles.size();
            alert_id[w] = alerts[w] ? r.collision_alert.obstacles[0].id : -1;
        }
        printf("dynamic obstacle planned cost %.6f vs %.6f end_y %.4f vs %.4f alerts %d (id %d) vs %d\n", cost[0], cost[1], endy[0], endy[1],
               alerts[0], alert_id[0], alerts[1]);
        // the obstacle list a caller passes IS the neighbour list (a caller that filters differently from broadcastMsgs):
        // agent 0 of a fourth mission is told about agents 2 and 5 only
        Mission fourth = mission;
        fourth.agents[0].desired_goal_point = point3d(0.4f, 0.7f, 1.f);
        TrajPlanner p4(nh, param, fourth, fourth.agents[0]);
        Obstacles few;
        for (int j : {5, 2}) {
            Obstacle o;
            o.type = ObstacleType::AGENT; o.id = j; o.position = fourth.agents[j].start_point; o.goal_point = o.position;
            o.radius = (float)fourth.agents[j].radius; o.downwash = (float)fourth.agents[j].downwash;
            few.push_back(o);
        }
        p4.setObstacles(few);
        Agent a4 = fourth.agents[0];
        a4.next_waypoint = point3d(1.53f, 0.16f, 1.f);
        p4.plan(a4, octree, distmap, ros::Time(), false);
        std::vector<int32_t> nidx((size_t)N * (N - 1 + DLSC_MAX_OBSTACLES)), ncnt(N);
        dlsc_get_neighbours(p4.deviceContext(), nidx.data(), ncnt.data());
        printf("caller list kept %d\n", ncnt[0] == 2 && nidx[0] == 2 && nidx[1] == 5 ? 1 : 0);
    }
    // TrajOptimizer::solve with explicit constraints (one LSC plane, no SFC)
    {
        Agent a = mission.agents[0];
        a.current_goal_point = point3d(1.f, 0.f, 1.f); a.next_waypoint = a.current_goal_point;
        CollisionConstraints cons(param, mission, a.radius, a.max_vel);
        Obstacles obs(1);
        obs[0].type = ObstacleType::AGENT; obs[0].id = 1; obs[0].position = point3d(1.5f, 0.f, 1.f);
        cons.initializeLSC(obs);
        for (int m = 0; m < param.M; m++)
            for (int i = 0; i <= param.n; i++) cons.setLSC(0, m, i, LSC(point3d(1.5f, 0.f, 1.f), point3d(1.f, 0.f, 0.f), 0.4));   // x >= 1.9: active (free optimum ends at 1.878)
        traj_t init(param.M, param.n, param.dt);
        for (int m = 0; m < param.M; m++)
            for (int i = 0; i <= param.n; i++) init[m][i] = a.current_state.position;
        Eigen::MatrixXd B;
        TrajOptimizer opt(param, mission, B);
        TrajOptResult r = opt.solve(a, cons, init, true);
        const point3d e = r.desired_traj.lastPoint();
        printf("solve end %.6f %.6f %.6f cost %.9f\n", e.x(), e.y(), e.z(), r.total_qp_cost);
    }
    return 0;
}
