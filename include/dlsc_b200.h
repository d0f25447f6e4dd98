/*
 * dlsc_b200.h -- C ABI of libdlsc_b200.so: the B200-native (sm_100a) replan hot path of
 * dlsc_gc_planner:  horizon shift -> neighbour list -> LSC (batched GJK) -> SFC box expansion
 * -> intermediate-goal line search -> piecewise-Bernstein min-jerk QP -> state step.
 *
 * This is the drop-in boundary.  The reference runs the path serially per agent on the CPU:
 *   TrajPlanner::plan / planImpl                      (reference src/traj_planner.cpp:35-63, 108-133)
 *   TrajPlanner::generateDLSCGC                       (src/traj_planner.cpp:603-666)
 *   CollisionConstraints::constructSFCFromInitialTraj (src/collision_constraints.cpp:502-536)
 *   GoalOptimizer::solve                              (src/goal_optimizer.cpp:7-136)
 *   TrajOptimizer::solve                              (src/traj_optimizer.cpp:18-165)
 *   AgentManager::doStep / getAgent                   (src/agent_manager.cpp:34-69, 249-263)
 * Here every agent of a lock-step replan is processed by one dlsc_step() call.  Plain pointers
 * and sizes only; int status returns (0 = ok, <0 = error, text via dlsc_last_error); no
 * exceptions cross the ABI.  Host buffers are caller-owned, device buffers context-owned.
 * There is no CPU fallback: every entry point fails when no CUDA device is usable.
 *
 * Multi-GPU: one context per process/GPU owns the contiguous agent block
 * [agent_begin, agent_begin + n_local) of a swarm of n_agents.  The per-agent "record" (what
 * AgentManager::getAgent exports to the other agents) lives in one device array of n_agents
 * fixed-size records; between steps the caller all-gathers it in place over NCCL
 * (dlsc_records_device / dlsc_record_floats give pointer and stride).
 */
#ifndef DLSC_B200_H
#define DLSC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLSC_ABI_VERSION 4

/* Planner parameters: the subset of MATP::Param / MATP::Mission the hot path reads
 * (reference src/param.cpp:5-117, src/mission.cpp:104-112; launch-file values SURVEY.md s5). */
typedef struct dlsc_params {
    int32_t M;              /* traj/M   segments (2..16)                                   */
    int32_t n;              /* traj/n   polynomial degree; only 5 is supported             */
    int32_t phi;            /* traj/phi; only 3 is supported                               */
    int32_t dim;            /* world/dimension 2|3                                         */
    int32_t use_sfc;        /* world/use_octomap                                           */
    int32_t max_nbr;        /* capacity of the per-agent neighbour list (K)                */
    double dt;              /* traj/dt == multisim/time_step                               */
    double world_min[3];    /* mission world box (float32-representable)                   */
    double world_max[3];
    double world_res;       /* world/resolution                                            */
    double grid_res;        /* grid/resolution                                             */
    double z_2d;            /* world/z_2d                                                  */
    double comm_range;      /* communication/range; <= 0: unlimited                        */
    double w_control;       /* opt/control_input_weight                                    */
    double w_terminal;      /* opt/terminal_weight                                         */
    double reset_threshold; /* plan/reset_threshold                                        */
    int32_t qp_max_iter;    /* interior-point iteration cap (0 -> 80)                      */
    int32_t qp_solver;      /* 0: dual active set, interior-point fallback (default); 1: interior point only;
                               2 / 3: as 0 with the warp-per-agent first scan forced on / off (0 picks by block size) */
    double qp_screen_slack; /* LSC row screens (exact, see dlsc_qp_gi.cuh / dlsc_qp.cuh).  0 -> default (0.5);
                               > 0: the active set skips rows whose slack at the initial trajectory exceeds the
                               iterate's deviation from it, and the interior-point fallback starts from the rows
                               with slack below this value [m]; < 0: every row is evaluated in every scan.  */
    int32_t qp_active_max;  /* the dual active set hands an agent over to the interior point once this many rows are
                               active at the same time (0 -> 32, the capacity; smaller values exercise the hand-over) */
    int32_t reserved0;
} dlsc_params;

/* per-agent status bits (dlsc_get_status) */
#define DLSC_OK               0
#define DLSC_QP_MAXITER       1   /* -> TrajPlanner failsafe: desired_traj = initial_traj  (traj_planner.cpp:749-777) */
#define DLSC_QP_NUMERIC       2
#define DLSC_SFC_INIT_FAILED  4   /* "Invalid initial SFC" (collision_constraints.cpp:445-447) */
#define DLSC_GOAL_INFEASIBLE  8   /* PlanningReport::QPFAILED from GoalOptimizer (goal_optimizer.cpp:122,132) */
#define DLSC_SFC_REUSED      16   /* informational: previous box kept (collision_constraints.cpp:529-532) */
#define DLSC_NBR_OVERFLOW    32   /* more neighbours in range than max_nbr */
#define DLSC_QP_IPM_USED     64   /* informational: the active-set solver gave up, interior point solved it */

/* stage bits for dlsc_run_stages / indices for dlsc_get_timings */
#define DLSC_STAGE_PREDICT  1   /* obstacle prediction + initial trajectory (traj_planner.cpp:290-336, 409-441) */
#define DLSC_STAGE_NBR      2   /* comm-range neighbour list (multi_sync_simulator.cpp:481-503) */
#define DLSC_STAGE_LSC      4
#define DLSC_STAGE_SFC      8
#define DLSC_STAGE_GOAL    16
#define DLSC_STAGE_QP      32
#define DLSC_STAGE_ALL     63
#define DLSC_N_STAGES       6

typedef struct dlsc_ctx dlsc_ctx;

/* Per-step inputs of the local agent block, SoA, host memory.  NULL members keep the device
 * copy (e.g. after dlsc_advance()).  Mirrors the Agent struct handed to TrajPlanner::plan
 * (include/sp_const.hpp:141-160) plus is_disturbed. */
typedef struct dlsc_agents {
    const float* pos;          /* [n_local][3] current_state.position   */
    const float* vel;          /* [n_local][3]                          */
    const float* acc;          /* [n_local][3]                          */
    const float* waypoint;     /* [n_local][3] next_waypoint            */
    const uint8_t* disturbed;  /* [n_local]                             */
} dlsc_agents;

/* Constant per-agent properties (mission JSON "quadrotors"), local block, host memory. */
typedef struct dlsc_agent_props {
    const double* radius;       /* [n_local] */
    const double* downwash;
    const double* max_vel;
    const double* max_acc;      /* first element of the JSON array only (mission.cpp:123-124) */
    const double* nominal_vel;
} dlsc_agent_props;

const char* dlsc_last_error(void);
int dlsc_abi_version(void);
/* CUDA toolkit version the library was built with; exported only by the sm_100a build (loaders use it to refuse
 * anything that is not the CUDA product, e.g. the test-only host simulator). */
int dlsc_cuda_build(void);
int dlsc_device_count(void);

/* Create a context on CUDA device `device` for agents [agent_begin, agent_begin+n_local) of a
 * swarm of n_agents.  Replaces the per-agent TrajPlanner/TrajOptimizer/CollisionConstraints
 * constructors (traj_planner.cpp:4-33, traj_optimizer.cpp:4-16). */
int dlsc_create(const dlsc_params* params, int n_agents, int agent_begin, int n_local, int device,
                dlsc_ctx** out);
void dlsc_destroy(dlsc_ctx* ctx);

/* Run on an existing CUDA stream (cudaStream_t as void*); default: a context-owned stream. */
int dlsc_set_stream(dlsc_ctx* ctx, void* cuda_stream);
void* dlsc_get_stream(dlsc_ctx* ctx);

/* Distance grid: what DynamicEDTOctomap::getDistanceAndClosestObstacle serves
 * (call site collision_constraints.cpp:880; built in map_manager.cpp:61-82).
 * dist [ncell] metres, obst [ncell][3] map-cell index of the nearest occupied cell or -1;
 * cell (x,y,z) -> (x*dims[1]+y)*dims[2]+z; map x = floor(coord/res) - min_key. */
int dlsc_set_edt(dlsc_ctx* ctx, const float* dist, const int32_t* obst, const int32_t dims[3],
                 const int32_t min_key[3], double res);

/* Build the same grid on the device instead (once per mission), replacing MapManager::updateOctreeFromCSV
 * (src/map_manager.cpp:264-316: CSV rows "cx,cy,cz,sx,sy,sz" -> occupied voxels) and MapManager::setGlobalMap
 * (src/map_manager.cpp:61-82: DynamicEDTOctomap(maxdist = 1.0, octree, world_min, world_max).update()).
 * boxes [n_boxes][6] float as parsed from the CSV; maxdist in metres (the reference passes 1.0), capped to
 * int(maxdist/res + 1) <= 16 cells.  Exact Euclidean distances; ties between equidistant occupied cells go to the
 * lowest linear cell index (dynamicEDT3D's own tie-breaking is third-party and unpinned).
 * The grid extent comes from the context's world box (dlsc_edt_dims).  dlsc_build_edt_occupancy takes a ready
 * occupancy grid instead (1 byte per cell, non-zero = occupied, layout as dlsc_set_edt), e.g. from an octomap .bt
 * file expanded by the caller.  dlsc_get_edt reads the current grid back in the dlsc_set_edt layout. */
int dlsc_edt_dims(const dlsc_ctx* ctx, int32_t dims[3], int32_t min_key[3]);
int dlsc_build_edt(dlsc_ctx* ctx, const float* boxes, int n_boxes, double maxdist);
int dlsc_build_edt_occupancy(dlsc_ctx* ctx, const uint8_t* occ, double maxdist);
int dlsc_get_edt(dlsc_ctx* ctx, float* dist, int32_t* obst);
/* device milliseconds of the three distance-transform passes of the last dlsc_build_edt* call */
double dlsc_edt_build_ms(const dlsc_ctx* ctx);

int dlsc_set_agent_props(dlsc_ctx* ctx, const dlsc_agent_props* props);

/* Reset planner state of the local block (planner_seq = 0, initialize_sfc = true,
 * current_goal_point = start, next_waypoint = start: agent_manager.cpp:4-32) and write the
 * local records.  start [n_local][3]. */
int dlsc_reset(dlsc_ctx* ctx, const float* start);

/* Dynamic (non-agent) obstacles: what MultiSyncSimulator::broadcastMsgs puts in front of every agent's obstacle list
 * (src/multi_sync_simulator.cpp:470-480, Obstacle in include/obstacle.hpp:13-37) and TrajPlanner::setObstacles receives.
 * On the device path they get: constant-velocity prediction (src/traj_planner.cpp:303-305), size prediction
 * (obstacleSizePredictionWithConstAcc :338-368), LSCs from normalVectorDynamicObs (:617-627, 1129-1148), the waypoint
 * trap (checkWaypointTrap :708-735), no row in the goal LP (src/goal_optimizer.cpp:176-178), and one slack variable per
 * (obstacle, segment) in the QP (src/traj_optimizer.cpp:272-283, 317-331, 436-448).  Call whenever the obstacle states
 * change (every step for moving obstacles); n = 0 or obs = NULL removes them.  The obstacles occupy the first n slots of
 * every agent's list (entries n_agents + o in dlsc_get_neighbours), so max_nbr must exceed n (n <= DLSC_MAX_OBSTACLES). */
#define DLSC_MAX_OBSTACLES 16
typedef struct dlsc_obstacles {
    int32_t n;
    const float* pos;          /* [n][3] Obstacle::position */
    const float* vel;          /* [n][3] Obstacle::velocity */
    const double* radius;      /* [n] */
    const double* downwash;    /* [n] */
    const double* max_acc;     /* [n] */
} dlsc_obstacles;
typedef struct dlsc_obstacle_params {
    double slack_collision_weight;   /* opt/slack_collision_weight (> 0)   src/param.cpp:80  */
    double uncertainty_horizon;      /* obs/uncertainty_horizon            src/param.cpp:66  */
    int32_t size_prediction;         /* obs/size_prediction                src/param.cpp:65  */
    int32_t reserved0;
} dlsc_obstacle_params;
int dlsc_set_obstacles(dlsc_ctx* ctx, const dlsc_obstacles* obs, const dlsc_obstacle_params* params);
/* after a step: slack variables [n_local][n][M] (<= 0; 0 for agents that kept their initial trajectory), the
 * checkWaypointTrap outcome [n_local], the obstacles' predicted trajectories [n][M][6][3] */
int dlsc_get_slack(dlsc_ctx* ctx, double* slack);
int dlsc_get_trap(dlsc_ctx* ctx, uint8_t* trapped);
int dlsc_get_obstacle_pred(dlsc_ctx* ctx, float* traj);
/* Monte-Carlo batches (BASELINE configs[4]): several independent missions replanning in lockstep inside one
 * context.  group[i] = mission index of local agent i (0 <= group < 2^24); agents only become neighbours of
 * agents of the same mission (the reference runs one MultiSyncSimulator per mission).  Call after dlsc_reset
 * (which puts every agent in mission 0); the index travels inside the agent record, so it is exchanged by
 * the same all-gather as the trajectories. */
int dlsc_set_groups(dlsc_ctx* ctx, const int32_t* group /* [n_local] */);

/* Upload this step's agent states / waypoints (TrajPlanner::plan arguments + setNextWaypoint). */
int dlsc_set_agents(dlsc_ctx* ctx, const dlsc_agents* agents);

/* Records: n_agents x dlsc_record_floats() float32, layout
 *   [0, M*P*3) previous desired trajectory | pos 3 | vel 3 | current_goal 3 | radius | downwash | pad.
 * Single GPU: nothing to do.  Multi GPU: all-gather the local slice
 * (records + agent_begin*stride, n_local*stride floats) in place between steps. */
float* dlsc_records_device(dlsc_ctx* ctx);
int dlsc_record_floats(const dlsc_ctx* ctx);
/* Use caller-allocated device memory for the records (e.g. a torch tensor that NCCL gathers). */
int dlsc_bind_records(dlsc_ctx* ctx, float* device_ptr);
/* Host -> device / device -> host copy of `count` records starting at agent `first` (global index). */
int dlsc_set_records(dlsc_ctx* ctx, int first, int count, const float* host);
int dlsc_get_records(dlsc_ctx* ctx, int first, int count, float* host);

/* Multi-GPU record exchange over NVLink peer memory instead of NCCL (one process per GPU on one node; replaces the
 * copies of MultiSyncSimulator::broadcastMsgs, src/multi_sync_simulator.cpp:468-514).  Every rank moves its records
 * into an IPC-shareable block (dlsc_p2p_export returns its 64-byte cudaIpcMemHandle_t), the handles are exchanged by
 * the caller (any transport) and opened with dlsc_p2p_connect (handles = [world][64] bytes in rank order).
 * dlsc_exchange_records then is the per-step all-gather: a kernel stores the local slice into every rank's record
 * array over NVLink and raises a per-rank step flag, a second kernel waits for the flags of all ranks; stream
 * ordered, no host synchronisation.  Records are double buffered, so ranks may run one exchange apart; every rank
 * must call it the same number of times, and a host-side overwrite of the records (dlsc_set_records) must be
 * separated from the peers' exchanges by dlsc_sync + a barrier of the caller.  dlsc_p2p_status synchronises and
 * reports a peer that never arrived (10 s timeout in the wait kernel).  world <= 16. */
int dlsc_p2p_export(dlsc_ctx* ctx, void* handle64);
int dlsc_p2p_connect(dlsc_ctx* ctx, int world, int rank, const void* handles);
int dlsc_exchange_records(dlsc_ctx* ctx);
int dlsc_p2p_status(dlsc_ctx* ctx);
/* unmap the peers' blocks: call on every rank, then barrier, before any rank calls dlsc_destroy */
int dlsc_p2p_disconnect(dlsc_ctx* ctx);

/* One replan of every local agent (= TrajPlanner::plan for each agent of the block).  Stream
 * ordered; returns after enqueueing.  dlsc_step == dlsc_run_stages(DLSC_STAGE_ALL) + seq++. */
int dlsc_step(dlsc_ctx* ctx);
int dlsc_run_stages(dlsc_ctx* ctx, int stage_mask);
/* AgentManager::doStep (agent_manager.cpp:34-69): state := desired_traj.getStateAt(dt) on the
 * device, and refresh the local records (prev_traj, pos, vel, goal) for the next step. */
int dlsc_advance(dlsc_ctx* ctx);
/* Refresh the local records from the current device state without moving the agents (used when
 * the host supplies the next states through dlsc_set_agents). */
int dlsc_publish_records(dlsc_ctx* ctx);
/* Run the given stages for the local agents [first, first+count) only (used by the per-agent
 * compatibility classes when the caller replans one agent at a time). */
int dlsc_run_stages_subset(dlsc_ctx* ctx, int stage_mask, int first, int count);
int dlsc_sync(dlsc_ctx* ctx);
int dlsc_get_seq(const dlsc_ctx* ctx);
int dlsc_set_seq(dlsc_ctx* ctx, int seq);

/* Results of the last step, local block, to host memory (blocking). */
int dlsc_get_traj(dlsc_ctx* ctx, float* traj /* [n_local][M][P][3] */);
int dlsc_get_qp_x(dlsc_ctx* ctx, double* x /* [n_local][dim][M][P] */);
int dlsc_get_cost(dlsc_ctx* ctx, double* cost /* [n_local] */);
int dlsc_get_violation(dlsc_ctx* ctx, double* viol /* [n_local] */);
int dlsc_get_qp_iters(dlsc_ctx* ctx, int32_t* iters /* [n_local] */);
int dlsc_get_status(dlsc_ctx* ctx, int32_t* status /* [n_local] */);
int dlsc_get_goal(dlsc_ctx* ctx, float* goal /* [n_local][3] current_goal_point */);
/* Deliver the trajectories straight into host memory: `host` is a caller-owned buffer [n_local][M][6][3] floats, pinned
 * (cudaHostAlloc / cudaHostRegister; a pageable buffer is pinned here).  From then on every replan writes each agent's
 * result (TrajOptResult::desired_traj, src/traj_optimizer.cpp:71-83, or the failsafe) to it as soon as that agent's QP
 * finishes, so the device->host transfer overlaps the QPs still running; the buffer is complete after dlsc_sync().
 * dlsc_get_traj keeps working.  NULL unbinds.  The buffer must outlive the binding. */
int dlsc_bind_traj_host(dlsc_ctx* ctx, float* host);
int dlsc_get_state(dlsc_ctx* ctx, float* pos, float* vel, float* acc /* each [n_local][3] or NULL */);
int dlsc_get_init_traj(dlsc_ctx* ctx, float* traj /* [n_local][M][P][3] */);
int dlsc_get_pred_traj(dlsc_ctx* ctx, float* traj /* [n_agents][M][P][3] */);
int dlsc_get_neighbours(dlsc_ctx* ctx, int32_t* idx /* [n_local][K] */, int32_t* cnt /* [n_local] */);
/* LSCs in the oracle / reference layout: normal [n_local][K][M][3], anchor [n_local][K][M][P][3],
 * d [n_local][K][M][P]  (CollisionConstraints::getLSC, collision_constraints.cpp:600-626). */
int dlsc_get_lsc(dlsc_ctx* ctx, float* normal, float* anchor, double* d);
int dlsc_get_sfc(dlsc_ctx* ctx, float* sfc /* [n_local][M][6] min xyz, max xyz */);
int dlsc_set_sfc(dlsc_ctx* ctx, const float* sfc, const uint8_t* init_flag /* [n_local] or NULL */);
/* Stage inputs for single-stage use (per-kernel tests; TrajOptimizer::solve / CollisionConstraints of the
 * compatibility layer): overwrite what the earlier stages would have produced.  Host arrays. */
int dlsc_set_init_traj(dlsc_ctx* ctx, const float* traj /* [n_local][M][P][3] */);
int dlsc_set_pred_traj(dlsc_ctx* ctx, const float* traj /* [n_agents][M][P][3] */);
int dlsc_set_neighbours(dlsc_ctx* ctx, const int32_t* idx /* [n_local][K] */, const int32_t* cnt /* [n_local] */);
/* normal [n_local][K][M][3], anchor_last [n_local][K][3] (anchor of segment M-1), d [n_local][K][M][P];
 * anchors of the segments < M-1 are the neighbours' predicted control points (dlsc_set_pred_traj). */
int dlsc_set_lsc(dlsc_ctx* ctx, const float* normal, const float* anchor_last, const double* d);

/* mean device milliseconds per stage over the steps since the last call (CUDA events recorded on the
 * context's stream around every stage while dlsc_enable_timing(ctx,1); resolved, with one
 * synchronisation, by dlsc_get_timings); order: predict, nbr, lsc, sfc, goal, qp. */
int dlsc_enable_timing(dlsc_ctx* ctx, int on);
int dlsc_get_timings(dlsc_ctx* ctx, double ms[DLSC_N_STAGES], int* n_steps);
/* kernels launched by this context so far */
int64_t dlsc_launch_count(const dlsc_ctx* ctx);
/* work counters of the last step summed over the local block (for roofline arithmetic):
 * [0] neighbour pairs, [1] GJK iterations, [2] lattice vertices of the tested SFC boxes, [3] QP iterations,
 * [4] QP rows, [5] SFC box tests answered from the lattice-vertex mask, [6] from the 16-byte EDT records,
 * [7] of [5]: answered "free" by the O(1) summed-area query alone, [8] lattice vertices of the non-redundant
 * SFC tests (initial box + slabs; SURVEY s8(d) algorithmic figure), [9..15] reserved */
#define DLSC_N_COUNTERS 16
int dlsc_get_counters(dlsc_ctx* ctx, int64_t counters[DLSC_N_COUNTERS]);

/* Waypoints already resident on the device ([n_local][3] float32): stream-ordered device copy. */
int dlsc_set_waypoints_device(dlsc_ctx* ctx, const float* device_ptr);
/* Sustained FP64 FMA rate of the device in TFLOP/s (dependent DFMA chains on every SM for ~ms): the
 * denominator of the FP64 roofline of the LSC and QP kernels (MEASURED_PEAKS.json has no FP64 figure). */
int dlsc_measure_fp64_peak(dlsc_ctx* ctx, double* tflops);

/* Per-kernel parity entry point for the GJK core of the LSC stage (SURVEY s8(b) "single-stage entry points"):
 * closest point of conv{pts[h][0..5]} to the origin for n hulls, by the very device function k_lsc calls
 * (gjk::hull_origin, dlsc_math.cuh).  Replaces closestPointsBetweenPointAndConvexHull -> gjk
 * (include/geometry.hpp:276-306, src/openGJK/openGJK.cpp:674-780).  Host arrays; pts [n][6][3] f64,
 * v [n][3] f64 witness vector, iters / simplex [n] (iteration count, final simplex size; may be NULL),
 * leaves [n] (bit set of the decision-tree leaves the call went through, numbering in dlsc_math.cuh; may be NULL). */
int dlsc_gjk_batch(dlsc_ctx* ctx, const double* pts, int n, double* v, int32_t* iters, int32_t* simplex,
                   uint64_t* leaves);

/* ---- waypoint provider ------------------------------------------------------------------
 * Where every agent's next_waypoint comes from, for agent-only missions: comm-range groups, PIBT on the grid_res lattice
 * and the waypoint update rules.  Replaces MultiSyncSimulator::decentralizedMAPP (src/multi_sync_simulator.cpp:308-466),
 * GridBasedPlanner::planMAPF / runMAPF / updateGridMap / updatePlanResult (src/grid_based_planner.cpp:64-164, 292-453) and
 * MAPF::PIBT (src/mapf/pibt.cpp:13-219).  Host code (a serial priority-inheritance search, once per replan step for the
 * whole swarm); separate handle, same error convention (dlsc_wp_last_error).  start / desired_goal [n_agents][3];
 * agent_radius / agent_downwash: agent 0's, as in the reference (multi_sync_simulator.cpp:375, grid_based_planner.cpp:31). */
typedef struct dlsc_wp dlsc_wp;
const char* dlsc_wp_last_error(void);
int dlsc_wp_create(const dlsc_params* params, int n_agents, const float* start, const float* desired_goal, double agent_radius,
                   double agent_downwash, dlsc_wp** out);
void dlsc_wp_destroy(dlsc_wp* wp);
int dlsc_wp_dims(const dlsc_wp* wp, int32_t dims[3]);                 /* lattice size; node id = w*d*z + w*y + x */
/* Lattice nodes from the distance grid (dlsc_set_edt layout; NULL, NULL: no map, every node free). */
int dlsc_wp_set_grid(dlsc_wp* wp, const float* dist, const int32_t* obst, const int32_t dims[3], const int32_t min_key[3], double res);
int dlsc_wp_set_nodes(dlsc_wp* wp, const int32_t dims[3], const uint8_t* exists /* [w*d*h], non-zero = node present */);
int dlsc_wp_get_nodes(const dlsc_wp* wp, uint8_t* exists /* [w*d*h] */);
/* PIBT alone on node ids (per-kernel parity entry); plan_out [max_t][n]; returns the number of configurations or < 0. */
int dlsc_wp_pibt(dlsc_wp* wp, int n, const int32_t* start, const int32_t* current, const int32_t* goal, int max_t, int32_t* plan_out);
/* per-kernel entries of the dynamic-obstacle side of the planner: warning flags of the lattice nodes (Grid drops the edges
 * from a clear node into a warning node, third_party/grid-pathfinding/graph/src/graph.cpp:371-431; NULL clears them) and
 * PIBT with each agent's closest obstacle of interest (src/mapf/pibt.cpp:16, 185-191, 230-236; obs_node < 0: none) */
int dlsc_wp_set_warning(dlsc_wp* wp, const uint8_t* warning /* [w*d*h] */);
/* Dynamic obstacles of the next dlsc_wp_step calls (n = 0 removes them): GridBasedPlanner::planMAPF's obstacle argument
 * (src/grid_based_planner.cpp:64-94) -- warning nodes (updateGridMap :140-150), obstacle distance tables (:165-174), the
 * obstacles of interest of every agent and its escape goal (updateDOI :192-247, updateGoal :250-299).  Obstacle types other
 * than DYN_REAL (the motion-capture type of the physical experiments, which also removes nodes) are covered.
 * dlsc_wp_set_alerts hands over the agents' collision alerts (TrajOptResult::collision_alert, the obstacle ids; count [N],
 * ids [N][stride]; NULL clears): an agent with alerts takes those obstacles as its candidates instead of the ones that reach
 * its waypoint (:200-226). */
int dlsc_wp_set_obstacles(dlsc_wp* wp, int n, const float* pos, const float* vel, const double* radius, const double* max_acc,
                          double uncertainty_horizon);
int dlsc_wp_set_alerts(dlsc_wp* wp, const int32_t* count, const int32_t* ids, int stride);
int dlsc_wp_get_warning(const dlsc_wp* wp, uint8_t* warning /* [w*d*h] of the last step */);
int dlsc_wp_pibt_obs(dlsc_wp* wp, int n, const int32_t* start, const int32_t* current, const int32_t* goal, const int32_t* obs_node,
                     const float* obs_dist, int max_t, int32_t* plan_out);
/* One decentralizedMAPP call for the whole swarm: pos / goal_cur [N][3], traj [N][M][6][3] (NULL before the first replan),
 * waypoint [N][3] in / out. */
int dlsc_wp_step(dlsc_wp* wp, const float* pos, const float* goal_cur, const float* traj, float* waypoint);
int64_t dlsc_wp_pibt_timesteps(const dlsc_wp* wp);                    /* PIBT timesteps simulated by the last dlsc_wp_step */

/* Device pointers of per-step input / output arrays, for callers that keep data on the GPU. */
float* dlsc_waypoint_device(dlsc_ctx* ctx);   /* [n_local][3] */
float* dlsc_traj_device(dlsc_ctx* ctx);       /* [n_local][M][P][3] */

#ifdef __cplusplus
}
#endif
#endif /* DLSC_B200_H */
