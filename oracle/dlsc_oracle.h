/*
 * dlsc_oracle.h -- C ABI of the CPU ORACLE for the dlsc_gc_planner replan hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a plain, serial, CPU restatement of the
 * reference algorithm (file:line citations are in dlsc_oracle.cpp).  It is the checker
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  Nothing under dlsc_gc_planner_b200/ (the product) may include, link or call it.
 *
 * Parity pinning status (see DESIGN.md "Oracle"):
 *   - GJK core: pinned bit-for-bit against the reference's own openGJK.cpp compiled
 *     unmodified into oracle/_ref/ (tests/test_oracle_gjk_ref.py).
 *   - composed path (LSC+SFC+goal+QP+state step): pinned against the reference's golden
 *     log log/result_1742185870.978562_DLSCGC_10agents.csv (tests/test_oracle_golden.py).
 *   - QP objective value / dynamicEDT3D tie-breaking / CPLEX tolerances: parity unpinned
 *     (third-party code absent from /root/reference); cross-checked against HiGHS.
 *   - dynamic (non-agent) obstacle path: parity unpinned (the reference holds no run, vector or test
 *     with dynamic obstacles); restated line by line, slack QP cross-checked against HiGHS
 *     (tests/test_dynamic_obstacles.py).
 */
#ifndef DLSC_ORACLE_H
#define DLSC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Planner parameters: the subset of MATP::Param / MATP::Mission the hot path reads
 * (reference src/param.cpp:5-117, src/mission.cpp:104-112). */
typedef struct orc_params {
    int32_t M;            /* traj/M   number of segments                         */
    int32_t n;            /* traj/n   polynomial degree (only 5 supported)       */
    int32_t phi;          /* traj/phi (only 3 supported)                         */
    int32_t dim;          /* world/dimension 2|3                                 */
    int32_t use_sfc;      /* world/use_octomap                                   */
    int32_t reserved0;
    double dt;            /* traj/dt == multisim/time_step                       */
    double world_min[3];  /* mission world box (float32-representable values)    */
    double world_max[3];
    double world_res;     /* world/resolution                                    */
    double grid_res;      /* grid/resolution                                     */
    double z_2d;          /* world/z_2d                                          */
    double comm_range;    /* communication/range, <=0: unlimited                 */
    double w_control;     /* opt/control_input_weight                            */
    double w_terminal;    /* opt/terminal_weight                                 */
    double reset_threshold; /* plan/reset_threshold                              */
} orc_params;

/* Euclidean distance grid with nearest-obstacle index: the data DynamicEDTOctomap holds.
 * cell (x,y,z) -> linear (x*ny + y)*nz + z.  dist in metres (float), obst = map-cell
 * coordinates of the nearest occupied cell or (-1,-1,-1) when none within max dist. */
typedef struct orc_edt {
    int32_t dims[3];
    int32_t min_key[3];   /* floor(world_min / res) : map x = floor(coord/res) - min_key */
    double res;
    const float* dist;
    const int32_t* obst;  /* [ncell][3] */
} orc_edt;

/* status bits returned per agent by orc_step / stage functions */
#define ORC_OK                0
#define ORC_QP_MAXITER        1
#define ORC_QP_NUMERIC        2
#define ORC_SFC_INIT_FAILED   4   /* "Invalid initial SFC" (collision_constraints.cpp:445) */
#define ORC_GOAL_INFEASIBLE   8   /* goal LP infeasible and not a numerical error         */
#define ORC_SFC_REUSED       16   /* informational: previous box reused (:529-532)        */
#define ORC_NBR_OVERFLOW     32

/* ---- GJK (openGJK.cpp:674-780 with body 2 = {origin}) ---- */
double orc_gjk_hull_origin(const double* pts, int npts, double v[3], int* iters, int* simplex_n);

/* ---- geometry (geometry.hpp:184-274) : float points in, out ---- */
void orc_closest_segments(const float a0[3], const float a1[3], const float b0[3], const float b1[3],
                          float p1[3], float p2[3], double* dist);

/* ---- horizon shift (traj_planner.cpp:290-336, 409-441) ----
 * init_traj: own view, pred_traj: how others see the agent.  [N][M][P][3] float */
void orc_predict(const orc_params* p, int N, int seq, const float* pos, const float* vel,
                 const float* prev_traj, const uint8_t* disturbed, float* init_traj, float* pred_traj);

/* neighbour lists by Chebyshev range, ascending index (multi_sync_simulator.cpp:481-503) */
int orc_neighbours(const orc_params* p, int N, const float* pos, int max_nbr, int32_t* nbr_idx,
                   int32_t* nbr_cnt);

/* ---- LSC (traj_planner.cpp:603-666) ----
 * out: normal [N][K][M][3] f32, anchor [N][K][M][P][3] f32, d [N][K][M][P] f64 (K = max_nbr) */
void orc_lsc_batch(const orc_params* p, int N, const float* init_traj, const float* pred_traj,
                   const float* goal_cur, const double* radius, const double* downwash, int max_nbr,
                   const int32_t* nbr_idx, const int32_t* nbr_cnt, float* normal, float* anchor,
                   double* d, int64_t* gjk_iter_hist /* [32] or NULL */);

/* ---- SFC (collision_constraints.cpp:435-452, 502-536, 781-901, 1023-1093) ---- */
int orc_sfc_expand(const orc_params* p, const orc_edt* edt, const float box_in[6], double margin,
                   double max_vel, float box_out[6], int64_t* n_lookups);
void orc_sfc_batch(const orc_params* p, const orc_edt* edt, int N, const uint8_t* init_flag,
                   const float* pos, const float* init_traj, const float* goal_cur,
                   const float* waypoint, const double* radius, const double* max_vel,
                   float* sfc /* [N][M][6] in/out */, int32_t* status, int64_t* n_lookups);

/* ---- goal line search (goal_optimizer.cpp:7-136, 138-198) ---- */
void orc_goal_batch(const orc_params* p, int N, const uint8_t* disturbed, const float* pos,
                    const float* waypoint, const float* sfc, int max_nbr, const int32_t* nbr_cnt,
                    const float* normal, const float* anchor, const double* d,
                    float* goal_cur /* in/out */, int32_t* status);

/* ---- QP (traj_optimizer.cpp:225-551) for one agent ----
 * traj_out [M][P][3] f32 (solution truncated to float like :71-83), x_out [dim][M][P] f64 */
int orc_qp_solve(const orc_params* p, const float pos[3], const float vel[3], const float acc[3],
                 const float goal[3], const float waypoint[3], double radius, double max_vel,
                 double max_acc, double nominal_vel, const float* sfc /* [M][6] */, int K,
                 const float* normal, const float* anchor, const double* d, const float* init_traj,
                 float* traj_out, double* x_out, double* cost, double* max_violation, int* iters);

/* full lock-step replan of all agents, serial per agent like multi_sync_simulator.cpp:516-524.
 * n_threads>1 uses one agent per thread (OpenMP) for the QP/LSC/SFC loops. */
typedef struct orc_step_io {
    int N;
    int seq;                 /* planner_seq after increment (1 on the first replan)  */
    int max_nbr;
    int n_threads;
    /* inputs */
    const float* pos;        /* [N][3] */
    const float* vel;
    const float* acc;
    const float* waypoint;   /* next_waypoint [N][3] */
    const uint8_t* disturbed;/* [N] */
    const double* radius;    /* [N] */
    const double* downwash;
    const double* max_vel;
    const double* max_acc;
    const double* nominal_vel;
    const orc_edt* edt;      /* may be NULL when !use_sfc */
    /* state, in/out */
    float* goal_cur;         /* current_goal_point [N][3] */
    float* prev_traj;        /* in: previous desired traj; out: new desired traj [N][M][P][3] */
    float* sfc;              /* [N][M][6] */
    uint8_t* sfc_init_flag;  /* [N] initialize_sfc flag (traj_planner.cpp:20-23, 439, 693-695) */
    /* outputs */
    float* init_traj;        /* [N][M][P][3] */
    float* pred_traj;
    int32_t* nbr_idx;        /* [N][max_nbr] */
    int32_t* nbr_cnt;        /* [N] */
    float* lsc_normal;       /* [N][K][M][3] */
    float* lsc_anchor;       /* [N][K][M][P][3] */
    double* lsc_d;           /* [N][K][M][P] */
    double* qp_x;            /* [N][dim][M][P] double solution */
    double* cost;            /* [N] */
    double* max_violation;   /* [N] */
    int32_t* qp_iters;       /* [N] */
    int32_t* status;         /* [N] */
    double* stage_seconds;   /* [5]: predict+nbr, lsc, sfc, goal, qp (wall) or NULL */
    /* dynamic (non-agent) obstacles, n_dyn = 0 when absent.  They take the first n_dyn obstacle slots of every agent
     * (list entries N + o), like the obstacle list MultiSyncSimulator::broadcastMsgs builds (multi_sync_simulator.cpp:476-480). */
    int n_dyn;
    int dyn_size_prediction;          /* obs/size_prediction                    */
    double dyn_uncertainty_horizon;   /* obs/uncertainty_horizon                */
    double slack_collision_weight;    /* opt/slack_collision_weight             */
    const float* dyn_pos;             /* [n_dyn][3] Obstacle::position          */
    const float* dyn_vel;             /* [n_dyn][3]                             */
    const double* dyn_radius;         /* [n_dyn]                                */
    const double* dyn_downwash;
    const double* dyn_max_acc;
    float* comm_box;         /* state in/out [N][6]: CollisionConstraints::communication_range (zero until first built) or NULL */
    double* qp_slack;        /* out [N][n_dyn][M] slack variables of the QP, or NULL */
    uint8_t* trap;           /* out [N] checkWaypointTrap found the waypoint trapped, or NULL */
} orc_step_io;
void orc_step(const orc_params* p, orc_step_io* io);
/* replan only agents [a_begin, a_end) (bounded CPU-baseline samples of large swarms) */
void orc_step_range(const orc_params* p, orc_step_io* io, int a_begin, int a_end);

/* Trajectory::getStateAt (trajectory.cpp:111-170): state[9] = pos, vel, acc (float) */
void orc_state_at(const orc_params* p, const float* traj, double t, float state[9]);

/* CSV boxes -> occupancy (map_manager.cpp:264-316) -> exact EDT capped like
 * DynamicEDTOctomap(maxdist=1.0) (map_manager.cpp:75-79).  boxes [nb][6] = cx,cy,cz,sx,sy,sz (float
 * values as parsed).  dims/min_key out; dist/obst caller-allocated after orc_edt_dims(). */
void orc_edt_dims(const orc_params* p, int32_t dims[3], int32_t min_key[3]);
void orc_edt_build(const orc_params* p, int nb, const float* boxes, float* dist, int32_t* obst);

/* QP constants for inspection: Q_base [P][P] (traj_optimizer.cpp:172-187) */
void orc_q_base(const orc_params* p, double* Q);

#ifdef __cplusplus
}
#endif
#endif
