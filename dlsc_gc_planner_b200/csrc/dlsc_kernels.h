// dlsc_kernels.h -- launch interface between the context (dlsc_api.cu) and the two kernel
// translation units: dlsc_kernels_exact.cu (-fmad=false: bit-exact float32/float64 stages) and
// dlsc_kernels_qp.cu (FP64 interior-point QP).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dlsc_qp_tables.h"
#include "dlsc_types.h"

namespace dlsc {

// device arrays of one context (local block unless noted)
struct DevState {
    float* rec;                // [N][rec]  (all agents)
    float* acc;                // [NL][3]
    float *stage_pos, *stage_vel;   // [NL][3] staging of dlsc_set_agents
    float* waypoint;           // [NL][3]
    float* goal_new;           // [NL][3] current_goal_point after this step's goal stage (committed to the
                               //         records by dlsc_advance / dlsc_publish_records)
    uint8_t* disturbed;        // [NL]
    uint8_t* sfc_init;         // [NL]
    double *radius, *downwash, *max_vel, *max_acc, *nominal_vel;   // [NL]
    float* pred_traj;          // [N + kMaxDyn][M][P][3]  (rows N.. : constant-velocity predictions of the dynamic obstacles)
    float* init_traj;          // [NL][M][P][3]
    int32_t* nbr_idx;          // [NL][K]
    int32_t* nbr_cnt;          // [NL]
    int* nbr_cell_start;       // [kNbrMaxCells + 1] uniform-grid neighbour search: first slot of every cell in nbr_sorted
    int* nbr_sorted;           // [N] agent indices sorted by cell
    float4* nbr_sorted_pos;    // [N] {x, y, z, group} of the agents in the same order
    float* lsc_normal;         // [NL][K][M][3]
    double* lsc_d;             // [NL][K][M][P]
    float* lsc_anchor_last;    // [NL][K][3]
    uint32_t* lsc_queue;       // [NL K] last-segment items | [NL K (M-1)] GJK items k_lsc could not finish (k_lsc_rest)
    float* lsc_near;           // [NL][K][M] QP row screen: smallest normalised slack of the item's rows at the initial trajectory
    float* sfc;                // [NL][M][6]
    float* traj;               // [NL][M][P][3]  QP result (or failsafe)
    float* traj_host;          // device view of a caller's mapped pinned buffer of the same shape (dlsc_bind_traj_host), or null
    double* qp_x;              // [NL][D][M][P]
    double *cost, *viol;       // [NL]
    int32_t *qp_iters, *status;
    unsigned long long* counters;   // [8]
    double* qp_scratch;        // [qp_ctas][qp_scratch_doubles]
    int* qp_next;              // [4]: [0] light / [3] heavy agents in qp_list_gi, [1] length of qp_list, [2] work counter of k_qp
    int* qp_list;              // [NL] agents queued for the interior-point fallback
    int* qp_list_gi;           // [NL] agents the fast path (k_qp_fast) could not finish: they run the dual active set
    double* qp_seed;           // [NL][4] {violation, row id, screened, violated rows} of the scan at the unconstrained optimum
    // dynamic (non-agent) obstacles (dlsc_set_obstacles)
    float *dyn_pos, *dyn_vel;  // [kMaxDyn][3]
    double *dyn_radius, *dyn_downwash, *dyn_max_acc;   // [kMaxDyn]
    double* dyn_size;          // [kMaxDyn][M][P] predicted sizes (obstacleSizePredictionWithConstAcc)
    float* comm_box;           // [NL][6] CollisionConstraints::communication_range (zero until first built)
    double* qp_slack;          // [NL][kMaxDyn][M] slack variables of the last QP
    uint8_t* trap;             // [NL] checkWaypointTrap outcome of the last step
    EdtDev edt;
};

struct QpLaunch { int ctas, threads; size_t smem; size_t scratch_doubles; size_t gi_smem; size_t fast_smem; int sms; };

void launch_predict(const DevParams& P, const DevState& S, int seq, cudaStream_t st);
int launch_neighbours(const DevParams& P, const DevState& S, cudaStream_t st, int parts = 3);   // 1: grid build, 2: search; returns launches
int launch_lsc(const DevParams& P, const DevState& S, cudaStream_t st);   // returns launches
void launch_sfc(const DevParams& P, const DevState& S, cudaStream_t st);
void launch_goal(const DevParams& P, const DevState& S, cudaStream_t st);
void launch_dyn_predict(const DevParams& P, const DevState& S, cudaStream_t st);   // n_dyn > 0: obstacle predictions, comm boxes, list heads
void launch_trap(const DevParams& P, const DevState& S, cudaStream_t st);          // n_dyn > 0: checkWaypointTrap, before the goal stage
void launch_goal_copy(const DevParams& P, const DevState& S, cudaStream_t st);   // goal_new := record goal
void launch_advance(const DevParams& P, const DevState& S, bool move, cudaStream_t st);
void launch_edt_pack(const float* dist, const int32_t* obst, int4* cells, size_t ncell, cudaStream_t st);
// dlsc_kernels_edt.cu: occupancy raster of CSV boxes and the exact capped EDT with nearest-obstacle index
void launch_edt_raster(const float* boxes_dev, int nb, double res, const int dims[3], const int min_key[3], uint8_t* occ, cudaStream_t st);
int launch_edt_build(const uint8_t* occ, uint32_t* tmp_a, uint32_t* tmp_b, uint8_t* tmp_col, int4* cells, const int dims[3], double res,
                     int maxd, cudaStream_t st);   // returns the number of launches, < 0: unsupported window
void launch_edt_unpack(const int4* cells, float* dist, int32_t* obst, size_t ncell, cudaStream_t st);
void launch_edt_mask(const EdtDev& E, double margin, uint8_t* mask, int* unsafe, cudaStream_t st);
void launch_sat_build(const EdtDev& E, int32_t* sat, cudaStream_t st);   // 4 launches
void launch_expand_anchor(const DevParams& P, const DevState& S, float* anchor_out, cudaStream_t st);
void launch_set_state(const DevParams& P, float* rec, const float* pos, const float* vel, cudaStream_t st);   // pos / vel [NL][3] device, or null
void launch_get_state(const DevParams& P, const float* rec, float* pos, float* vel, cudaStream_t st);         // records -> dense pos / vel
void launch_reset(const DevParams& P, const DevState& S, const float* start_dev, cudaStream_t st);

void launch_gjk_batch(const double* pts, int n, double* v, int32_t* iters, int32_t* simplex, unsigned long long* leaves, cudaStream_t st);

constexpr int kP2PMaxWorld = 16;
void launch_p2p_push(const float* src, size_t n_floats, size_t slice_off_floats, int world, int rank, unsigned long long step,
                     float* const* dst, unsigned long long* const* flag, unsigned* done, cudaStream_t st);
void launch_p2p_wait(const unsigned long long* flags, int world, unsigned long long step, int* err, cudaStream_t st);

double measure_fp64_peak(int device, cudaStream_t st);
QpLaunch qp_launch_config(const DevParams& P, const QpTab& T, int device);
int launch_qp(const DevParams& P, const DevState& S, const QpTab& T, const QpLaunch& L, cudaStream_t st);   // S.qp_next / lists / scratch of this view

}  // namespace dlsc
