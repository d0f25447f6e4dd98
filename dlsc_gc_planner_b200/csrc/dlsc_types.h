// dlsc_types.h -- plain-old-data shared by host code, kernels and the test-only host simulator.
#pragma once
#include <stdint.h>

namespace dlsc {

constexpr int kP = 6;                // control points per segment (n + 1, n = 5)
constexpr int kMaxM = 16;            // segments
constexpr int kMaxPts = kMaxM * kP;
constexpr int kMaxDyn = 16;          // dynamic (non-agent) obstacles per context

// Device-side copy of the planner parameters plus derived constants.
struct DevParams {
    int M, D, use_sfc, K;            // K = max_nbr
    int N, begin, NL;                // swarm size, first local agent, local count
    int rec;                         // floats per record
    int qp_max_iter;
    double qp_screen;                // LSC working-set screen [m]; <= 0: all rows
    int qp_solver;                   // 0: dual active set with interior-point fallback, 1: interior point only
    int qp_active_max;               // active-set capacity before the hand-over to the interior point (<= kGiQ)
    int n_dyn;                       // dynamic obstacles: they take the first n_dyn obstacle slots of every agent (list entries N + o)
    int dyn_reserved;
    double slack_w;                  // opt/slack_collision_weight
    double dyn_horizon;              // obs/uncertainty_horizon
    double dt, world_res, grid_res, z_2d, comm_range, w_control, w_terminal, reset_threshold;
    double world_min[3], world_max[3];       // double(float(x))
    float tk[kMaxPts];               // (float) of the time accumulated by `time += dt/n` (trajectory.cpp:84-90)
};

// record layout (floats): traj[M*P*3] | pos 3 | vel 3 | goal 3 | radius | downwash | group | pad to x4
// group = mission index of the agent (Monte-Carlo batches: agents only see agents of their own mission)
struct RecLayout {
    int traj, pos, vel, goal, radius, downwash, group, size;
};
inline RecLayout rec_layout(int M) {
    RecLayout r;
    r.traj = 0;
    r.pos = M * kP * 3;
    r.vel = r.pos + 3;
    r.goal = r.vel + 3;
    r.radius = r.goal + 3;
    r.downwash = r.radius + 1;
    r.group = r.downwash + 1;
    r.size = (r.group + 1 + 3) / 4 * 4;
    return r;
}

// EDT grid view: one 16-byte record per cell {dist (float bits), ox, oy, oz}
#ifdef __CUDACC__
struct EdtDev {
    int dims[3];
    int min_key[3];
    double res, inv_res;
    const int4* cells;
    const float* centre[3];      // centre[ax][c] = (float)(((double)(c + min_key[ax]) + 0.5) * res)
    // Lattice-vertex mask (dlsc_stages.cuh, edt_vertex_mask): one byte per lattice vertex v, bit s = outcome of
    // the isObstacleInSFC vertex test when the vertex coordinate falls into cell v - (s&1, s>>1&1, s>>2).
    // Layout [dims0+1][dims1+1][zs], zs = dims2+1 rounded up to 16.  nullptr = not built / not usable.
    const uint8_t* vmask;
    const int32_t* sat;          // summed-area table over "mask byte != 0" [(dims0+2)][(dims1+2)][(dims2+2)], or nullptr
    int zs;
    double mask_margin;          // the margin (agent radius) the mask was built for
};
#endif

// interior-point stopping rule (same constants as the CPU oracle, oracle/dlsc_oracle.cpp ipm_solve):
// primal residual, dual residual relative to (1 + |g|_inf), mean complementarity
#ifndef DLSC_QP_TOL_RD
#define DLSC_QP_TOL_RD 1e-13
#endif
#ifndef DLSC_QP_TOL_MU
#define DLSC_QP_TOL_MU 1e-12
#endif
constexpr double kQpTolRp = 1e-10, kQpTolRd = DLSC_QP_TOL_RD, kQpTolMu = DLSC_QP_TOL_MU;

// status bits (mirror include/dlsc_b200.h)
constexpr int kStQpMaxIter = 1, kStQpNumeric = 2, kStSfcInitFailed = 4, kStGoalInfeasible = 8,
              kStSfcReused = 16, kStNbrOverflow = 32, kStQpIpmUsed = 64;

}  // namespace dlsc
