// dlsc_qp_tables.h -- constant structure of the piecewise-Bernstein min-jerk QP
// (reference src/traj_optimizer.cpp:172-223 builds Q_base / Aeq_base once in the constructor; here the
// equality constraints are eliminated analytically and every remaining row pattern is tabulated once
// per context so that the kernel is a set of table-driven gathers with no atomics).
//
// Variables.  x[k][m][i]: control point i of segment m, axis k (reference order, traj_optimizer.cpp:230-232).
// The equality rows (initial state :335-352, C0/C1/C2 continuity :338-381, stop-at-end :515-524) leave
//   per axis  nyd = 3M-2  free coordinates  y[3m+j] = x[m][3+j]  (m < M-1),  y[3(M-1)] = x[M-1][3..5]
// and          x[m][0] = y5(m-1),  x[m][1] = 2 y5 - y4,  x[m][2] = 4 y5 - 4 y4 + y3   (m >= 1)
//              x[0][0..2] = c0, c1, c2  (constants from the initial state)
// (equal segment times; same elimination as the CPU oracle, oracle/dlsc_oracle.cpp "XExpr").
//
// Rows.  "pair rows" are two-sided  lo <= a.x <= hi  with a pattern that does not depend on the agent:
//   family 0  box      x[k][m][i]                      world box :251-265, SFC faces :384-410,
//                                                      waypoint comm-range rows :502-512   (merged)
//   family 1  velocity (n/dt)(x[i+1]-x[i])             :452-468
//   family 2  accel.   (n(n-1)/dt^2)(x[i+2]-2x[i+1]+x[i])   :470-487
//   family 3  comm     x[m][n] - x[mi][0], mi <= m     :490-501
// LSC rows (:412-450) couple the D coordinates of one control point and are handled per point.
#pragma once
#include <stdint.h>
#include <vector>
#ifndef __CUDACC__
struct uint4;
#endif

namespace dlsc {

struct QpTabHost {
    int D = 0, M = 0, nyd = 0, ny = 0, npt = 0, nx = 0, np = 0, ntri = 0, ntri_local = 0;
    int use_comm = 0;
    std::vector<int8_t> xm_nv, xm_cidx;          // [npt]
    std::vector<int16_t> xm_idx;                  // [npt][3] axis-local y
    std::vector<double> xm_coef;                  // [npt][3]
    std::vector<uint8_t> pr_fam, pr_axis, pr_nnz; // [np]
    std::vector<int16_t> pr_pt;                   // [np] axis-local point of a box row (else -1)
    std::vector<int16_t> pr_idx;                  // [np][6] global y
    std::vector<double> pr_val;                   // [np][6]
    std::vector<double> pr_cc;                    // [np][3] coefficients on c0,c1,c2 of the row's axis
    std::vector<int> yi_ptr;  std::vector<int16_t> yi_row;  std::vector<double> yi_coef;   // pair rows per global y
    std::vector<int> yp_ptr;  std::vector<int16_t> yp_pt;   std::vector<double> yp_coef;   // points per local y
    std::vector<int> wi_ptr;  std::vector<int16_t> wi_row;  std::vector<double> wi_coef;   // pair rows per W entry (global lower tri)
    std::vector<int> wp_ptr;  std::vector<int16_t> wp_pt;   std::vector<double> wp_coef;   // points per local (a>=b)
    std::vector<uint8_t> tri_p;                   // [ntri] row index p of packed lower-triangle entry e
    std::vector<uint32_t> pr_desc;                // [np] fam | axis<<2 | pa<<4 | pb<<12 (points of the row's stencil)
    std::vector<uint32_t> nz_hdr;                 // [nnzw][4] header of a structurally non-zero entry of W
    std::vector<double> nz_h;                     // [nnzw] its H value
    std::vector<uint16_t> nz_e;                   // packed indices e of the structurally non-zero entries of W
    int nnzw = 0;
    double scv = 0, sca = 0;                      // velocity / acceleration row scales n/dt, n(n-1)/dt^2
    std::vector<double> H1;                       // [nyd][nyd]  2*w_u*sum_m T_m' Q T_m
    std::vector<double> Hinv;                     // [M][nyd][nyd]  inverse of the per-axis Hessian H1 + 2 w_T sum_{m >= M-ts} e e', ts = 1..M
    std::vector<double> Y0;                       // [M][nyd][4]  unconstrained optimum of one axis: y0 = Y0[ts-1] (c0, c1, c2, goal)
    std::vector<double> Q2;                       // [6][6]      2*w_u*Q
    std::vector<double> Qb;                       // [6][6]      Q_base
};

// M, D, dt, weights, comm_range>0 ?
void build_qp_tables(int M, int D, double dt, double w_control, double w_terminal, bool use_comm, QpTabHost& T);

// Q_base = B Z B' dt^(-2 phi + 1)  (traj_optimizer.cpp:172-187, polynomial.hpp:280-293), n = 5, phi = 3
void build_q_base(double dt, double Q[36]);

// Device view (pointers into one device allocation)
struct QpTab {
    int D, M, nyd, ny, npt, nx, np, ntri, ntri_local, use_comm;
    const int8_t *xm_nv, *xm_cidx;
    const int16_t* xm_idx;
    const double* xm_coef;
    const uint8_t *pr_fam, *pr_axis, *pr_nnz;
    const int16_t *pr_pt, *pr_idx;
    const double *pr_val, *pr_cc;
    const int *yi_ptr; const int16_t* yi_row; const double* yi_coef;
    const int *yp_ptr; const int16_t* yp_pt; const double* yp_coef;
    const int *wi_ptr; const int16_t* wi_row; const double* wi_coef;
    const int *wp_ptr; const int16_t* wp_pt; const double* wp_coef;
    const uint8_t* tri_p;
    const uint16_t* nz_e;
    const uint32_t* pr_desc;
    const uint4* nz_hdr;
    const double* nz_h;
    int nnzw;
    int row_npl, row_bv, row_ba, row_bc;   // rows per axis, first velocity / acceleration / comm row inside an axis
    double scv, sca;                       // n/dt, n(n-1)/dt^2
    const double *H1, *Q2, *Hinv, *Y0;
};

}  // namespace dlsc
