// dlsc_qp.cuh -- batched piecewise-Bernstein min-jerk QP, one CTA per agent.
//
// Replaces TrajOptimizer::solve / populatebyrow + CPLEX (reference src/traj_optimizer.cpp:18-165,
// 225-527).  Same problem, different machinery:
//   * equality rows are eliminated analytically (dlsc_qp_tables.h) -> ny = D(3M-2) unknowns;
//   * dense primal-dual interior point (Mehrotra predictor-corrector) on  min 1/2 y'Hy + g'y,  G y <= h;
//   * the reduced KKT matrix  W = H + G' diag(z/s) G  (ny <= 128) lives in shared memory as a packed
//     lower triangle, is assembled by table-driven gathers (no atomics), factorised in place as
//     L D L' (one __syncthreads per column) and solved by one warp with register-resident right-hand
//     sides and warp shuffles;  FP64 pipe, no tensor cores (systems of 39..84 unknowns).
//   * LSC rows (57 x neighbours per agent, almost none active) go through an exact working-set screen:
//     the QP is solved on the rows whose slack at the starting point is below a threshold, every row is
//     then checked at the solution and, if one is violated, the solve is repeated on a wider set.  A
//     relaxed optimum that is feasible for the full problem is the full optimum, so nothing is approximated.
//   * row state (slack s, dual z, directions) stays in a per-CTA global scratch slab that is reused for
//     every agent the CTA processes (L1/L2 resident); the compact LSC row list is sorted by control point
//     so that its contributions are accumulated per point without atomics (deterministic).
//
// The core is __host__ __device__: the device build runs it with a 128-thread CTA, the test-only host
// simulator with a single "thread" (tests/hostsim), so the arithmetic can be checked without a GPU.
#pragma once
#include "dlsc_math.cuh"
#include "dlsc_qp_tables.h"
#include "dlsc_types.h"

#ifndef DLSC_DYN_TAU_MULT
#define DLSC_DYN_TAU_MULT 2.0
#endif
namespace dlsc {

// ------------------------------------------------------------------------------------------------
// CTA abstraction
// ------------------------------------------------------------------------------------------------
struct Cta {
    int tid, nthr;
    double* red;     // >= 3 * 32 doubles of shared scratch
#ifdef DLSC_QP_CYCLES
    long long* ticks = nullptr;                    // diagnostic build: cycles of thread 0 per phase; ticks[0] = last stamp
    DLSC_HD void tick(int k) const {
#ifdef __CUDA_ARCH__
        if (ticks) { const long long now = clock64(); ticks[k] += now - ticks[0]; ticks[0] = now; }
#endif
    }
#else
    DLSC_HD void tick(int) const {}
#endif
    bool warp = false;   // true: the "CTA" is one warp (nthr == 32): warp-level synchronisation only, red unused
    DLSC_HD void sync() const {
#ifdef __CUDA_ARCH__
        if (warp) __syncwarp(); else __syncthreads();
#endif
    }
    DLSC_HD void wsync() const {      // threads of one warp only
#ifdef __CUDA_ARCH__
        __syncwarp();
#endif
    }
    // op: 0 sum, 1 max, 2 min.  Deterministic (fixed tree).  All threads get the result.
    DLSC_HD void reduce3(double& a, int opa, double& b, int opb, double& c, int opc) const {
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ta = __shfl_xor_sync(0xffffffffu, a, o);
            const double tb = __shfl_xor_sync(0xffffffffu, b, o);
            const double tc = __shfl_xor_sync(0xffffffffu, c, o);
            a = comb(a, ta, opa); b = comb(b, tb, opb); c = comb(c, tc, opc);
        }
        if (warp) return;
        const int w = tid >> 5, nw = nthr >> 5;
        __syncthreads();
        if ((tid & 31) == 0) { red[w] = a; red[32 + w] = b; red[64 + w] = c; }
        __syncthreads();
        a = red[0]; b = red[32]; c = red[64];
        for (int i = 1; i < nw; i++) { a = comb(a, red[i], opa); b = comb(b, red[32 + i], opb); c = comb(c, red[64 + i], opc); }
#endif
    }
    // one value only (same tree)
    DLSC_HD void reduce1(double& a, int opa) const {
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a = comb(a, __shfl_xor_sync(0xffffffffu, a, o), opa);
        if (warp) return;
        const int w = tid >> 5, nw = nthr >> 5;
        __syncthreads();
        if ((tid & 31) == 0) red[w] = a;
        __syncthreads();
        a = red[0];
        for (int i = 1; i < nw; i++) a = comb(a, red[i], opa);
#endif
    }
    // (value, id): the largest value, ties to the smallest id.  All threads get the result.
    // m: a third value max-reduced, n: a fourth value summed, on the same tree.
    DLSC_HD void reduce_argmax(double& v, double& id, double& m, double& n) const {
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double tv = __shfl_xor_sync(0xffffffffu, v, o);
            const double ti = __shfl_xor_sync(0xffffffffu, id, o);
            const double tm = __shfl_xor_sync(0xffffffffu, m, o);
            const double tn = __shfl_xor_sync(0xffffffffu, n, o);
            if (tv > v || (tv == v && ti < id)) { v = tv; id = ti; }
            m = tm > m ? tm : m;
            n += tn;
        }
        if (warp) return;
        const int w = tid >> 5, nw = nthr >> 5;
        __syncthreads();
        if ((tid & 31) == 0) { red[w] = v; red[32 + w] = id; red[64 + w] = m; red[16 + w] = n; }   // nw <= 16
        __syncthreads();
        v = red[0]; id = red[32]; m = red[64]; n = red[16];
        for (int i = 1; i < nw; i++) {
            const double tv = red[i], ti = red[32 + i], tm = red[64 + i];
            if (tv > v || (tv == v && ti < id)) { v = tv; id = ti; }
            m = tm > m ? tm : m;
            n += red[16 + i];
        }
#endif
    }
    static DLSC_HD double comb(double x, double y, int op) {
        return op == 0 ? x + y : (op == 1 ? (x < y ? y : x) : (y < x ? y : x));
    }
};

// ------------------------------------------------------------------------------------------------
// per-agent inputs / outputs
// ------------------------------------------------------------------------------------------------
struct QpIn {
    V3 pos, vel, acc, goal, wp;
    double radius, max_vel, max_acc, nominal_vel;
    const float* sfc;          // [M][6]
    const float* init_traj;    // [M][P][3]
    int K;                     // neighbour count
    const int32_t* nbr_idx;    // [K] global indices
    const float* normal;       // [K][M][3]
    const double* d;           // [K][M][P]
    const float* anchor_last;  // [K][3]
    const float* pred_traj;    // [N][M][P][3] (anchors of segments < M-1)
    const float* near;         // [K][M] row screen of k_lsc (or null): smallest normalised slack of the item's rows at the
                               //        initial trajectory; the rows cannot be violated while |x_pt - init_pt| stays below it
};
struct QpOut {
    float* traj;               // [M][P][3]
    float* traj_host;          // the same slice of a caller's mapped host buffer (dlsc_bind_traj_host), or null
    double* x;                 // [D][M][P] or null
    double* cost; double* viol; int32_t* iters; int32_t* status;
    long long* rows;           // active inequality rows (one-sided count) or null
    double* slack;             // [n_dyn][M] slack variables of the dynamic-obstacle rows, or null
};

// shared-memory carve-up (doubles unless noted)
struct QpSmem {
    double *W, *invp, *pan, *y, *dy, *rd, *x, *dx, *ax1, *ax2, *V1, *V2, *DD, *S, *cst, *red;
    float* sfcs;               // dual active set only: staged copy of the agent's SFC boxes [M][6]
    double* dev;               // dual active set only: [M] largest |x_pt - init_pt| per segment of the current iterate
    double* esl;               // dual active set only: [kMaxDyn][M] slack variables of the dynamic-obstacle rows (or null: all zero)
    int* off;                  // [npt + 2] segment offsets of the LSC row list (+ scratch word)
    uint8_t* act;              // [M][Kcap]
};
// dual active-set (Goldfarb-Idnani, Schur-complement form) working storage; aliases the W region
constexpr int kGiQ = 32;                               // max simultaneously active rows
constexpr int kGiTri = kGiQ * (kGiQ + 1) / 2;
struct GiSmem {
    double *Hinv;                                      // [nyd][nyd] of this agent's terminal-segment count
    double *Ls, *Sm;                                   // packed lower: Cholesky of S = A H^-1 A', and S itself
    double *yc;                                        // [kGiQ+1][9] y-space coefficients of the active rows (+ candidate)
    double *bq, *u, *r, *v, *l;                        // [kGiQ+1]
    double *li;                                        // [kGiQ+1] reciprocals of the Cholesky diagonal (dlsc_qp_gi.cuh)
    double *ty;                                        // scalars: [0] step, [1] flag
    int16_t* yi;                                       // [kGiQ+1][9] y indices (-1 = unused)
    int* id;                                           // [kGiQ+1] row ids
};
// Q: capacity of the active set (rows); a template parameter so that a larger-capacity instantiation costs nothing to add --
// one was measured for the slack-heavy agents of dynamic-obstacle missions and lost to the interior point (profiles/r2_history.md)
template <int Q = kGiQ>
DLSC_HD size_t gi_doubles(const QpTab& T) {
    return (size_t)T.nyd * T.nyd + 2 * (Q * (Q + 1) / 2) + 9 * (Q + 1) + 6 * (Q + 1) + 8 + (9 * (Q + 1) + 3) / 4 + (Q + 2) / 2 + 2;
}
template <int Q = kGiQ>
DLSC_HD void gi_carve(const QpTab& T, double* base, GiSmem& g) {
    double* p = base;
    g.Hinv = p; p += T.nyd * T.nyd;
    g.Ls = p; p += Q * (Q + 1) / 2; g.Sm = p; p += Q * (Q + 1) / 2;
    g.yc = p; p += 9 * (Q + 1);
    g.bq = p; p += Q + 1; g.u = p; p += Q + 1; g.r = p; p += Q + 1; g.v = p; p += Q + 1; g.l = p; p += Q + 1; g.li = p; p += Q + 1;
    g.ty = p; p += 8;
    g.yi = reinterpret_cast<int16_t*>(p); p += (9 * (Q + 1) + 3) / 4;
    g.id = reinterpret_cast<int*>(p);
}
DLSC_HD size_t qp_w_doubles(const QpTab& T) { return (size_t)T.ntri > gi_doubles(T) ? (size_t)T.ntri : gi_doubles(T); }

DLSC_HD size_t qp_smem_doubles(const QpTab& T) {
    const size_t v12 = (2 * (size_t)T.np > 8 * (size_t)T.ny) ? 2 * (size_t)T.np : 8 * (size_t)T.ny;   // V1|V2, aliased by pan
    return qp_w_doubles(T) + 4 * (size_t)T.ny + 16 + 4 * (size_t)T.nx + v12 + (size_t)T.np + 6 * (size_t)T.npt + 16 + 96 +
           ((size_t)T.npt + 4) / 2 + 1;
}
DLSC_HD size_t qp_smem_bytes(const QpTab& T, int Kcap) {
    return qp_smem_doubles(T) * sizeof(double) + (((size_t)T.M * Kcap + 15) / 16) * 16;
}
DLSC_HD void qp_smem_carve(const QpTab& T, int Kcap, double* base, QpSmem& s) {
    double* p = base;
    s.W = p; p += qp_w_doubles(T);
    s.invp = p; p += T.ny + 16; s.y = p; p += T.ny; s.dy = p; p += T.ny; s.rd = p; p += T.ny;
    s.x = p; p += T.nx; s.dx = p; p += T.nx; s.ax1 = p; p += T.nx; s.ax2 = p; p += T.nx;
    {   // V1 and V2 are dead while W is being factorised: the panel buffers of ldl_factor alias them
        const size_t v12 = (2 * (size_t)T.np > 8 * (size_t)T.ny) ? 2 * (size_t)T.np : 8 * (size_t)T.ny;
        s.V1 = p; s.V2 = p + T.np; s.pan = p; p += v12;
    }
    s.DD = p; p += T.np;
    s.S = p; p += 6 * T.npt;
    s.cst = p; p += 16;
    s.red = p; p += 96;
    s.sfcs = nullptr; s.dev = nullptr; s.esl = nullptr;
    s.off = reinterpret_cast<int*>(p); p += (T.npt + 4) / 2 + 1;
    s.act = reinterpret_cast<uint8_t*>(p);
    (void)Kcap;
}
// per-CTA global scratch (doubles): LSC row list [npt*Kcap] x {n0,n1,n2,b,s,z,c,ds,dz} + point index (int);
// pair rows [np] x 13
// + dynamic obstacles: slack group of every LSC row (int) and the per-slack block (kDynDoubles per slack variable)
constexpr int kDynSlots = 18;                          // y unknowns a slack group touches: 3 axes x (3 of segment m-1, 3 of segment m)
constexpr int kDynDoubles = 11 + kDynSlots + 3;        // e, de, sb, zb, dsb, dzb, ccb, wee, rhe, rde, pb | u[18] | row index per point (6 ints)
DLSC_HD size_t qp_scratch_doubles(const QpTab& T, int Kcap) {
    const size_t LS = (size_t)T.npt * Kcap;
    return 9 * LS + (LS + 1) / 2 + 13 * (size_t)T.np + (LS + 1) / 2 + (size_t)kDynDoubles * kMaxDyn * T.M;
}
// y index of local slot (axis k, slot) of a slack group of segment m, or -1: slots 0..2 = free points of segment m-1
// (control points 0..2 of segment m depend on them through the continuity rows), slots 3..5 = free points of segment m
// (one unknown for the last segment)
DLSC_HD int dyn_slot_y(const QpTab& T, int m, int k, int slot) {
    if (slot < 3) return m >= 1 ? k * T.nyd + 3 * (m - 1) + slot : -1;
    if (m == T.M - 1) return slot == 3 ? k * T.nyd + 3 * (T.M - 1) : -1;
    return k * T.nyd + 3 * m + slot - 3;
}

// ------------------------------------------------------------------------------------------------
// L D L' of the packed lower triangle W (row-major: (i,k) at i(i+1)/2+k), in place, rank-4 panels.
// After the call column j holds the unscaled entries Lt(i,j) = l_ij * d_j and invp[j] = 1/d_j.
//   panel   : every thread owns one row i >= j0 of the 4-column panel and eliminates it against the
//             4x4 diagonal block, which it factors redundantly from shared memory (no barrier inside
//             the panel); unscaled and scaled panel rows go to pu / ps (4 doubles per row).
//   trailing: A(i,k) -= sum_t pu(i,t) ps(k,t), one warp per row, lanes over k  (4 FMA per load/store).
// Two barriers per 4 columns.  Returns false (uniformly) when a pivot is not positive.
// pan: 8*ny doubles of shared scratch.
// ------------------------------------------------------------------------------------------------
DLSC_HD bool ldl_factor(const Cta& c, double* W, double* invp, double* pan, int ny, const uint8_t* tri_p) {
    double* pu = pan;
    double* ps = pan + 4 * ny;
    double* blk = invp + ny;          // 16 doubles: [0..9] eliminated block (row-major lower), [10..13] 1/pivot, [14] ok
    for (int j0 = 0; j0 < ny; j0 += 4) {
        const int nb = (ny - j0 < 4) ? ny - j0 : 4;
        // ---- 4x4 diagonal block: eliminated by one thread, broadcast through shared memory ----
        double d[4][4];     // d[s][t], s >= t: unscaled block columns
        double ip[4];
        if (c.tid == 0) {
            bool ok = true;
#pragma unroll
            for (int s2 = 0; s2 < 4; s2++)
#pragma unroll
                for (int t = 0; t <= s2; t++) d[s2][t] = (s2 < nb) ? W[(j0 + s2) * (j0 + s2 + 1) / 2 + j0 + t] : (s2 == t ? 1.0 : 0.0);
#pragma unroll
            for (int t = 0; t < 4; t++) {
#pragma unroll
                for (int s2 = t; s2 < 4; s2++) {
                    double v = d[s2][t];
#pragma unroll
                    for (int u = 0; u < t; u++) v -= d[s2][u] * (d[t][u] * ip[u]);
                    d[s2][t] = v;
                }
                if (!(d[t][t] > 0)) ok = false;
                ip[t] = 1.0 / d[t][t];
            }
#pragma unroll
            for (int s2 = 0; s2 < 4; s2++)
#pragma unroll
                for (int t = 0; t <= s2; t++) blk[s2 * (s2 + 1) / 2 + t] = d[s2][t];
#pragma unroll
            for (int t = 0; t < 4; t++) blk[10 + t] = ip[t];
            blk[14] = ok ? 1.0 : 0.0;
        }
        c.sync();
#pragma unroll
        for (int s2 = 0; s2 < 4; s2++)
#pragma unroll
            for (int t = 0; t <= s2; t++) d[s2][t] = blk[s2 * (s2 + 1) / 2 + t];
#pragma unroll
        for (int t = 0; t < 4; t++) ip[t] = blk[10 + t];
        if (blk[14] == 0.0) return false;
        // ---- panel rows ----
        for (int i = j0 + c.tid; i < ny; i += c.nthr) {
            const int row = i * (i + 1) / 2 + j0;
            double a[4];
#pragma unroll
            for (int t = 0; t < 4; t++) a[t] = (t < nb && j0 + t <= i) ? W[row + t] : 0.0;
#pragma unroll
            for (int t = 1; t < 4; t++) {
                double v = a[t];
#pragma unroll
                for (int u = 0; u < t; u++) v -= a[u] * (d[t][u] * ip[u]);
                a[t] = v;
            }
#pragma unroll
            for (int t = 0; t < 4; t++) {
                // rows inside the diagonal block are written after the barrier (other threads may still be
                // reading the block from W)
                if (t < nb && i >= j0 + nb) W[row + t] = a[t];
                const bool below = i > j0 + t;          // strictly below the diagonal of column j0+t
                pu[i * 4 + t] = (below && t < nb) ? a[t] : 0.0;
                ps[i * 4 + t] = (below && t < nb) ? a[t] * ip[t] : 0.0;
            }
            if (i < j0 + nb) invp[i] = ip[i - j0];
        }
        c.sync();
#pragma unroll
        for (int s2 = 0; s2 < 4; s2++)
#pragma unroll
            for (int t = 0; t <= s2; t++)
                if (s2 < nb && (c.nthr == 1 || c.tid == s2)) W[(j0 + s2) * (j0 + s2 + 1) / 2 + j0 + t] = d[s2][t];
        // ---- trailing update: one 4x4 tile of A(i,k) -= sum_t pu(i,t) ps(k,t) per thread (64 FMA per
        //      16 loads of W; the panel rows of the tile's 4 rows / 4 columns sit in registers) ----
        const int k0 = j0 + nb;
        if (k0 < ny) {
            const int I0 = k0 >> 2, nT = ((ny + 3) >> 2) - I0, ntile = nT * (nT + 1) / 2;
            for (int t = c.tid; t < ntile; t += c.nthr) {
                const int Ir = tri_p[t], Kr = t - Ir * (Ir + 1) / 2;
                const int i0 = (I0 + Ir) << 2, kk0 = (I0 + Kr) << 2;
                double u[4][4], q[4][4];
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int t2 = 0; t2 < 4; t2++) {
                        u[r][t2] = (i0 + r < ny) ? pu[(i0 + r) * 4 + t2] : 0.0;
                        q[r][t2] = (kk0 + r < ny) ? ps[(kk0 + r) * 4 + t2] : 0.0;
                    }
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int i = i0 + r;
                    if (i < ny) {
                        const int row = i * (i + 1) / 2;
#pragma unroll
                        for (int cc = 0; cc < 4; cc++) {
                            const int k = kk0 + cc;
                            if (k <= i) W[row + k] -= u[r][0] * q[cc][0] + u[r][1] * q[cc][1] + u[r][2] * q[cc][2] + u[r][3] * q[cc][3];
                        }
                    }
                }
            }
        }
        c.sync();
    }
    return true;
}

// Solve (L D L') w = b in place.  Device: warp 0 only (caller syncs before and after).
DLSC_HD void ldl_solve(const Cta& c, const double* W, const double* invp, double* b, int ny) {
#ifdef __CUDA_ARCH__
    if (c.tid >= 32) return;
    const int lane = c.tid;
    double r0 = (lane < ny) ? b[lane] : 0.0, r1 = (lane + 32 < ny) ? b[lane + 32] : 0.0;
    double r2 = (lane + 64 < ny) ? b[lane + 64] : 0.0, r3 = (lane + 96 < ny) ? b[lane + 96] : 0.0;
    const int i0 = lane, i1 = lane + 32, i2 = lane + 64, i3 = lane + 96;
    const int o0 = i0 * (i0 + 1) / 2, o1 = i1 * (i1 + 1) / 2, o2 = i2 * (i2 + 1) / 2, o3 = i3 * (i3 + 1) / 2;
    // forward: z_i = r_i - sum_{j<i} Lt(i,j) invp_j z_j
    for (int j = 0; j < ny; j++) {
        const int src = j & 31, t = j >> 5;
        const double zj = __shfl_sync(0xffffffffu, t == 0 ? r0 : (t == 1 ? r1 : (t == 2 ? r2 : r3)), src);
        const double q = zj * invp[j];
        if (i0 > j && i0 < ny) r0 -= W[o0 + j] * q;
        if (i1 > j && i1 < ny) r1 -= W[o1 + j] * q;
        if (i2 > j && i2 < ny) r2 -= W[o2 + j] * q;
        if (i3 > j && i3 < ny) r3 -= W[o3 + j] * q;
    }
    // backward: w_i = invp_i (z_i - sum_{k>i} Lt(k,i) w_k)
    for (int i = ny - 1; i >= 0; i--) {
        const int src = i & 31, t = i >> 5;
        const double ai = __shfl_sync(0xffffffffu, t == 0 ? r0 : (t == 1 ? r1 : (t == 2 ? r2 : r3)), src);
        const double wi = ai * invp[i];
        const int row = i * (i + 1) / 2;
        if (i0 < i) r0 -= W[row + i0] * wi;
        if (i1 < i) r1 -= W[row + i1] * wi;
        if (i2 < i) r2 -= W[row + i2] * wi;
        if (i3 < i) r3 -= W[row + i3] * wi;
    }
    if (i0 < ny) b[i0] = r0 * invp[i0];
    if (i1 < ny) b[i1] = r1 * invp[i1];
    if (i2 < ny) b[i2] = r2 * invp[i2];
    if (i3 < ny) b[i3] = r3 * invp[i3];
#else
    for (int j = 0; j < ny; j++) {
        const double q = b[j] * invp[j];
        for (int i = j + 1; i < ny; i++) b[i] -= W[i * (i + 1) / 2 + j] * q;
    }
    for (int i = ny - 1; i >= 0; i--) {
        const double wi = b[i] * invp[i];
        for (int k = 0; k < i; k++) b[k] -= W[i * (i + 1) / 2 + k] * wi;
    }
    for (int i = 0; i < ny; i++) b[i] *= invp[i];
#endif
}

// ------------------------------------------------------------------------------------------------
// the solver
// ------------------------------------------------------------------------------------------------
DLSC_HD int sym_idx(int k, int kk) { return k >= kk ? k * (k + 1) / 2 + kk : kk * (kk + 1) / 2 + k; }

// x = c + T y  (or dx = T dy when cst == nullptr); T hard-wired from the elimination in dlsc_qp_tables.h:
//   x[m][3+j] = y[3m+j] (m < M-1), x[M-1][3..5] = y[3(M-1)], x[m][0] = y5(m-1), x[m][1] = 2 y5 - y4,
//   x[m][2] = 4 y5 - 4 y4 + y3 (m >= 1), x[0][0..2] = c0, c1, c2
DLSC_HD void map_x(const Cta& c, const QpTab& T, const double* y, const double* cst, double* x) {
    const int M = T.M, nyd = T.nyd, npt = T.npt;
    for (int k = 0; k < T.D; k++) {
        const double* yk = y + k * nyd;
        for (int pt = c.tid; pt < npt; pt += c.nthr) {
            const int m = pt / kP, i = pt - m * kP;
            double v;
            if (i >= 3) v = (m == M - 1) ? yk[3 * (M - 1)] : yk[3 * m + i - 3];
            else if (m == 0) v = cst ? cst[k * 3 + i] : 0.0;
            else {
                const double* q = yk + 3 * (m - 1);
                v = (i == 0) ? q[2] : (i == 1 ? 2.0 * q[2] - q[1] : 4.0 * q[2] - 4.0 * q[1] + q[0]);
            }
            x[k * npt + pt] = v;
        }
    }
}

// uy[p] = (T' ax)[p] for global unknown p
DLSC_HD double gather_y(const QpTab& T, int p, const double* ax) {
    const int M = T.M, nyd = T.nyd, npt = T.npt;
    const int k = p / nyd, a = p - k * nyd;
    const double* xk = ax + k * npt;
    if (a == 3 * (M - 1)) { const double* q = xk + (M - 1) * kP; return q[3] + q[4] + q[5]; }
    const int m = a / 3, j = a - 3 * m;
    const double* q = xk + m * kP;          // q[6..8] = points 0..2 of segment m+1
    if (j == 0) return q[3] + q[8];
    if (j == 1) return q[4] - q[7] - 4.0 * q[8];
    return q[5] + q[6] + 2.0 * q[7] + 4.0 * q[8];
}

struct QpConst {
    double hi_v, hi_a, hi_c, wpr;
    int ts;
};

// ---- two-sided rows with constant patterns, evaluated by stencil arithmetic in x-space ----
// row index inside an axis (see dlsc_qp_tables.cpp): box | velocity | acceleration | comm
struct PairRow { int fam, pa, pb; };
DLSC_HD PairRow pair_decode(const QpTab& T, int idx) {
    PairRow r; r.pb = 0;
    if (idx < T.row_bv) { r.fam = 0; r.pa = idx + 3; }
    else if (idx < T.row_ba) { const int t = idx - T.row_bv + 2; r.fam = 1; r.pa = (t / 5) * kP + t % 5; }
    else if (idx < T.row_bc) { const int t = idx - T.row_ba + 1; r.fam = 2; r.pa = (t / 4) * kP + t % 4; }
    else {
        int t = idx - T.row_bc, mi = 0;
        while (t >= T.M - mi) { t -= T.M - mi; mi++; }
        r.fam = 3; r.pa = (mi + t) * kP + (kP - 1); r.pb = mi * kP;
    }
    return r;
}
DLSC_HD double pair_eval(const QpTab& T, const PairRow& r, const double* xk) {
    if (r.fam == 0) return xk[r.pa];
    if (r.fam == 1) return T.scv * (xk[r.pa + 1] - xk[r.pa]);
    if (r.fam == 2) return T.sca * (xk[r.pa + 2] - 2.0 * xk[r.pa + 1] + xk[r.pa]);
    return xk[r.pa] - xk[r.pb];
}
// (G_pair' V) at point pt of one axis; Vk = V + k*row_npl
DLSC_HD double pair_transpose(const QpTab& T, const double* Vk, int pt) {
    const int M = T.M, m = pt / kP, i = pt - m * kP;
    double v = 0.0;
    if (pt >= 3) v += Vk[pt - 3];
    const double* vv = Vk + T.row_bv + 5 * m - 2;       // velocity row (m, j) at vv[j], valid unless m == 0 && j < 2
    if (i >= 1 && (m > 0 || i - 1 >= 2)) v += T.scv * vv[i - 1];
    if (i <= 4 && (m > 0 || i >= 2)) v -= T.scv * vv[i];
    const double* va = Vk + T.row_ba + 4 * m - 1;       // acceleration row (m, j) at va[j], valid unless m == 0 && j == 0
    if (i >= 2 && (m > 0 || i - 2 >= 1)) v += T.sca * va[i - 2];
    if (i >= 1 && i <= 4 && (m > 0 || i - 1 >= 1)) v -= 2.0 * T.sca * va[i - 1];
    if (i <= 3 && (m > 0 || i >= 1)) v += T.sca * va[i];
    if (T.use_comm) {
        const double* vc = Vk + T.row_bc;               // comm row (mi, m') at vc[mi*M - mi(mi-1)/2 + m' - mi]
        if (i == kP - 1) for (int mi = 0; mi <= m; mi++) v += vc[mi * M - mi * (mi - 1) / 2 + m - mi];
        if (i == 0) for (int m2 = m; m2 < M; m2++) v -= vc[m * M - m * (m - 1) / 2 + m2 - m];
    }
    return v;
}

DLSC_HD void pair_bounds(const DevParams& P, const QpTab& T, const QpIn& in, const QpConst& qc, int k, const PairRow& r,
                         double& lo, double& hi) {
    if (r.fam == 0) {
        const int m = r.pa / kP, i = r.pa - m * kP;
        hi = P.world_max[k]; lo = P.world_min[k];
        if (P.use_sfc) {
            const double bl = (double)in.sfc[m * 6 + k], bh = (double)in.sfc[m * 6 + 3 + k];
            if (bl > lo) lo = bl;
            if (bh < hi) hi = bh;
        }
        if (T.use_comm && i == kP - 1) {
            const double w = (double)v3_get(in.wp, k);
            if (w - qc.wpr > lo) lo = w - qc.wpr;
            if (w + qc.wpr < hi) hi = w + qc.wpr;
        }
    } else if (r.fam == 1) { hi = qc.hi_v; lo = -qc.hi_v; }
    else if (r.fam == 2) { hi = qc.hi_a; lo = -qc.hi_a; }
    else { hi = qc.hi_c; lo = -qc.hi_c; }
}

// LSC row (pt, cc) of this agent:  -n.x <= b  with  b = -(n.anchor + d)   (traj_optimizer.cpp:412-450)
struct LscRowData { double n0, n1, n2, b; };
DLSC_HD LscRowData lsc_row_data(const DevParams& P, const QpIn& in, int pt, int cc, int n_dyn) {
    const int m = pt / kP, i = pt - m * kP;
    const float* nr = in.normal + ((size_t)cc * P.M + m) * 3;
    // agents: witness point of the segment case for the last segment; dynamic obstacles (slots < n_dyn): always the
    // predicted control point (traj_planner.cpp:625)
    const float* an = (m < P.M - 1 || cc < n_dyn) ? in.pred_traj + ((size_t)in.nbr_idx[cc] * (P.M * kP) + pt) * 3
                                                    : in.anchor_last + cc * 3;
    LscRowData r;
    r.n0 = (double)nr[0]; r.n1 = (double)nr[1]; r.n2 = (P.D == 3) ? (double)nr[2] : 0.0;
    double b = -in.d[((size_t)cc * P.M + m) * kP + i];
    b -= r.n0 * (double)an[0];
    b -= r.n1 * (double)an[1];
    if (P.D == 3) b -= r.n2 * (double)an[2];
    r.b = b;
    return r;
}

// ------------------------------------------------------------------------------------------------
// Dual active-set solver (Goldfarb & Idnani 1983, range-space / Schur-complement form).
//   min 1/2 y'Hy + g'y  s.t.  a_i'y <= b_i.   H = blockdiag_k(H1 + terminal) is constant per terminal-segment
//   count, so H^-1 (one nyd x nyd block) is tabulated; starting from y0 = -H^-1 g (all rows inactive) the
//   most violated row p is driven to its bound along  z = H^-1 (a_p - A' r),  r = S^-1 A H^-1 a_p,
//   S = A H^-1 A' over the active rows A (q <= 32; Cholesky of S updated by one appended row per added
//   constraint, refactored on the rare drop).  The replan QPs have ~1 active row at the optimum (median 0),
//   so this needs a handful of O(n*nyd) steps instead of interior-point iterations with an n^3 factorisation.
// Returns 0 = optimal, kStQpMaxIter = infeasible, -1 = give up (caller falls back to the interior point).
// y in sm.y, x in sm.x on return.  Row cache: l_n0/l_n1/l_n2/l_b [npt*Kc] (b = +inf for unused rows).
// ------------------------------------------------------------------------------------------------
DLSC_HD int qp_dual_active_set(const Cta& c, const DevParams& P, const QpTab& T, const QpIn& in, const QpConst& qc,
                               const QpSmem& sm, const double* __restrict__ l_n0, const double* __restrict__ l_n1,
                               const double* __restrict__ l_n2, const double* __restrict__ l_b,
                               const double* __restrict__ pb_hi, const double* __restrict__ pb_lo, int* iters_out,
                               double* viol_out) {
    const int M = P.M, D = P.D, ny = T.ny, nyd = T.nyd, npt = T.npt, np = T.np, npl = T.row_npl, Kc = P.K, K = in.K;
    const bool D3 = (D == 3);
    GiSmem g;
    gi_carve(T, sm.W, g);
    const double tol = 1e-9;
    // H^-1 block of this agent; g = T' grad f(x(0)) was left in sm.rd by the caller
    {
        const double* src = T.Hinv + (size_t)(qc.ts - 1) * nyd * nyd;
        for (int e = c.tid; e < nyd * nyd; e += c.nthr) g.Hinv[e] = src[e];
    }
    c.sync();
    for (int p = c.tid; p < ny; p += c.nthr) {                 // y0 = -H^-1 g
        const int k = p / nyd, a = p - k * nyd;
        const double* hr = g.Hinv + a * nyd;
        const double* gk = sm.rd + k * nyd;
        double v = 0.0;
        for (int b = 0; b < nyd; b++) v += hr[b] * gk[b];
        sm.y[p] = -v;
    }
    int q = 0, iters = 0, status = -1;
    double viol_p = 0.0, u_p = 0.0;
    bool same_p = false;
    c.sync();
    for (int guard = 0; guard < 400; guard++) {
        if (!same_p) {
            // ---- most violated row ----
            map_x(c, T, sm.y, sm.cst, sm.x);
            c.sync();
            double best = -1e300, best_id = 1e300;
            for (int k = 0; k < D; k++)
                for (int idx = c.tid; idx < npl; idx += c.nthr) {
                    const int r = k * npl + idx;
                    const double act = pair_eval(T, pair_decode(T, idx), sm.x + k * npt);
                    const double vh = act - pb_hi[r], vl = pb_lo[r] - act;
                    if (vh > best) { best = vh; best_id = 2.0 * r; }
                    if (vl > best) { best = vl; best_id = 2.0 * r + 1.0; }
                }
            for (int pt = 3; pt < npt; pt++) {
                const double x0 = sm.x[pt], x1 = sm.x[npt + pt], x2 = D3 ? sm.x[2 * npt + pt] : 0.0;
                for (int cc = c.tid; cc < K; cc += c.nthr) {
                    const int o = pt * Kc + cc;
                    const double v = -(l_n0[o] * x0 + l_n1[o] * x1 + l_n2[o] * x2) - l_b[o];
                    if (v > best) { best = v; best_id = 2.0 * np + o; }
                }
            }
            double vmax = best, d0 = 0.0, d1 = 0.0;
            c.reduce3(vmax, 1, d0, 0, d1, 0);
            double idsel = (best == vmax) ? best_id : 1e300;
            d0 = 0.0; d1 = 0.0;
            c.reduce3(idsel, 2, d0, 0, d1, 0);
            *viol_out = vmax;
            if (!(vmax > tol)) { status = 0; break; }
            // ---- candidate row p: x-space form -> y-space form (thread 0) ----
            if (c.tid == 0) {
                const int id = (int)idsel;
                int xi[3] = {0, 0, 0}; double xc[3] = {0, 0, 0}; int nxe = 0; double bp;
                if (id < 2 * np) {
                    const int r = id >> 1, side = id & 1, k = r / npl, idx = r - k * npl;
                    const PairRow pr = pair_decode(T, idx);
                    const double sg = side ? -1.0 : 1.0;
                    const int base = k * npt;
                    if (pr.fam == 0) { xi[0] = base + pr.pa; xc[0] = sg; nxe = 1; }
                    else if (pr.fam == 1) { xi[0] = base + pr.pa + 1; xc[0] = sg * T.scv; xi[1] = base + pr.pa; xc[1] = -sg * T.scv; nxe = 2; }
                    else if (pr.fam == 2) {
                        xi[0] = base + pr.pa + 2; xc[0] = sg * T.sca; xi[1] = base + pr.pa + 1; xc[1] = -2.0 * sg * T.sca;
                        xi[2] = base + pr.pa; xc[2] = sg * T.sca; nxe = 3;
                    } else { xi[0] = base + pr.pa; xc[0] = sg; xi[1] = base + pr.pb; xc[1] = -sg; nxe = 2; }
                    bp = side ? -pb_lo[r] : pb_hi[r];
                } else {
                    const int o = id - 2 * np, pt = o / Kc;
                    xi[0] = pt; xc[0] = -l_n0[o]; xi[1] = npt + pt; xc[1] = -l_n1[o]; nxe = 2;
                    if (D3) { xi[2] = 2 * npt + pt; xc[2] = -l_n2[o]; nxe = 3; }
                    bp = l_b[o];
                }
                double* yc = g.yc + 9 * kGiQ; int16_t* yi = g.yi + 9 * kGiQ;
                int nt = 0;
                for (int e = 0; e < nxe; e++) {
                    const int k = xi[e] / npt, pt = xi[e] - k * npt, m = pt / kP, i = pt - m * kP, yb = k * nyd;
                    const double cf = xc[e];
                    if (i >= 3) { yi[nt] = (int16_t)(yb + ((m == M - 1) ? 3 * (M - 1) : 3 * m + i - 3)); yc[nt++] = cf; }
                    else if (m > 0) {
                        const int q3 = yb + 3 * (m - 1);
                        if (i == 0) { yi[nt] = (int16_t)(q3 + 2); yc[nt++] = cf; }
                        else if (i == 1) { yi[nt] = (int16_t)(q3 + 2); yc[nt++] = 2.0 * cf; yi[nt] = (int16_t)(q3 + 1); yc[nt++] = -cf; }
                        else { yi[nt] = (int16_t)(q3 + 2); yc[nt++] = 4.0 * cf; yi[nt] = (int16_t)(q3 + 1); yc[nt++] = -4.0 * cf;
                               yi[nt] = (int16_t)q3; yc[nt++] = cf; }
                    }
                }
                for (; nt < 9; nt++) { yi[nt] = -1; yc[nt] = 0.0; }
                g.bq[kGiQ] = bp; g.id[kGiQ] = id;
            }
            viol_p = vmax; u_p = 0.0;
            c.sync();
            // ---- w = H^-1 a_p (into sm.dy) ----
            for (int p = c.tid; p < ny; p += c.nthr) {
                const int k = p / nyd, a = p - k * nyd;
                const double* hr = g.Hinv + a * nyd;
                const double* yc = g.yc + 9 * kGiQ; const int16_t* yi = g.yi + 9 * kGiQ;
                double v = 0.0;
#pragma unroll
                for (int t = 0; t < 9; t++) {
                    const int j = yi[t];
                    if (j >= 0 && j / nyd == k) v += yc[t] * hr[j - k * nyd];
                }
                sm.dy[p] = v;
            }
            c.sync();
        }
        iters++;
        // ---- small dense step (thread 0): r, step lengths, active-set update; t_y = A' r into sm.ax1[0..ny) ----
        if (c.tid == 0) {
            const double* yc = g.yc + 9 * kGiQ; const int16_t* yi = g.yi + 9 * kGiQ;
            double apw = 0.0;
            for (int t = 0; t < 9; t++) if (yi[t] >= 0) apw += yc[t] * sm.dy[yi[t]];
            double ll = 0.0;
            for (int j = 0; j < q; j++) {
                double vj = 0.0;
                for (int t = 0; t < 9; t++) { const int jj = g.yi[9 * j + t]; if (jj >= 0) vj += g.yc[9 * j + t] * sm.dy[jj]; }
                g.v[j] = vj;
                double a = vj;
                for (int k = 0; k < j; k++) a -= g.Ls[j * (j + 1) / 2 + k] * g.l[k];
                a /= g.Ls[j * (j + 1) / 2 + j];
                g.l[j] = a; ll += a * a;
            }
            for (int j = q - 1; j >= 0; j--) {
                double a = g.l[j];
                for (int k = j + 1; k < q; k++) a -= g.Ls[k * (k + 1) / 2 + j] * g.r[k];
                g.r[j] = a / g.Ls[j * (j + 1) / 2 + j];
            }
            const double zn = apw - ll;
            double t1 = 1e300; int kdrop = -1;
            for (int j = 0; j < q; j++)
                if (g.r[j] > 0) { const double tt = g.u[j] / g.r[j]; if (tt < t1) { t1 = tt; kdrop = j; } }
            const bool dependent = !(zn > 1e-12 * (apw > 1e-300 ? apw : 1e-300));
            const double t2 = dependent ? 1e300 : viol_p / zn;
            const double t = t1 < t2 ? t1 : t2;
            double flag;             // 0: full step, new scan | 1: partial, same p | 2: infeasible | 3: give up
            if (!(t < 1e299)) flag = 2.0;
            else {
                for (int j = 0; j < q; j++) g.u[j] -= t * g.r[j];
                u_p += t;
                for (int e = 0; e < ny; e++) sm.ax1[e] = 0.0;
                for (int j = 0; j < q; j++)
                    for (int tt = 0; tt < 9; tt++) { const int jj = g.yi[9 * j + tt]; if (jj >= 0) sm.ax1[jj] += g.r[j] * g.yc[9 * j + tt]; }
                if (t2 <= t1) {
                    // full step: p becomes active
                    if (q == kGiQ) flag = 3.0;
                    else {
                        for (int k = 0; k < q; k++) { g.Ls[q * (q + 1) / 2 + k] = g.l[k]; g.Sm[q * (q + 1) / 2 + k] = g.v[k]; }
                        g.Ls[q * (q + 1) / 2 + q] = sqrt(zn); g.Sm[q * (q + 1) / 2 + q] = apw;
                        for (int tt = 0; tt < 9; tt++) { g.yc[9 * q + tt] = yc[tt]; g.yi[9 * q + tt] = yi[tt]; }
                        g.bq[q] = g.bq[kGiQ]; g.id[q] = g.id[kGiQ]; g.u[q] = u_p;
                        flag = 0.0;
                    }
                } else {
                    // partial step: row kdrop leaves the active set; S loses a row/column, refactor
                    for (int j = kdrop; j < q - 1; j++) {
                        for (int tt = 0; tt < 9; tt++) { g.yc[9 * j + tt] = g.yc[9 * (j + 1) + tt]; g.yi[9 * j + tt] = g.yi[9 * (j + 1) + tt]; }
                        g.bq[j] = g.bq[j + 1]; g.id[j] = g.id[j + 1]; g.u[j] = g.u[j + 1];
                    }
                    for (int i2 = 0, ii = 0; i2 < q; i2++) {
                        if (i2 == kdrop) continue;
                        for (int k2 = 0, kk = 0; k2 <= i2; k2++) {
                            if (k2 == kdrop) continue;
                            g.Ls[ii * (ii + 1) / 2 + kk] = g.Sm[i2 * (i2 + 1) / 2 + k2];     // Ls as temp for the shrunk S
                            kk++;
                        }
                        ii++;
                    }
                    for (int e = 0; e < (q - 1) * q / 2; e++) g.Sm[e] = g.Ls[e];
                    bool okf = true;
                    for (int j = 0; j < q - 1 && okf; j++) {
                        double dj = g.Sm[j * (j + 1) / 2 + j];
                        for (int k = 0; k < j; k++) dj -= g.Ls[j * (j + 1) / 2 + k] * g.Ls[j * (j + 1) / 2 + k];
                        if (!(dj > 0)) { okf = false; break; }
                        dj = sqrt(dj);
                        g.Ls[j * (j + 1) / 2 + j] = dj;
                        for (int i2 = j + 1; i2 < q - 1; i2++) {
                            double a = g.Sm[i2 * (i2 + 1) / 2 + j];
                            for (int k = 0; k < j; k++) a -= g.Ls[i2 * (i2 + 1) / 2 + k] * g.Ls[j * (j + 1) / 2 + k];
                            g.Ls[i2 * (i2 + 1) / 2 + j] = a / dj;
                        }
                    }
                    flag = okf ? 1.0 : 3.0;
                }
            }
            g.ty[0] = dependent ? 0.0 : t;      // primal step length
            g.ty[1] = flag;
            g.ty[2] = zn;
        }
        c.sync();
        const double tp = g.ty[0], flag = g.ty[1];
        if (flag == 2.0) { status = kStQpMaxIter; break; }
        if (flag == 3.0) { status = -1; break; }
        // ---- y -= t (w - H^-1 A' r) ----
        if (tp != 0.0)
            for (int p = c.tid; p < ny; p += c.nthr) {
                const int k = p / nyd, a = p - k * nyd;
                const double* hr = g.Hinv + a * nyd;
                const double* tk = sm.ax1 + k * nyd;
                double hz = 0.0;
                if (q > 0 || flag == 1.0)
                    for (int b = 0; b < nyd; b++) hz += hr[b] * tk[b];
                sm.y[p] -= tp * (sm.dy[p] - hz);
            }
        if (flag == 0.0) { q++; same_p = false; }
        else { q--; same_p = true; viol_p -= tp * g.ty[2]; }
        c.sync();
    }
    map_x(c, T, sm.y, sm.cst, sm.x);
    c.sync();
    *iters_out = iters;
    return status;
}

// one agent
// skip_gi: the dual active set already ran on this agent (dlsc_qp_gi.cuh) and gave up
// DYN: dynamic obstacles present (slots < P.n_dyn): their LSC rows carry the slack variable of their (obstacle, segment)
// group, eps <= 0 with cost w (M-m)/M eps^2.  The slack block is eliminated from every Newton system: with
// u_g = sum_{r in g} D_r a_r and W_ee[g] = h_g + sum D_r + D_bound the reduced matrix is W - sum_g u_g u_g' / W_ee[g]
// (u_g lives on <= 18 unknowns), the slack step follows from dy.  The DYN = false instantiation contains none of it.
template <bool DYN = false>
DLSC_HD void qp_agent(const Cta& c, const DevParams& P, const QpTab& T, const QpIn& in, const QpOut& out,
                      const QpSmem& sm, double* scratch, bool skip_gi = false) {
    const int M = P.M, D = P.D, ny = T.ny, nyd = T.nyd, npt = T.npt, nx = T.nx, np = T.np, Kc = P.K;
    const int K = in.K;
    const int n = kP - 1;
    const bool D3 = (D == 3);
    const size_t LS = (size_t)npt * Kc;
    double* __restrict__ l_n0 = scratch; double* __restrict__ l_n1 = scratch + LS; double* __restrict__ l_n2 = scratch + 2 * LS;
    double* __restrict__ l_b = scratch + 3 * LS; double* __restrict__ l_s = scratch + 4 * LS; double* __restrict__ l_z = scratch + 5 * LS;
    double* __restrict__ l_c = scratch + 6 * LS; double* __restrict__ l_ds = scratch + 7 * LS; double* __restrict__ l_dz = scratch + 8 * LS;
    int* __restrict__ l_pt = reinterpret_cast<int*>(scratch + 9 * LS);
    double* pr = scratch + 9 * LS + (LS + 1) / 2;
    double* __restrict__ ps_hi = pr; double* __restrict__ pz_hi = pr + np; double* __restrict__ pc_hi = pr + 2 * np;
    double* __restrict__ ps_lo = pr + 3 * np; double* __restrict__ pz_lo = pr + 4 * np; double* __restrict__ pc_lo = pr + 5 * np;
    double* __restrict__ pb_hi = pr + 6 * np; double* __restrict__ pb_lo = pr + 7 * np; double* __restrict__ pd_sh = pr + 8 * np;
    double* __restrict__ pd_zh = pr + 9 * np; double* __restrict__ pd_sl = pr + 10 * np; double* __restrict__ pd_zl = pr + 11 * np;
    double* __restrict__ p_act = pr + 12 * np;
    const int npl = T.row_npl;
    // dynamic-obstacle block (DYN only)
    const int nd = DYN ? P.n_dyn : 0, NS = nd * M;
    int* __restrict__ l_g = reinterpret_cast<int*>(pr + 13 * np);                 // slack group of a working-set row, or -1
    double* dq = pr + 13 * np + (LS + 1) / 2;
    const int NSm = kMaxDyn * M;
    double* __restrict__ q_e = dq; double* __restrict__ q_de = dq + NSm; double* __restrict__ q_sb = dq + 2 * NSm;
    double* __restrict__ q_zb = dq + 3 * NSm; double* __restrict__ q_dsb = dq + 4 * NSm; double* __restrict__ q_dzb = dq + 5 * NSm;
    double* __restrict__ q_ccb = dq + 6 * NSm; double* __restrict__ q_wee = dq + 7 * NSm; double* __restrict__ q_rhe = dq + 8 * NSm;
    double* __restrict__ q_rde = dq + 9 * NSm; double* __restrict__ q_pb = dq + 10 * NSm; double* __restrict__ q_u = dq + 11 * NSm;
    int* __restrict__ q_row = reinterpret_cast<int*>(dq + (11 + kDynSlots) * NSm);   // [NS][6] working-set row of the group at point i, or -1
    auto slack_h = [&](int g) { return 2.0 * P.slack_w * ((double)(M - g % M) / M); };
    // group quantities from the rows' current (D, c, z): W_ee, u, dual residual, right-hand side; cb = the bound row's c
    auto dyn_groups = [&](bool first) {
        for (int g = c.tid; g < NS; g += c.nthr) {
            const double Db = q_zb[g] / q_sb[g];
            double wee = slack_h(g) + Db, zsum = q_zb[g], ce = q_ccb[g];
            double u[kDynSlots];
#pragma unroll
            for (int t = 0; t < kDynSlots; t++) u[t] = 0.0;
            for (int i = 0; i < kP; i++) {
                const int r = q_row[g * kP + i];
                if (r < 0) continue;
                const double dd = l_dz[r];
                zsum += l_z[r]; ce += l_ds[r];
                if (!first) continue;
                wee += dd;
                const double w3[3] = {-l_n0[r] * dd, -l_n1[r] * dd, -l_n2[r] * dd};
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    double* uk = u + k * 6;
                    const double w = w3[k];
                    if (i == 0) uk[2] += w;
                    else if (i == 1) { uk[2] += 2.0 * w; uk[1] -= w; }
                    else if (i == 2) { uk[2] += 4.0 * w; uk[1] -= 4.0 * w; uk[0] += w; }
                    else if (g % M == M - 1) uk[3] += w;
                    else uk[i] += w;
                }
            }
            if (first) {
                // stored scaled: u / sqrt(W_ee), rhs / sqrt(W_ee) -- the three places that use them need no division
                const double isw = 1.0 / sqrt(wee);
                q_wee[g] = isw;
                for (int t = 0; t < kDynSlots; t++) q_u[g * kDynSlots + t] = u[t] * isw;
                q_rde[g] = slack_h(g) * q_e[g] + zsum;
            }
            q_rhe[g] = (-q_rde[g] - ce) * q_wee[g];
        }
    };
    // right-hand side of the reduced system: dy[p] -= sum_g u_g[p] rhe[g] / wee[g] (every unknown belongs to <= 2 segments)
    auto dyn_reduce_rhs = [&]() {
        for (int p = c.tid; p < ny; p += c.nthr) {
            const int k = p / nyd, a = p - k * nyd;
            double acc = 0.0;
            for (int which = 0; which < 2; which++) {
                const int m = which ? a / 3 + 1 : a / 3, slot = which ? a % 3 : 3 + a % 3;
                if (m > M - 1 || (which == 0 && m == M - 1 && slot != 3)) continue;
                for (int o = 0; o < nd; o++) {
                    const int g = o * M + m;
                    acc += q_u[g * kDynSlots + k * 6 + slot] * q_rhe[g];
                }
            }
            sm.dy[p] -= acc;
        }
    };
    // slack step from dy
    auto dyn_back = [&]() {
        for (int g = c.tid; g < NS; g += c.nthr) {
            const int m = g % M;
            double acc = q_rhe[g];
            for (int k = 0; k < D; k++)
                for (int slot = 0; slot < 6; slot++) {
                    const int yi = dyn_slot_y(T, m, k, slot);
                    if (yi >= 0) acc -= q_u[g * kDynSlots + k * 6 + slot] * sm.dy[yi];
                }
            q_de[g] = acc * q_wee[g];
        }
    };
    // ---- constants of this agent ----
    QpConst qc;
    qc.hi_v = in.max_vel; qc.hi_a = in.max_acc; qc.hi_c = 0.5 * P.comm_range - in.radius;
    qc.wpr = 0.5 * P.comm_range - kEpsF;
    {
        const double ideal = v3_norm(in.goal - in.pos) / in.nominal_vel;              // traj_optimizer.cpp:543-551
        int ts = (int)((M * P.dt - ideal + kEps) / P.dt);
        if (ts < 1) ts = 1;
        if (ts > M) ts = M;
        qc.ts = ts;
    }
    for (int k = c.tid; k < D; k += c.nthr) {                                         // :335-352
        const double c0 = (double)v3_get(in.pos, k);
        const double c1 = c0 + (double)v3_get(in.vel, k) * P.dt / n;
        const double c2 = (double)v3_get(in.acc, k) * P.dt * P.dt / (n * (n - 1)) + 2 * c1 - c0;
        sm.cst[k * 3] = c0; sm.cst[k * 3 + 1] = c1; sm.cst[k * 3 + 2] = c2;
    }
    // usable LSC (m, c) pairs: neighbour present and normal not ~0 (:422-424)
    for (int e = c.tid; e < M * Kc; e += c.nthr) {
        const int m = e / Kc, cc = e - m * Kc;
        uint8_t a = 0;
        if (cc < K) a = !(v3_norm(v3_load(in.normal + ((size_t)cc * M + m) * 3)) < kEpsF);
        sm.act[e] = a;
    }
    for (int p = c.tid; p < ny; p += c.nthr) sm.y[p] = 0.0;
    c.sync();

    const double wT = P.w_terminal, wT2 = 2.0 * P.w_terminal;
    auto grad_x = [&](const double* x, int k, int pt) -> double {   // d/dx of  w_u x'Qx + w_T (x_n - goal)^2
        const int m = pt / kP, i = pt - m * kP;
        const double* xs = x + k * npt + m * kP;
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < kP; j++) v += T.Q2[i * kP + j] * xs[j];
        if (i == n && m >= M - qc.ts) v += wT2 * (xs[n] - (double)v3_get(in.goal, k));
        return v;
    };
    // ---- g_inf: gradient of the objective at y = 0 ----
    map_x(c, T, sm.y, sm.cst, sm.x);
    c.sync();
    for (int k = 0; k < D; k++)
        for (int pt = c.tid; pt < npt; pt += c.nthr) sm.ax1[k * npt + pt] = grad_x(sm.x, k, pt);
    c.sync();
    double g_inf = 0.0;
    for (int p = c.tid; p < ny; p += c.nthr) {
        const double gp = gather_y(T, p, sm.ax1);
        sm.rd[p] = gp;                              // linear term g of 1/2 y'Hy + g'y
        if (fabs(gp) > g_inf) g_inf = fabs(gp);
    }
    { double d0 = 0.0, d1 = 0.0; c.reduce3(g_inf, 1, d0, 0, d1, 0); }
    // pair-row bounds
    for (int k = 0; k < D; k++)
        for (int idx = c.tid; idx < npl; idx += c.nthr) {
            double lo, hi;
            pair_bounds(P, T, in, qc, k, pair_decode(T, idx), lo, hi);
            pb_hi[k * npl + idx] = hi; pb_lo[k * npl + idx] = lo;
        }

    int status = kStQpMaxIter;
    int it_total = 0;
    double viol_lsc = 0.0;
    long long rows_total = 0;
    // agents pushed around by a dynamic obstacle end far from their initial trajectory: a wider first working set saves
    // the second attempt (measured on the 4096-agent forest with 8 obstacles: x1 needs retries, 7.4 ms; x2 4.5 ms; x4 5.0 ms)
    double tau = DYN ? DLSC_DYN_TAU_MULT * P.qp_screen : P.qp_screen;
    const int max_it = P.qp_max_iter;

    // ---- primary solver: dual active set on all rows (LSC rows cached as (n, b) per (point, neighbour)) ----
    bool solved = false, solved_by_gi = false;
    if (P.qp_solver != 1 && !skip_gi) {
        for (int pt = 3; pt < npt; pt++) {
            const int m = pt / kP;
            for (int cc = c.tid; cc < K; cc += c.nthr) {
                const int o = pt * Kc + cc;
                if (sm.act[m * Kc + cc]) {
                    const LscRowData r = lsc_row_data(P, in, pt, cc, 0);
                    l_n0[o] = r.n0; l_n1[o] = r.n1; l_n2[o] = r.n2; l_b[o] = r.b;
                } else { l_n0[o] = 0.0; l_n1[o] = 0.0; l_n2[o] = 0.0; l_b[o] = 1e300; }
            }
        }
        c.sync();
        int gi_it = 0;
        double gi_viol = 0.0;
        const int st = qp_dual_active_set(c, P, T, in, qc, sm, l_n0, l_n1, l_n2, l_b, pb_hi, pb_lo, &gi_it, &gi_viol);
        it_total += gi_it;
        rows_total += 2LL * np + (long long)K * (npt - 3);
        if (st >= 0) { status = st; solved = true; solved_by_gi = true; viol_lsc = gi_viol > 0 ? gi_viol : 0.0; }
        c.sync();
    }
    // ---- fallback: interior point on the screened working set ----
    for (int attempt = 0; attempt < 4 && !solved; attempt++) {
        const bool all_rows = !(tau > 0) || attempt == 3;
        // ---- starting point: free control points of the initial trajectory ----
        for (int p = c.tid; p < ny; p += c.nthr) {
            const int k = p / nyd, a = p - k * nyd;
            const int m = a / 3, j = (m == M - 1) ? 2 : a - 3 * m;
            sm.y[p] = (double)in.init_traj[(m * kP + 3 + j) * 3 + k];
        }
        c.sync();
        map_x(c, T, sm.y, sm.cst, sm.x);
        c.sync();
        // ---- working set of LSC rows, sorted by control point ----
        for (int pt = c.tid; pt < npt; pt += c.nthr) {
            int cnt = 0;
            if (pt >= 3) {
                const int m = pt / kP;
                const double x0 = sm.x[pt], x1 = sm.x[npt + pt], x2 = D3 ? sm.x[2 * npt + pt] : 0.0;
                for (int cc = 0; cc < K; cc++) {
                    if (!sm.act[m * Kc + cc]) continue;
                    const LscRowData r = lsc_row_data(P, in, pt, cc, nd);
                    const double s0 = r.b + (r.n0 * x0 + r.n1 * x1 + r.n2 * x2);
                    if (all_rows || s0 < tau || cc < nd) cnt++;                     // dynamic-obstacle rows: always in the working set
                }
            }
            sm.off[pt + 1] = cnt;
        }
        c.sync();
        if (c.tid == 0) {
            sm.off[0] = 0;
            for (int pt = 0; pt < npt; pt++) sm.off[pt + 1] += sm.off[pt];
        }
        c.sync();
        const int nl = sm.off[npt];
        if (DYN) {                                  // slack variables start at 0, their bound rows like every other row
            for (int g = c.tid; g < NS; g += c.nthr) { q_e[g] = 0.0; q_sb[g] = 1e-2; q_zb[g] = 1.0; }
            for (int e = c.tid; e < NS * kP; e += c.nthr) q_row[e] = -1;
            c.sync();
        }
        for (int pt = c.tid; pt < npt; pt += c.nthr) {
            if (pt < 3 || sm.off[pt + 1] == sm.off[pt]) continue;
            const int m = pt / kP;
            int o = sm.off[pt];
            const double x0 = sm.x[pt], x1 = sm.x[npt + pt], x2 = D3 ? sm.x[2 * npt + pt] : 0.0;
            for (int cc = 0; cc < K; cc++) {
                if (!sm.act[m * Kc + cc]) continue;
                const LscRowData r = lsc_row_data(P, in, pt, cc, nd);
                const double act = -(r.n0 * x0 + r.n1 * x1 + r.n2 * x2);
                double s0 = r.b - act;
                if (!(all_rows || s0 < tau || cc < nd)) continue;
                if (s0 < 1e-2) s0 = 1e-2;
                l_n0[o] = r.n0; l_n1[o] = r.n1; l_n2[o] = r.n2; l_b[o] = r.b; l_s[o] = s0; l_z[o] = 1.0 / s0 * 1e-2;
                l_pt[o] = pt;
                if (DYN) {
                    const int g = cc < nd ? cc * M + m : -1;
                    l_g[o] = g;
                    if (g >= 0) q_row[g * kP + (pt - m * kP)] = o;
                }
                o++;
            }
        }
        // ---- pair rows: slack / dual initialisation ----
        for (int k = 0; k < D; k++)
        for (int idx = c.tid; idx < npl; idx += c.nthr) {
            const int r = k * npl + idx;
            const double act = pair_eval(T, pair_decode(T, idx), sm.x + k * npt);
            double s = pb_hi[r] - act; if (s < 1e-2) s = 1e-2;
            ps_hi[r] = s; pz_hi[r] = 1.0 / s * 1e-2;
            s = act - pb_lo[r]; if (s < 1e-2) s = 1e-2;
            ps_lo[r] = s; pz_lo[r] = 1.0 / s * 1e-2;
        }
        const double n_rows = 2.0 * np + nl + NS;
        rows_total += (long long)n_rows;
        const double inv_rows = 1.0 / n_rows;
        c.sync();

        status = kStQpMaxIter;
        int it = 0;
        bool acceptable = false;
        for (it = 0; it < max_it; it++) {
            // ============ pass 1: residuals, G'z, affine rhs, D = z/s ============
            double rp_inf = 0.0, mu = 0.0;
            for (int k = 0; k < D; k++)
            for (int idx = c.tid; idx < npl; idx += c.nthr) {
                const int r = k * npl + idx;
                const double act = pair_eval(T, pair_decode(T, idx), sm.x + k * npt);
                p_act[r] = act;                 // row activity of this iterate, reused by the later passes
                const double sh = ps_hi[r], zh = pz_hi[r], sl = ps_lo[r], zl = pz_lo[r];
                const double rph = act + sh - pb_hi[r], rpl = -act + sl + pb_lo[r];
                rp_inf = fmax(rp_inf, fmax(fabs(rph), fabs(rpl)));
                mu += sh * zh + sl * zl;
                const double dh = zh / sh, dl = zl / sl;
                sm.V1[r] = zh - zl;
                sm.V2[r] = dh * (rph - sh) - dl * (rpl - sl);      // (-s z + z rp)/s = (z/s)(rp - s)
                sm.DD[r] = dh + dl;
            }
            for (int r = c.tid; r < nl; r += c.nthr) {
                const int pt = l_pt[r];
                const double n0 = l_n0[r], n1 = l_n1[r], n2 = l_n2[r], s = l_s[r], z = l_z[r];
                const double eg = (DYN && l_g[r] >= 0) ? q_e[l_g[r]] : 0.0;
                const double act = -(n0 * sm.x[pt] + n1 * sm.x[npt + pt] + (D3 ? n2 * sm.x[2 * npt + pt] : 0.0)) + eg;
                const double rp = act + s - l_b[r];
                rp_inf = fmax(rp_inf, fabs(rp));
                mu += s * z;
                const double dd = z / s;
                l_ds[r] = dd * (rp - s);      // c_aff (temporary)
                l_dz[r] = dd;                 // D      (temporary)
            }
            if (DYN)
                for (int g = c.tid; g < NS; g += c.nthr) {          // bound rows eps <= 0
                    const double rpb = q_e[g] + q_sb[g];
                    rp_inf = fmax(rp_inf, fabs(rpb));
                    mu += q_sb[g] * q_zb[g];
                    q_ccb[g] = q_zb[g] / q_sb[g] * (rpb - q_sb[g]);
                }
            c.sync();
            if (DYN) dyn_groups(true);
            for (int pt = c.tid; pt < npt; pt += c.nthr) {
                double u10 = 0, u11 = 0, u12 = 0, u20 = 0, u21 = 0, u22 = 0, S0 = 0, S1 = 0, S2 = 0, S3 = 0, S4 = 0, S5 = 0;
                for (int r = sm.off[pt]; r < sm.off[pt + 1]; r++) {
                    const double n0 = l_n0[r], n1 = l_n1[r], n2 = l_n2[r], z = l_z[r], cv = l_ds[r], dd = l_dz[r];
                    u10 -= n0 * z; u11 -= n1 * z; u12 -= n2 * z;
                    u20 -= n0 * cv; u21 -= n1 * cv; u22 -= n2 * cv;
                    S0 += dd * n0 * n0; S1 += dd * n1 * n0; S2 += dd * n1 * n1;
                    S3 += dd * n2 * n0; S4 += dd * n2 * n1; S5 += dd * n2 * n2;
                }
                // x-space accumulators of this point: objective gradient + pair rows + LSC rows
                sm.ax1[pt] = grad_x(sm.x, 0, pt) + pair_transpose(T, sm.V1, pt) + u10;
                sm.ax2[pt] = pair_transpose(T, sm.V2, pt) + u20;
                sm.ax1[npt + pt] = grad_x(sm.x, 1, pt) + pair_transpose(T, sm.V1 + npl, pt) + u11;
                sm.ax2[npt + pt] = pair_transpose(T, sm.V2 + npl, pt) + u21;
                if (D3) {
                    sm.ax1[2 * npt + pt] = grad_x(sm.x, 2, pt) + pair_transpose(T, sm.V1 + 2 * npl, pt) + u12;
                    sm.ax2[2 * npt + pt] = pair_transpose(T, sm.V2 + 2 * npl, pt) + u22;
                }
                double* Sp = sm.S + pt * 6;
                Sp[0] = S0; Sp[1] = S1; Sp[2] = S2; Sp[3] = S3; Sp[4] = S4; Sp[5] = S5;
            }
            c.sync();
            double rd_inf = 0.0;
            for (int p = c.tid; p < ny; p += c.nthr) {
                const double r1 = gather_y(T, p, sm.ax1);
                const double r2 = gather_y(T, p, sm.ax2);
                sm.rd[p] = r1;
                sm.dy[p] = -r1 - r2;
                rd_inf = fmax(rd_inf, fabs(r1));
            }
            if (DYN)
                for (int g = c.tid; g < NS; g += c.nthr) rd_inf = fmax(rd_inf, fabs(q_rde[g]));
            c.reduce3(rp_inf, 1, mu, 0, rd_inf, 1);
            mu *= inv_rows;
#if defined(DLSC_QP_TRACE) && !defined(__CUDA_ARCH__)
            printf("att %d it %d rp %.3e rd %.3e mu %.3e ginf %.3e rows %.0f\n", attempt, it, rp_inf, rd_inf, mu, g_inf, n_rows);
#endif
            if (rp_inf <= kQpTolRp && rd_inf <= kQpTolRd * (1.0 + g_inf) && mu <= kQpTolMu) { status = 0; break; }
            if (rp_inf <= 1e-10 && rd_inf <= 1e-9 * (1.0 + g_inf) && mu <= 1e-11) acceptable = true;   // as in the oracle

            // ============ W = H + G' D G  (packed lower triangle) ============
            for (int e = c.tid; e < T.ntri; e += c.nthr) sm.W[e] = 0.0;
            c.sync();
            for (int ei = c.tid; ei < T.nnzw; ei += c.nthr) {
#ifdef __CUDA_ARCH__
                const uint4 h = __ldg(T.nz_hdr + ei);
#else
                const uint4 h = T.nz_hdr[ei];
#endif
                double v = T.nz_h[ei];
                const int e = h.x & 0xffff, k = (h.x >> 16) & 3, kk = (h.x >> 18) & 3;
                if (((h.x >> 20) & 1) && (int)h.w >= M - qc.ts) v += wT2;
                const int wo = h.y & 0xffffff, wn = h.y >> 24, po = h.z & 0xffffff, pn = h.z >> 24;
                for (int t = 0; t < wn; t++) v += T.wi_coef[wo + t] * sm.DD[T.wi_row[wo + t]];
                const int si = sym_idx(k, kk);
                for (int t = 0; t < pn; t++) v += T.wp_coef[po + t] * sm.S[T.wp_pt[po + t] * 6 + si];
                sm.W[e] = v;
            }
            c.sync();
            if (DYN) {
                // W -= sum_g u_g u_g' / W_ee[g] (u stored scaled).  The groups of a segment share their 18 unknowns and
                // adjacent segments share 9 of them, so: even segments, barrier, odd segments -- every W entry has one
                // writer per pass; obstacles in a fixed order inside the thread, so the sums are deterministic.
                constexpr int kPairs = kDynSlots * (kDynSlots + 1) / 2;
                for (int par = 0; par < 2; par++) {
                    const int nseg = (M - par + 1) / 2;
                    for (int w = c.tid; w < nseg * kPairs; w += c.nthr) {
                        const int m = 2 * (w / kPairs) + par, pq = w % kPairs;
                        int a = 0;
                        while ((a + 1) * (a + 2) / 2 <= pq) a++;
                        const int b = pq - a * (a + 1) / 2;
                        if (a / 6 >= D) continue;
                        const int ya = dyn_slot_y(T, m, a / 6, a % 6), yb = dyn_slot_y(T, m, b / 6, b % 6);
                        if (ya < 0 || yb < 0) continue;
                        double acc = 0.0;
                        for (int o = 0; o < nd; o++) {
                            const int g = o * M + m;
                            acc += q_u[g * kDynSlots + a] * q_u[g * kDynSlots + b];
                        }
                        const int hi = ya > yb ? ya : yb, lo = ya > yb ? yb : ya;
                        sm.W[hi * (hi + 1) / 2 + lo] -= acc;
                    }
                    c.sync();
                }
                dyn_reduce_rhs();
                c.sync();
            }
            if (!ldl_factor(c, sm.W, sm.invp, sm.pan, ny, T.tri_p)) { status = acceptable ? 0 : kStQpNumeric; break; }

            // ============ predictor ============
            ldl_solve(c, sm.W, sm.invp, sm.dy, ny);
            c.sync();
            if (DYN) dyn_back();
            map_x(c, T, sm.dy, nullptr, sm.dx);
            c.sync();
            double a_aff = 1.0;
            for (int k = 0; k < D; k++)
            for (int idx = c.tid; idx < npl; idx += c.nthr) {
                const int r = k * npl + idx;
                const double act = p_act[r], gd = pair_eval(T, pair_decode(T, idx), sm.dx + k * npt);
                const double sh = ps_hi[r], zh = pz_hi[r], sl = ps_lo[r], zl = pz_lo[r];
                const double dsh = -(act + sh - pb_hi[r]) - gd, dsl = -(-act + sl + pb_lo[r]) + gd;
                const double dzh = -zh - (zh / sh) * dsh, dzl = -zl - (zl / sl) * dsl;   // (-s z - z ds)/s
                if (dsh < 0) a_aff = fmin(a_aff, -sh / dsh);
                if (dzh < 0) a_aff = fmin(a_aff, -zh / dzh);
                if (dsl < 0) a_aff = fmin(a_aff, -sl / dsl);
                if (dzl < 0) a_aff = fmin(a_aff, -zl / dzl);
                pd_sh[r] = dsh; pd_zh[r] = dzh; pd_sl[r] = dsl; pd_zl[r] = dzl;
            }
            for (int r = c.tid; r < nl; r += c.nthr) {
                const int pt = l_pt[r];
                const double n0 = l_n0[r], n1 = l_n1[r], n2 = l_n2[r], s = l_s[r], z = l_z[r];
                const int gr = DYN ? l_g[r] : -1;
                const double act = -(n0 * sm.x[pt] + n1 * sm.x[npt + pt] + (D3 ? n2 * sm.x[2 * npt + pt] : 0.0)) + (gr >= 0 ? q_e[gr] : 0.0);
                const double gd = -(n0 * sm.dx[pt] + n1 * sm.dx[npt + pt] + (D3 ? n2 * sm.dx[2 * npt + pt] : 0.0)) + (gr >= 0 ? q_de[gr] : 0.0);
                const double ds = -(act + s - l_b[r]) - gd;
                const double dz = -z - l_dz[r] * ds;          // l_dz still holds D = z/s
                if (ds < 0) a_aff = fmin(a_aff, -s / ds);
                if (dz < 0) a_aff = fmin(a_aff, -z / dz);
                l_ds[r] = ds; l_c[r] = dz;                    // affine directions (l_c as temporary)
            }
            if (DYN)
                for (int g = c.tid; g < NS; g += c.nthr) {
                    const double sb = q_sb[g], zb = q_zb[g];
                    const double ds = -(q_e[g] + sb) - q_de[g], dz = -zb - zb / sb * ds;
                    if (ds < 0) a_aff = fmin(a_aff, -sb / ds);
                    if (dz < 0) a_aff = fmin(a_aff, -zb / dz);
                    q_dsb[g] = ds; q_dzb[g] = dz;
                }
            { double d0 = 0.0, d1 = 0.0; c.reduce3(a_aff, 2, d0, 0, d1, 0); }
            double mu_aff = 0.0;
            for (int r = c.tid; r < np; r += c.nthr) {
                const double dsh = pd_sh[r], dzh = pd_zh[r], dsl = pd_sl[r], dzl = pd_zl[r];
                mu_aff += (ps_hi[r] + a_aff * dsh) * (pz_hi[r] + a_aff * dzh) + (ps_lo[r] + a_aff * dsl) * (pz_lo[r] + a_aff * dzl);
                pc_hi[r] = dsh * dzh; pc_lo[r] = dsl * dzl;
            }
            for (int r = c.tid; r < nl; r += c.nthr) {
                const double ds = l_ds[r], dz = l_c[r];
                mu_aff += (l_s[r] + a_aff * ds) * (l_z[r] + a_aff * dz);
                l_c[r] = ds * dz;
            }
            if (DYN)
                for (int g = c.tid; g < NS; g += c.nthr) {
                    mu_aff += (q_sb[g] + a_aff * q_dsb[g]) * (q_zb[g] + a_aff * q_dzb[g]);
                    q_pb[g] = q_dsb[g] * q_dzb[g];
                }
            { double d0 = 0.0, d1 = 0.0; c.reduce3(mu_aff, 0, d0, 0, d1, 0); }
            mu_aff *= inv_rows;
            const double ratio = mu > 0 ? mu_aff / mu : 0.0;
            const double sig_mu = ratio * ratio * ratio * mu;

            // ============ corrector ============
            for (int r = c.tid; r < np; r += c.nthr) {
                const double act = p_act[r];
                const double sh = ps_hi[r], zh = pz_hi[r], sl = ps_lo[r], zl = pz_lo[r];
                const double rph = act + sh - pb_hi[r], rpl = -act + sl + pb_lo[r];
                const double rch = sh * zh + pc_hi[r] - sig_mu, rcl = sl * zl + pc_lo[r] - sig_mu;
                sm.V2[r] = (-rch + zh * rph) / sh - (-rcl + zl * rpl) / sl;
            }
            for (int r = c.tid; r < nl; r += c.nthr) {
                const int pt = l_pt[r];
                const double n0 = l_n0[r], n1 = l_n1[r], n2 = l_n2[r], s = l_s[r], z = l_z[r];
                const double eg = (DYN && l_g[r] >= 0) ? q_e[l_g[r]] : 0.0;
                const double act = -(n0 * sm.x[pt] + n1 * sm.x[npt + pt] + (D3 ? n2 * sm.x[2 * npt + pt] : 0.0)) + eg;
                const double rp = act + s - l_b[r];
                const double rc = s * z + l_c[r] - sig_mu;
                l_ds[r] = (-rc + z * rp) / s;                 // c_corr (temporary); l_c keeps rc's corrector part
            }
            if (DYN)
                for (int g = c.tid; g < NS; g += c.nthr) {
                    const double sb = q_sb[g], zb = q_zb[g];
                    q_ccb[g] = (-(sb * zb + q_pb[g] - sig_mu) + zb * (q_e[g] + sb)) / sb;
                }
            c.sync();
            if (DYN) dyn_groups(false);
            for (int pt = c.tid; pt < npt; pt += c.nthr) {
                double u20 = 0, u21 = 0, u22 = 0;
                for (int r = sm.off[pt]; r < sm.off[pt + 1]; r++) {
                    const double cv = l_ds[r];
                    u20 -= l_n0[r] * cv; u21 -= l_n1[r] * cv; u22 -= l_n2[r] * cv;
                }
                sm.ax2[pt] = pair_transpose(T, sm.V2, pt) + u20;
                sm.ax2[npt + pt] = pair_transpose(T, sm.V2 + npl, pt) + u21;
                if (D3) sm.ax2[2 * npt + pt] = pair_transpose(T, sm.V2 + 2 * npl, pt) + u22;
            }
            c.sync();
            for (int p = c.tid; p < ny; p += c.nthr) sm.dy[p] = -sm.rd[p] - gather_y(T, p, sm.ax2);
            c.sync();
            if (DYN) { dyn_reduce_rhs(); c.sync(); }
            ldl_solve(c, sm.W, sm.invp, sm.dy, ny);
            c.sync();
            if (DYN) dyn_back();
            map_x(c, T, sm.dy, nullptr, sm.dx);
            c.sync();
            double a_st = 1.0;
            for (int k = 0; k < D; k++)
            for (int idx = c.tid; idx < npl; idx += c.nthr) {
                const int r = k * npl + idx;
                const double act = p_act[r], gd = pair_eval(T, pair_decode(T, idx), sm.dx + k * npt);
                const double sh = ps_hi[r], zh = pz_hi[r], sl = ps_lo[r], zl = pz_lo[r];
                const double rch = sh * zh + pc_hi[r] - sig_mu, rcl = sl * zl + pc_lo[r] - sig_mu;
                const double dsh = -(act + sh - pb_hi[r]) - gd, dsl = -(-act + sl + pb_lo[r]) + gd;
                const double dzh = (-rch - zh * dsh) / sh, dzl = (-rcl - zl * dsl) / sl;
                if (dsh < 0) a_st = fmin(a_st, -sh / dsh);
                if (dzh < 0) a_st = fmin(a_st, -zh / dzh);
                if (dsl < 0) a_st = fmin(a_st, -sl / dsl);
                if (dzl < 0) a_st = fmin(a_st, -zl / dzl);
                pd_sh[r] = dsh; pd_zh[r] = dzh; pd_sl[r] = dsl; pd_zl[r] = dzl;
            }
            for (int r = c.tid; r < nl; r += c.nthr) {
                const int pt = l_pt[r];
                const double n0 = l_n0[r], n1 = l_n1[r], n2 = l_n2[r], s = l_s[r], z = l_z[r];
                const int gr = DYN ? l_g[r] : -1;
                const double act = -(n0 * sm.x[pt] + n1 * sm.x[npt + pt] + (D3 ? n2 * sm.x[2 * npt + pt] : 0.0)) + (gr >= 0 ? q_e[gr] : 0.0);
                const double gd = -(n0 * sm.dx[pt] + n1 * sm.dx[npt + pt] + (D3 ? n2 * sm.dx[2 * npt + pt] : 0.0)) + (gr >= 0 ? q_de[gr] : 0.0);
                const double rc = s * z + l_c[r] - sig_mu;
                const double ds = -(act + s - l_b[r]) - gd;
                const double dz = (-rc - z * ds) / s;
                if (ds < 0) a_st = fmin(a_st, -s / ds);
                if (dz < 0) a_st = fmin(a_st, -z / dz);
                l_ds[r] = ds; l_dz[r] = dz;
            }
            if (DYN)
                for (int g = c.tid; g < NS; g += c.nthr) {
                    const double sb = q_sb[g], zb = q_zb[g];
                    const double rc = sb * zb + q_pb[g] - sig_mu;
                    const double ds = -(q_e[g] + sb) - q_de[g], dz = (-rc - zb * ds) / sb;
                    if (ds < 0) a_st = fmin(a_st, -sb / ds);
                    if (dz < 0) a_st = fmin(a_st, -zb / dz);
                    q_dsb[g] = ds; q_dzb[g] = dz;
                }
            { double d0 = 0.0, d1 = 0.0; c.reduce3(a_st, 2, d0, 0, d1, 0); }
            a_st = fmin(1.0, 0.995 * a_st);
#if defined(DLSC_QP_TRACE) && !defined(__CUDA_ARCH__)
            printf("      a_aff %.3e mu_aff %.3e sig_mu %.3e a %.3e\n", a_aff, mu_aff, sig_mu, a_st);
#endif
            // ============ update ============
            for (int r = c.tid; r < np; r += c.nthr) {
                ps_hi[r] += a_st * pd_sh[r]; pz_hi[r] += a_st * pd_zh[r];
                ps_lo[r] += a_st * pd_sl[r]; pz_lo[r] += a_st * pd_zl[r];
            }
            for (int r = c.tid; r < nl; r += c.nthr) { l_s[r] += a_st * l_ds[r]; l_z[r] += a_st * l_dz[r]; }
            if (DYN)
                for (int g = c.tid; g < NS; g += c.nthr) { q_e[g] += a_st * q_de[g]; q_sb[g] += a_st * q_dsb[g]; q_zb[g] += a_st * q_dzb[g]; }
            for (int p = c.tid; p < ny; p += c.nthr) sm.y[p] += a_st * sm.dy[p];
            c.sync();
            map_x(c, T, sm.y, sm.cst, sm.x);
            c.sync();
        }
        it_total += it;
        if (status == kStQpMaxIter && acceptable) status = 0;
        if (status != 0) break;
        // ---- check EVERY LSC row at the solution (the screen is exact only if none is violated) ----
        viol_lsc = 0.0;
        for (int pt = c.tid; pt < npt; pt += c.nthr) {
            if (pt < 3) continue;
            const int m = pt / kP;
            const double x0 = sm.x[pt], x1 = sm.x[npt + pt], x2 = D3 ? sm.x[2 * npt + pt] : 0.0;
            for (int cc = 0; cc < K; cc++) {
                if (!sm.act[m * Kc + cc]) continue;
                const LscRowData r = lsc_row_data(P, in, pt, cc, nd);
                const double v = -(r.n0 * x0 + r.n1 * x1 + r.n2 * x2) - r.b + (cc < nd ? q_e[cc * M + m] : 0.0);
                viol_lsc = fmax(viol_lsc, v);
            }
        }
        { double d0 = 0.0, d1 = 0.0; c.reduce3(viol_lsc, 1, d0, 0, d1, 0); }
        if (all_rows || viol_lsc <= 1e-9) break;
        tau *= 4.0;
        status = kStQpMaxIter;
    }

    // ---- outputs: objective in x-space (constant included, like IloCplex::getObjValue :109) ----
    double obj = 0.0, viol = viol_lsc;
    for (int k = 0; k < D; k++)
        for (int pt = c.tid; pt < npt; pt += c.nthr) {
            const int m = pt / kP, i = pt - m * kP;
            const double* xs = sm.x + k * npt + m * kP;
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < kP; j++) v += T.Q2[i * kP + j] * xs[j];
            obj += 0.5 * v * xs[i];
            if (i == n && m >= M - qc.ts) {
                const double g = (double)v3_get(in.goal, k);
                obj += wT * xs[n] * xs[n] - wT2 * g * xs[n] + wT * g * g;
            }
        }
    for (int k = 0; k < D; k++)
        for (int idx = c.tid; idx < npl; idx += c.nthr) {
            const int r = k * npl + idx;
            const double act = pair_eval(T, pair_decode(T, idx), sm.x + k * npt);
            viol = fmax(viol, fmax(act - pb_hi[r], pb_lo[r] - act));
        }
    if (DYN) {
        for (int g = c.tid; g < NS; g += c.nthr) {
            obj += P.slack_w * ((double)(M - g % M) / M) * q_e[g] * q_e[g];
            viol = fmax(viol, q_e[g]);                                    // eps <= 0
            if (out.slack) out.slack[g] = (status == 0) ? q_e[g] : 0.0;
        }
    }
    { double d1 = 0.0; c.reduce3(obj, 0, viol, 1, d1, 0); }
    if (c.tid == 0) {
        *out.cost = obj; *out.viol = viol; *out.iters = it_total; *out.status |= status | ((P.qp_solver != 1 && !solved_by_gi) ? kStQpIpmUsed : 0);
        if (out.rows) *out.rows = rows_total;
    }
    const bool ok = (status == 0);
    for (int e = c.tid; e < npt; e += c.nthr) {
        float v0, v1, v2;
        if (ok) {                                                                     // :71-83
            v0 = (float)sm.x[e]; v1 = (float)sm.x[npt + e];
            v2 = D3 ? (float)sm.x[2 * npt + e] : (float)P.z_2d;
        } else {                                                                      // failsafe traj_planner.cpp:775-776
            v0 = in.init_traj[e * 3]; v1 = in.init_traj[e * 3 + 1]; v2 = in.init_traj[e * 3 + 2];
        }
        float* o = out.traj + e * 3;
        o[0] = v0; o[1] = v1; o[2] = v2;
        if (out.traj_host) { float* h = out.traj_host + e * 3; h[0] = v0; h[1] = v1; h[2] = v2; }   // streams out over PCIe while the other agents solve
    }
    if (out.x)
        for (int e = c.tid; e < nx; e += c.nthr) out.x[e] = sm.x[e];
    c.sync();
}

}  // namespace dlsc
