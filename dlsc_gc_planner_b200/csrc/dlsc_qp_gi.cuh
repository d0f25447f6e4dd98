// dlsc_qp_gi.cuh -- primary QP path: dual active set (Goldfarb-Idnani, Schur form) straight off the LSC / SFC
// arrays, one small CTA per agent.
//
// Same problem as TrajOptimizer::populatebyrow (reference src/traj_optimizer.cpp:225-527).  What the replan
// QPs look like in practice (4096-agent forest, measured): 60 % of the agents have NO active inequality at
// the optimum, 28 % have one, the maximum is ~11.  So the work per agent is
//   (A) y0 = -H^-1 g, the unconstrained optimum.  g is linear in (c0, c1, c2, goal) per axis, so y0 is a
//       tabulated [nyd x 4] map per terminal-segment count (QpTab::Y0) -- 4 FMAs per unknown;
//   (B) one pass over every inequality row to find the most violated one.  Rows are evaluated from their
//       sources (two-sided pattern rows by stencil arithmetic, LSC rows from normal / d / anchor arrays in a
//       flat neighbour-major index space so that consecutive threads read consecutive control points);
//       nothing is cached or copied per agent;
//   (C) only if something is violated: the H^-1 block is staged in shared memory and the active-set
//       iterations of dlsc_qp.cuh run (same algebra as qp_dual_active_set), re-scanning from the sources.
// Agents on which the active set gives up (more than kGiQ simultaneously active rows, loss of positive
// definiteness of the Schur complement; with dynamic obstacles already at kGiDynRows rows or kGiDynIters iterations)
// are queued for the interior-point kernel (dlsc_qp.cuh qp_agent), which carries the slack variables as well.
#pragma once
#include "dlsc_qp.cuh"

namespace dlsc {

constexpr int kSfcStage = kMaxM * 6 / 2;        // doubles holding the agent's [M][6] float boxes
template <int Q = kGiQ>
DLSC_HD size_t gi_smem_doubles(const QpTab& T, int Kcap) {
    return gi_doubles<Q>(T) + 3 * (size_t)T.ny + (size_t)T.nx + 16 + kSfcStage + kMaxM + 96 + (size_t)kMaxDyn * T.M + ((size_t)Kcap + 1) / 2;
}
template <int Q = kGiQ>
DLSC_HD void gi_smem_carve(const QpTab& T, double* base, QpSmem& s) {
    double* p = base;
    s.W = p; p += gi_doubles<Q>(T);
    s.y = p; p += T.ny; s.dy = p; p += T.ny; s.ax1 = p; p += T.ny;
    s.x = p; p += T.nx;
    s.cst = p; p += 16;
    s.sfcs = reinterpret_cast<float*>(p); p += kSfcStage;
    s.dev = p; p += kMaxM;
    s.red = p; p += 96;
    s.esl = p; p += kMaxDyn * T.M;
    s.off = reinterpret_cast<int*>(p);          // [Kcap] global indices of the neighbours
    s.invp = s.pan = s.rd = s.dx = s.ax2 = s.V1 = s.V2 = s.DD = s.S = nullptr;
    s.act = nullptr;
}
// the fast path (qp_agent_fast, one warp per agent) only needs y, x and the staged inputs
DLSC_HD size_t fast_smem_doubles(const QpTab& T, int Kcap) {
    return (size_t)T.ny + (size_t)T.nx + 16 + kSfcStage + kMaxM + ((size_t)Kcap + 1) / 2;
}
DLSC_HD void fast_smem_carve(const QpTab& T, double* base, QpSmem& s) {
    double* p = base;
    s.y = p; p += T.ny;
    s.x = p; p += T.nx;
    s.cst = p; p += 16;
    s.sfcs = reinterpret_cast<float*>(p); p += kSfcStage;
    s.dev = p; p += kMaxM;
    s.off = reinterpret_cast<int*>(p);
    s.W = s.dy = s.ax1 = s.red = s.esl = nullptr;
    s.invp = s.pan = s.rd = s.dx = s.ax2 = s.V1 = s.V2 = s.DD = s.S = nullptr;
    s.act = nullptr;
}

DLSC_HD float int_as_f32(int i) {
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}
DLSC_HD int f32_as_int(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4); return i;
#endif
}

#ifndef DLSC_GI_DYN_ITERS
#define DLSC_GI_DYN_ITERS 16
#endif
#ifndef DLSC_GI_DYN_ROWS
#define DLSC_GI_DYN_ROWS 12
#endif
constexpr int kGiDynIters = DLSC_GI_DYN_ITERS;    // ... and its iteration cap in that case
constexpr int kGiDynRows = DLSC_GI_DYN_ROWS;     // hand-over threshold of the active set when dynamic obstacles are present
constexpr double kGiTol = 1e-10;   // accepted violation [m]: the objective error it admits is (multiplier x tol) <= ~1e-8

struct PairRowD { int fam, k, pa, pb; };
DLSC_HD PairRowD pair_desc(const QpTab& T, int r) {
#ifdef __CUDA_ARCH__
    const uint32_t d = __ldg(T.pr_desc + r);
#else
    const uint32_t d = T.pr_desc[r];
#endif
    PairRowD o; o.fam = d & 3; o.k = (d >> 2) & 3; o.pa = (d >> 4) & 0xff; o.pb = (d >> 12) & 0xff;
    return o;
}

// violation of the two sides of pattern row r at x
DLSC_HD void pair_violation(const DevParams& P, const QpTab& T, const QpIn& in, const QpConst& qc, int r, const double* x,
                            double& vh, double& vl) {
    const PairRowD d = pair_desc(T, r);
    PairRow pr; pr.fam = d.fam; pr.pa = d.pa; pr.pb = d.pb;
    const double act = pair_eval(T, pr, x + d.k * T.npt);
    double lo, hi;
    pair_bounds(P, T, in, qc, d.k, pr, lo, hi);
    vh = act - hi; vl = lo - act;
}

// Dynamic-obstacle LSC rows carry a slack variable per (obstacle, segment): -n.x + eps <= b, eps <= 0, cost
// w (M-m)/M eps^2 (traj_optimizer.cpp:272-283, 317-331, 436-448).  The slack block of H is diagonal and every row
// touches at most one slack with coefficient +1, so the active set keeps the slack values beside y (sm.esl) and adds
// the slack terms to the few inner products that need them.  eps <= 0 holds by itself: eps = -(sum of the group's
// multipliers) / h and the multipliers never go negative.
// slack variable of inequality row `id`, or -1
DLSC_HD int gi_slack_of(const DevParams& P, const QpTab& T, int id) {
    if (P.n_dyn == 0 || id < 2 * T.np) return -1;
    const int o = id - 2 * T.np, pt = o / P.K, cc = o - pt * P.K;
    return cc < P.n_dyn ? cc * P.M + pt / kP : -1;
}
DLSC_HD double gi_slack_hinv(const DevParams& P, int s) {          // 1 / (2 w (M - m) / M)
    const int m = s % P.M;
    return 1.0 / (2.0 * P.slack_w * ((double)(P.M - m) / P.M));
}

// most violated row over all inequality rows; every thread returns the same (vmax, id)
//   id < 2 np: pattern row r = id >> 1, side id & 1 (0: upper, 1: lower);  else LSC row o = id - 2 np = pt * Kcap + cc
// Row screen: k_lsc stored, per (neighbour, segment) item, the smallest normalised slack of its rows at the agent's
// initial trajectory.  With dev[m] = the largest |x_pt - init_pt| over the constrained points of segment m of the
// iterate being scanned, a row of slack s can only be violated when dev[m] >= s (Cauchy-Schwarz), so items with
// slack > dev[m] + margin are not evaluated (margin 1e-3 m covers the rounding of the two evaluations).  Exact, and
// it adapts to every iterate: far from the initial trajectory more rows are evaluated, never fewer than needed.
constexpr float kScreenMargin = 1e-3f;
// dev2: squared per-segment deviations as float bit patterns (positive floats order like ints: gi_map_x_dev)
// mine: out, this thread evaluated the winning row (exactly one thread: every row is evaluated by one thread);
// row: out (may be null), the thread's best LSC row as lsc_row_data would return it (valid when its best is an LSC row)
template <bool DYN = false>
DLSC_HD void gi_scan(const Cta& c, const DevParams& P, const QpTab& T, const QpIn& in, const QpConst& qc, const double* x,
                     const int* sm_nbr, bool screened, const int* dev2, double& vmax_out, double& id_out,
                     double& nviol_out, bool& mine, LscRowData* row, const double* esl) {
    const int npt = T.npt, np = T.np, Kc = P.K, K = in.K, nd = DYN ? P.n_dyn : 0;
    const bool D3 = (P.D == 3);
    double best = -1e300, best_id = 1e300;
    int n_bad = 0;                                                 // violated rows seen by this thread
    for (int r = c.tid; r < np; r += c.nthr) {
        double vh, vl;
        pair_violation(P, T, in, qc, r, x, vh, vl);
        if (vh > best) { best = vh; best_id = 2.0 * r; }
        if (vl > best) { best = vl; best_id = 2.0 * r + 1.0; }
        n_bad += (vh > kGiTol) + (vl > kGiTol);
    }
    // LSC rows: one work item = one (neighbour, segment) = up to 6 rows sharing a normal; its d values and
    // anchors are contiguous in memory (traj_optimizer.cpp:412-450: -n.x <= -(n.anchor + d))
    const int M = P.M, items = K * M;
    for (int e = c.tid; e < items; e += c.nthr) {
        const int cc = e / M, m = e - cc * M;
        if (screened) {                                                             // row screen
            const float s = in.near[e] - kScreenMargin;
            if (s > 0.f && s * s > int_as_f32(dev2[m])) continue;
        }
        const float* nr = in.normal + ((size_t)cc * M + m) * 3;
        const V3 nv = v3_load(nr);
        const double* dd = in.d + ((size_t)cc * M + m) * kP;
        const bool last = (m == M - 1) && cc >= nd;                                 // dynamic obstacles: predicted points throughout
        const float* an = last ? in.anchor_last + cc * 3 : in.pred_traj + ((size_t)sm_nbr[cc] * npt + m * kP) * 3;
        const double es = (DYN && esl && cc < nd) ? esl[cc * M + m] : 0.0;
        double dv[kP]; float av[kP][3];
#pragma unroll
        for (int i = 0; i < kP; i++) {
            dv[i] = dd[i];
            const float* a = last ? an : an + i * 3;
            av[i][0] = a[0]; av[i][1] = a[1]; av[i][2] = a[2];
        }
        if (v3_norm(nv) < kEpsF) continue;                                          // :422-424
        const double n0 = (double)nv.x, n1 = (double)nv.y, n2 = D3 ? (double)nv.z : 0.0;
#pragma unroll
        for (int i = 0; i < kP; i++) {
            if (m == 0 && i < 3) continue;                                          // :417-419
            const int pt = m * kP + i;
            double b = -dv[i];                                                      // same order as lsc_row_data
            b -= n0 * (double)av[i][0];
            b -= n1 * (double)av[i][1];
            if (D3) b -= n2 * (double)av[i][2];
            const double v = -(n0 * x[pt] + n1 * x[npt + pt] + (D3 ? n2 * x[2 * npt + pt] : 0.0)) - b + es;
            const double id = 2.0 * np + (double)(pt * Kc + cc);
            if (v > best || (v == best && id < best_id)) {
                best = v; best_id = id;
                if (row) { row->n0 = n0; row->n1 = n1; row->n2 = n2; row->b = b; }
            }
            n_bad += (v > kGiTol);
        }
    }
    const double my_best = best, my_id = best_id;
    double nvd = (double)n_bad, unused = 0.0;
    c.reduce_argmax(best, best_id, unused, nvd);
    mine = (my_best == best && my_id == best_id);
    vmax_out = best; id_out = best_id; nviol_out = nvd;
}

// x = c + T y (map_x), and on the way the squared deviation of every constrained control point from the initial
// trajectory, max-reduced per segment into dev2[m] (float bit patterns through an integer atomic max: order
// independent, hence deterministic; inflated by 1e-5 so that the float rounding can only make the screen weaker).
// dev2 must be zero on entry (gi_prologue / the end of the previous scan).
DLSC_HD void gi_map_x_dev(const Cta& c, const DevParams& P, const QpTab& T, const QpIn& in, const double* y, const double* cst,
                          double* x, int* dev2, bool screened) {
    const int M = T.M, nyd = T.nyd, npt = T.npt, D = T.D;
    for (int pt = c.tid; pt < npt; pt += c.nthr) {
        const int m = pt / kP, i = pt - m * kP;
        double r2 = 0.0;
        for (int k = 0; k < D; k++) {
            const double* yk = y + k * nyd;
            double v;
            if (i >= 3) v = (m == M - 1) ? yk[3 * (M - 1)] : yk[3 * m + i - 3];
            else if (m == 0) v = cst[k * 3 + i];
            else {
                const double* q = yk + 3 * (m - 1);
                v = (i == 0) ? q[2] : (i == 1 ? 2.0 * q[2] - q[1] : 4.0 * q[2] - 4.0 * q[1] + q[0]);
            }
            x[k * npt + pt] = v;
            if (screened) { const double e = v - (double)in.init_traj[pt * 3 + k]; r2 += e * e; }
        }
        if (screened && pt >= 3) {                                  // points 0..2 of segment 0 carry no LSC rows
            const int bits = f32_as_int((float)r2 * 1.00001f + 1e-30f);
#ifdef __CUDA_ARCH__
            atomicMax(dev2 + m, bits);
#else
            if (bits > dev2[m]) dev2[m] = bits;
#endif
        }
    }
}

// map y -> x with the per-segment deviations, then one scan for the most violated row; leaves dev2 zeroed
template <bool DYN = false>
DLSC_HD void gi_map_and_scan(const Cta& c, const DevParams& P, const QpTab& T, const QpIn& in, const QpConst& qc,
                             const QpSmem& sm, bool screened, double& vmax, double& idsel, bool& mine, LscRowData* row,
                             double* nviol = nullptr) {
    int* dev2 = reinterpret_cast<int*>(sm.dev);
    gi_map_x_dev(c, P, T, in, sm.y, sm.cst, sm.x, dev2, screened);
    c.sync();
    c.tick(3);
    double nv = 0.0;
    gi_scan<DYN>(c, P, T, in, qc, sm.x, sm.off, screened, dev2, vmax, idsel, nv, mine, row, DYN ? sm.esl : nullptr);
    // every thread is past the reduction, so nobody reads dev2 any more: reset it for the next map.  In warp mode the
    // reduction is shuffles only, which order execution but are no memory barrier (racecheck flags the reset against
    // the scan's reads): __syncwarp makes the ordering explicit.
    if (c.warp) c.wsync();
    for (int m = c.tid; m < P.M; m += c.nthr) dev2[m] = 0;
    if (nviol) *nviol = nv;
    c.tick(5);
}

// Returns 0 = optimal, kStQpMaxIter = infeasible, -1 = give up.  y in sm.y, x in sm.x on return.
template <int Q = kGiQ, bool DYN = false>
DLSC_HD int gi_solve(const Cta& c, const DevParams& P, const QpTab& T, const QpIn& in, const QpConst& qc,
                     const QpSmem& sm, const double* seed, int* iters_out, double* viol_out) {
    const int M = P.M, D = P.D, ny = T.ny, nyd = T.nyd, npt = T.npt, np = T.np, Kc = P.K;
    const bool D3 = (D == 3);
    GiSmem g;
    gi_carve<Q>(T, sm.W, g);
    const double tol = kGiTol;
    int q = 0, iters = 0, status = -1;
    double viol_p = 0.0, u_p = 0.0;
    bool same_p = false, have_hinv = false;
    const bool screened = (P.qp_screen > 0) && (in.near != nullptr);      // row screen (gi_scan); qp_screen <= 0 disables it
    bool x_current = false;      // sm.x == map_x(sm.y)
    for (int guard = 0; guard < 400; guard++) {
        if (!same_p) {
            double vmax, idsel;
            bool mine;                         // this thread builds the candidate row
            LscRowData row;
            bool have_row = false;
            if (guard == 0 && seed) {          // the fast path already scanned the unconstrained optimum
                map_x(c, T, sm.y, sm.cst, sm.x);
                c.sync();
                vmax = seed[0]; idsel = seed[1];
                mine = (c.tid == 0);
            } else {
                gi_map_and_scan<DYN>(c, P, T, in, qc, sm, screened, vmax, idsel, mine, &row);
                have_row = true;
            }
            x_current = true;
            *viol_out = vmax;
            if (!(vmax > tol)) { status = 0; break; }
            if (!have_hinv) {                                   // first violated row: stage this agent's H^-1 block
                const double* src = T.Hinv + (size_t)(qc.ts - 1) * nyd * nyd;
                for (int e = c.tid; e < nyd * nyd; e += c.nthr) g.Hinv[e] = src[e];
                have_hinv = true;
            }
            // ---- candidate row p: x-space form -> y-space form, by the thread that evaluated it in the scan (it
            //      still holds the row's coefficients; thread 0 re-reads them after a seeded start) ----
            if (mine) {
                const int id = (int)idsel;
                int xi[3] = {0, 0, 0}; double xc[3] = {0, 0, 0}; int nxe = 0; double bp;
                if (id < 2 * np) {
                    const int r = id >> 1, side = id & 1;
                    const PairRowD pd = pair_desc(T, r);
                    PairRow pr; pr.fam = pd.fam; pr.pa = pd.pa; pr.pb = pd.pb;
                    double lo, hi;
                    pair_bounds(P, T, in, qc, pd.k, pr, lo, hi);
                    const double sg = side ? -1.0 : 1.0;
                    const int base = pd.k * npt;
                    if (pr.fam == 0) { xi[0] = base + pr.pa; xc[0] = sg; nxe = 1; }
                    else if (pr.fam == 1) { xi[0] = base + pr.pa + 1; xc[0] = sg * T.scv; xi[1] = base + pr.pa; xc[1] = -sg * T.scv; nxe = 2; }
                    else if (pr.fam == 2) {
                        xi[0] = base + pr.pa + 2; xc[0] = sg * T.sca; xi[1] = base + pr.pa + 1; xc[1] = -2.0 * sg * T.sca;
                        xi[2] = base + pr.pa; xc[2] = sg * T.sca; nxe = 3;
                    } else { xi[0] = base + pr.pa; xc[0] = sg; xi[1] = base + pr.pb; xc[1] = -sg; nxe = 2; }
                    bp = side ? -lo : hi;
                } else {
                    const int o = id - 2 * np, pt = o / Kc, cc = o - pt * Kc;
                    const LscRowData rw = have_row ? row : lsc_row_data(P, in, pt, cc, DYN ? P.n_dyn : 0);
                    xi[0] = pt; xc[0] = -rw.n0; xi[1] = npt + pt; xc[1] = -rw.n1; nxe = 2;
                    if (D3) { xi[2] = 2 * npt + pt; xc[2] = -rw.n2; nxe = 3; }
                    bp = rw.b;
                }
                double* yc = g.yc + 9 * Q; int16_t* yi = g.yi + 9 * Q;
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
                g.bq[Q] = bp; g.id[Q] = id;
            }
            viol_p = vmax; u_p = 0.0;
            c.sync();
            c.tick(8);
            // ---- w = H^-1 a_p (into sm.dy) ----
            for (int p = c.tid; p < ny; p += c.nthr) {
                const int k = p / nyd, a = p - k * nyd;
                const double* hr = g.Hinv + a * nyd;
                const double* yc = g.yc + 9 * Q; const int16_t* yi = g.yi + 9 * Q;
                double v = 0.0;
#pragma unroll
                for (int t = 0; t < 9; t++) {
                    const int j = yi[t];
                    if (j >= 0 && j / nyd == k) v += yc[t] * hr[j - k * nyd];
                }
                sm.dy[p] = v;
            }
            c.sync();
            c.tick(9);
        }
        iters++;
        if (DYN && iters > kGiDynIters) { status = -1; break; }      // uniform: long add / drop sequences go to the interior point too
        // ---- small dense step (first warp): r, step lengths, active-set update; t_y = A' r into sm.ax1[0..ny) ----
        // The q x q algebra stays with lane 0; the two gather loops over the rows' y-space terms (v = A w before it,
        // t_y = A' r after it) are spread over the lanes.  Every sum keeps the serial order of its terms, so the result
        // is bit-identical to a single-thread evaluation (the host simulator runs this with one "lane").
        const int W = c.nthr < 32 ? c.nthr : 32;
        if (c.tid < W) {
            // slack of the candidate row: (H^-1 a_p) has one more component, 1/h, in the slack block
            const int sp = DYN ? gi_slack_of(P, T, g.id[Q]) : -1;
            const double hp = sp >= 0 ? gi_slack_hinv(P, sp) : 0.0;
            for (int j = c.tid; j < q; j += W) {
                double vj = 0.0;
                for (int t = 0; t < 9; t++) { const int jj = g.yi[9 * j + t]; if (jj >= 0) vj += g.yc[9 * j + t] * sm.dy[jj]; }
                if (DYN && sp >= 0 && gi_slack_of(P, T, g.id[j]) == sp) vj += hp;
                g.v[j] = vj;
            }
            c.wsync();
            const double* yc = g.yc + 9 * Q; const int16_t* yi = g.yi + 9 * Q;
            if (c.tid == 0) {
                double apw = 0.0;
                for (int t = 0; t < 9; t++) if (yi[t] >= 0) apw += yc[t] * sm.dy[yi[t]];
                apw += hp;
                double ll = 0.0;
                for (int j = 0; j < q; j++) {
                    double a = g.v[j];
                    for (int k = 0; k < j; k++) a -= g.Ls[j * (j + 1) / 2 + k] * g.l[k];
                    a *= g.li[j];                              // 1 / L_jj, kept beside L: no division on the dependent chain
                    g.l[j] = a; ll += a * a;
                }
                for (int j = q - 1; j >= 0; j--) {
                    double a = g.l[j];
                    for (int k = j + 1; k < q; k++) a -= g.Ls[k * (k + 1) / 2 + j] * g.r[k];
                    g.r[j] = a * g.li[j];
                }
                const double zn = apw - ll;
                double t1 = 1e300; int kdrop = -1;
                for (int j = 0; j < q; j++)
                    if (g.r[j] > 0) { const double tt = g.u[j] / g.r[j]; if (tt < t1) { t1 = tt; kdrop = j; } }
                const bool dependent = !(zn > 1e-12 * (apw > 1e-300 ? apw : 1e-300));
                const double t2 = dependent ? 1e300 : viol_p / zn;
                const double t = t1 < t2 ? t1 : t2;
                const bool finite = (t < 1e299);
                if (finite) for (int j = 0; j < q; j++) g.u[j] -= t * g.r[j];
                if (DYN && P.n_dyn > 0 && finite && !dependent) {      // slack block of the primal step  z = H^-1 (a_p - A' r)
                    for (int j = 0; j < q; j++) {
                        const int sj = gi_slack_of(P, T, g.id[j]);
                        if (sj >= 0) sm.esl[sj] += t * g.r[j] * gi_slack_hinv(P, sj);
                    }
                    if (sp >= 0) sm.esl[sp] -= t * hp;
                }
                g.ty[0] = dependent ? 0.0 : t;
                g.ty[2] = zn;
                g.ty[3] = finite ? 1.0 : 0.0;
                g.ty[4] = (t2 <= t1) ? 1.0 : 0.0;          // add the candidate (else drop row kdrop)
                g.ty[5] = (double)kdrop;
                g.ty[6] = apw;
                g.ty[7] = t;
            }
            c.wsync();
            const bool finite = (g.ty[3] != 0.0);
            if (finite && q > 0) {                          // t_y = A' r, only read by the y update when q > 0
                for (int e = c.tid; e < ny; e += W) {
                    double acc = 0.0;
                    for (int j = 0; j < q; j++)
                        for (int tt = 0; tt < 9; tt++) if (g.yi[9 * j + tt] == e) acc += g.r[j] * g.yc[9 * j + tt];
                    sm.ax1[e] = acc;
                }
            }
            c.wsync();
            if (c.tid == 0) {
                double flag;             // 0: full step, new scan | 1: partial, same p | 2: infeasible | 3: give up
                if (!finite) flag = 2.0;
                else {
                    const double zn = g.ty[2], apw = g.ty[6];
                    const int kdrop = (int)g.ty[5];
                    u_p += g.ty[7];
                    if (g.ty[4] != 0.0) {
                        // with dynamic obstacles an agent beside an obstacle keeps a row active in most of its slack groups:
                        // the serial q x q algebra of this kernel is the wrong tool beyond a dozen rows, the interior point
                        // (which carries the slack variables) takes over.  Measured (4096 agents, 8 obstacles, QP stage):
                        // 8 rows / 8 iterations 3.9 ms, 12 / 16 4.0 ms, 20 / 32 4.6 ms, 28 / 64 6.6 ms
                        if (q >= ((DYN && P.qp_active_max > kGiDynRows) ? kGiDynRows : P.qp_active_max)) flag = 3.0;
                        else {
                            for (int k = 0; k < q; k++) { g.Ls[q * (q + 1) / 2 + k] = g.l[k]; g.Sm[q * (q + 1) / 2 + k] = g.v[k]; }
                            g.Ls[q * (q + 1) / 2 + q] = sqrt(zn); g.Sm[q * (q + 1) / 2 + q] = apw;
                            g.li[q] = 1.0 / g.Ls[q * (q + 1) / 2 + q];
                            for (int tt = 0; tt < 9; tt++) { g.yc[9 * q + tt] = yc[tt]; g.yi[9 * q + tt] = yi[tt]; }
                            g.bq[q] = g.bq[Q]; g.id[q] = g.id[Q]; g.u[q] = u_p;
                            flag = 0.0;
                        }
                    } else {
                        for (int j = kdrop; j < q - 1; j++) {
                            for (int tt = 0; tt < 9; tt++) { g.yc[9 * j + tt] = g.yc[9 * (j + 1) + tt]; g.yi[9 * j + tt] = g.yi[9 * (j + 1) + tt]; }
                            g.bq[j] = g.bq[j + 1]; g.id[j] = g.id[j + 1]; g.u[j] = g.u[j + 1];
                        }
                        for (int i2 = 0, ii = 0; i2 < q; i2++) {
                            if (i2 == kdrop) continue;
                            for (int k2 = 0, kk = 0; k2 <= i2; k2++) {
                                if (k2 == kdrop) continue;
                                g.Ls[ii * (ii + 1) / 2 + kk] = g.Sm[i2 * (i2 + 1) / 2 + k2];
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
                            g.li[j] = 1.0 / dj;
                            for (int i2 = j + 1; i2 < q - 1; i2++) {
                                double a = g.Sm[i2 * (i2 + 1) / 2 + j];
                                for (int k = 0; k < j; k++) a -= g.Ls[i2 * (i2 + 1) / 2 + k] * g.Ls[j * (j + 1) / 2 + k];
                                g.Ls[i2 * (i2 + 1) / 2 + j] = a * g.li[j];
                            }
                        }
                        flag = okf ? 1.0 : 3.0;
                    }
                }
                g.ty[1] = flag;
            }
        }
        c.sync();
        c.tick(10);
        const double tp = g.ty[0], flag = g.ty[1];
        if (flag == 2.0) { status = kStQpMaxIter; break; }
        if (flag == 3.0) { status = -1; break; }
        x_current = false;
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
        c.tick(11);
    }
    if (!x_current) {
        map_x(c, T, sm.y, sm.cst, sm.x);
        c.sync();
    }
    c.tick(6);
    *iters_out = iters;
    return status;
}

// Per-agent constants, staged inputs and the unconstrained optimum y0 = -H^-1 g = Y0[ts] (c0, c1, c2, goal) in sm.y.
// `in` is redirected to the staged copy of the SFC boxes.
template <bool DYN = false>
DLSC_HD void gi_prologue(const Cta& c, const DevParams& P, const QpTab& T, QpIn& in, const QpSmem& sm, QpConst& qc) {
    const int M = P.M, D = P.D, ny = T.ny, nyd = T.nyd;
    const int n = kP - 1;
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
    for (int cc = c.tid; cc < in.K; cc += c.nthr) sm.off[cc] = in.nbr_idx[cc];
    for (int m = c.tid; m < kMaxM * 2; m += c.nthr) reinterpret_cast<int*>(sm.dev)[m] = 0;
    if (DYN && sm.esl) for (int e = c.tid; e < P.n_dyn * M; e += c.nthr) sm.esl[e] = 0.0;
    if (P.use_sfc) {
        for (int e = c.tid; e < M * 6; e += c.nthr) sm.sfcs[e] = in.sfc[e];
        in.sfc = sm.sfcs;
    }
    c.sync();
    c.tick(1);
    {
        const double* Y = T.Y0 + (size_t)(qc.ts - 1) * nyd * 4;
        for (int p = c.tid; p < ny; p += c.nthr) {
            const int k = p / nyd, a = p - k * nyd;
            const double* ya = Y + a * 4;
            sm.y[p] = ya[0] * sm.cst[k * 3] + ya[1] * sm.cst[k * 3 + 1] + ya[2] * sm.cst[k * 3 + 2] +
                      ya[3] * (double)v3_get(in.goal, k);
        }
    }
    c.sync();
    c.tick(2);
}

// outputs: objective in x-space (constant included, like IloCplex::getObjValue :109), trajectory or failsafe
template <bool DYN = false>
DLSC_HD void gi_epilogue(const Cta& c, const DevParams& P, const QpTab& T, const QpIn& in, const QpOut& out,
                         const QpConst& qc, const QpSmem& sm, int status, int iters, double vmax) {
    const int M = P.M, D = P.D, npt = T.npt, nx = T.nx, np = T.np;
    const int n = kP - 1;
    const bool D3 = (D == 3);
    const double wT = P.w_terminal, wT2 = 2.0 * P.w_terminal;
    double obj = 0.0;
    for (int e = c.tid; e < nx; e += c.nthr) {
        const int k = e / npt, pt = e - k * npt;
        const int m = pt / kP, i = pt - m * kP;
        const double* xs = sm.x + k * npt + m * kP;
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < kP; j++) v += T.Q2[i * kP + j] * xs[j];
        obj += 0.5 * v * xs[i];
        if (i == n && m >= M - qc.ts) {
            const double gg = (double)v3_get(in.goal, k);
            obj += wT * xs[n] * xs[n] - wT2 * gg * xs[n] + wT * gg * gg;
        }
    }
    c.reduce1(obj, 0);
    c.tick(7);
    if (DYN && P.n_dyn > 0) {                                                         // :317-331
        if (c.tid == 0 && sm.esl)
            for (int o = 0; o < P.n_dyn; o++)
                for (int m = 0; m < M; m++) {
                    const double e = sm.esl[o * M + m];
                    if (e != 0.0) obj += P.slack_w * ((double)(M - m) / M) * e * e;
                }
        if (out.slack)
            for (int e = c.tid; e < P.n_dyn * M; e += c.nthr) out.slack[e] = (status == 0 && sm.esl) ? sm.esl[e] : 0.0;
    }
    if (c.tid == 0) {
        *out.cost = obj; *out.viol = vmax > 0 ? vmax : 0.0; *out.iters = iters; *out.status |= status;
        if (out.rows) *out.rows = 2LL * np + (long long)in.K * (npt - 3);
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

// One agent.  Returns true when the agent is finished (outputs written: optimal, or infeasible -> failsafe
// trajectory), false when the active set gave up and the interior point must take over (nothing written).
// seed: {vmax, id, screened, violated rows} of the scan at the unconstrained optimum when qp_agent_fast already did it, or null.
template <int Q = kGiQ, bool DYN = false>
DLSC_HD bool qp_agent_gi(const Cta& c, const DevParams& P, const QpTab& T, const QpIn& in_, const QpOut& out,
                         const QpSmem& sm, const double* seed = nullptr) {
    QpIn in = in_;
    QpConst qc;
    gi_prologue<DYN>(c, P, T, in, sm, qc);
    int iters = 0;
    double vmax = 0.0;
    const int status = gi_solve<Q, DYN>(c, P, T, in, qc, sm, seed, &iters, &vmax);
    if (status < 0) return false;
    gi_epilogue<DYN>(c, P, T, in, out, qc, sm, status, iters, vmax);
    return true;
}

// Fast path, one warp per agent: is the unconstrained optimum feasible (60 % of the agents of a swarm in
// transit)?  Then it is the optimum: outputs are written and true is returned.  Otherwise the most violated row
// goes to seed_out {vmax, id, screened} for qp_agent_gi and nothing is written.  sm: fast_smem_carve.
template <bool DYN = false>
DLSC_HD bool qp_agent_fast(const Cta& c, const DevParams& P, const QpTab& T, const QpIn& in_, const QpOut& out,
                           const QpSmem& sm, double* seed_out) {
    QpIn in = in_;
    QpConst qc;
    gi_prologue<DYN>(c, P, T, in, sm, qc);
    const bool screened = (P.qp_screen > 0) && (in.near != nullptr);
    double vmax, idsel, nviol = 0.0;
    bool mine;
    gi_map_and_scan<DYN>(c, P, T, in, qc, sm, screened, vmax, idsel, mine, nullptr, &nviol);
    if (vmax > kGiTol) {
        // seed[3]: number of rows violated at the unconstrained optimum -- a proxy for the active-set iterations
        // the agent will need, used to start the expensive agents first (k_qp_gi)
        if (c.tid == 0) { seed_out[0] = vmax; seed_out[1] = idsel; seed_out[2] = screened ? 1.0 : 0.0; seed_out[3] = nviol; }
        return false;
    }
    gi_epilogue<DYN>(c, P, T, in, out, qc, sm, 0, 0, vmax);
    return true;
}

}  // namespace dlsc
