// dlsc_qp_tables.cpp -- host-side builder of the constant QP structure tables (see dlsc_qp_tables.h).
#include "dlsc_qp_tables.h"

#include <cmath>
#include <map>
#include <utility>

#include "dlsc_types.h"

namespace dlsc {

namespace {
int n_choose_k(int n, int k) {      // polynomial.hpp:9-20
    if (k > n) return 0;
    if (k * 2 > n) k = n - k;
    if (k == 0) return 1;
    int r = n;
    for (int i = 2; i <= k; i++) { r *= (n - i + 1); r /= i; }
    return r;
}
int coef_derivative(int n, int phi) {   // polynomial.hpp:89-99
    if (n < phi) return 0;
    int c = 1;
    for (int i = 0; i < phi; i++) c *= n - i;
    return c;
}
struct Term { int idx; double coef; };
struct Expr { std::vector<Term> t; double cc[3] = {0, 0, 0}; };
void expr_add(Expr& e, int idx, double c) {
    for (auto& t : e.t) if (t.idx == idx) { t.coef += c; return; }
    e.t.push_back({idx, c});
}
}  // namespace

void build_q_base(double dt, double Q[36]) {
    const int P = kP, n = kP - 1, k = 3;
    double B[36] = {0}, Z[36] = {0}, T[36] = {0};
    for (int i = 0; i < P; i++)
        for (int j = 0; j < P; j++)
            if (j >= i) B[i * P + j] = n_choose_k(n, i) * n_choose_k(n - i, n - j) * std::pow(-1, j - i);
    for (int i = 0; i < P; i++)
        for (int j = 0; j < P; j++)
            if (i + j - 2 * k + 1 > 0)
                Z[i * P + j] = (double)coef_derivative(i, k) * coef_derivative(j, k) / (i + j - 2 * k + 1);
    for (int i = 0; i < P; i++)
        for (int j = 0; j < P; j++) {
            double s = 0;
            for (int a = 0; a < P; a++) s += B[i * P + a] * Z[a * P + j];
            T[i * P + j] = s;
        }
    const double sc = std::pow(dt, -2 * k + 1);
    for (int i = 0; i < P; i++)
        for (int j = 0; j < P; j++) {
            double s = 0;
            for (int a = 0; a < P; a++) s += T[i * P + a] * B[j * P + a];
            Q[i * P + j] = s * sc;
        }
}

void build_qp_tables(int M, int D, double dt, double w_control, double w_terminal, bool use_comm, QpTabHost& T) {
    const int n = kP - 1;
    T = QpTabHost();
    T.D = D; T.M = M; T.nyd = 3 * M - 2; T.ny = D * T.nyd; T.npt = M * kP; T.nx = D * T.npt;
    T.ntri = T.ny * (T.ny + 1) / 2; T.ntri_local = T.nyd * (T.nyd + 1) / 2;
    T.use_comm = use_comm ? 1 : 0;
    const int nyd = T.nyd;
    auto yid = [&](int m, int j) { return (m == M - 1) ? 3 * (M - 1) : 3 * m + j; };

    // x map (axis-local)
    std::vector<Expr> xe(T.npt);
    for (int i = 0; i < 3; i++) xe[i].cc[i] = 1.0;                       // x[0][0..2] = c0,c1,c2
    for (int m = 0; m < M; m++) {
        for (int j = 0; j < 3; j++) expr_add(xe[m * kP + 3 + j], yid(m, j), 1.0);
        if (m + 1 < M) {
            const int y3 = yid(m, 0), y4 = yid(m, 1), y5 = yid(m, 2);
            expr_add(xe[(m + 1) * kP + 0], y5, 1.0);
            expr_add(xe[(m + 1) * kP + 1], y5, 2.0); expr_add(xe[(m + 1) * kP + 1], y4, -1.0);
            expr_add(xe[(m + 1) * kP + 2], y5, 4.0); expr_add(xe[(m + 1) * kP + 2], y4, -4.0);
            expr_add(xe[(m + 1) * kP + 2], y3, 1.0);
        }
    }
    T.xm_nv.assign(T.npt, 0); T.xm_cidx.assign(T.npt, -1);
    T.xm_idx.assign(T.npt * 3, 0); T.xm_coef.assign(T.npt * 3, 0.0);
    for (int pt = 0; pt < T.npt; pt++) {
        T.xm_nv[pt] = (int8_t)xe[pt].t.size();
        for (size_t t = 0; t < xe[pt].t.size(); t++) {
            T.xm_idx[pt * 3 + t] = (int16_t)xe[pt].t[t].idx;
            T.xm_coef[pt * 3 + t] = xe[pt].t[t].coef;
        }
        for (int i = 0; i < 3; i++) if (xe[pt].cc[i] != 0.0) T.xm_cidx[pt] = (int8_t)i;
    }

    // objective constants
    T.Qb.assign(36, 0.0); T.Q2.assign(36, 0.0);
    build_q_base(dt, T.Qb.data());
    for (int i = 0; i < 36; i++) T.Q2[i] = 2.0 * w_control * T.Qb[i];
    T.H1.assign((size_t)nyd * nyd, 0.0);
    for (int m = 0; m < M; m++)
        for (int i = 0; i < kP; i++)
            for (int j = 0; j < kP; j++) {
                const double q = T.Q2[i * kP + j];
                for (auto& a : xe[m * kP + i].t)
                    for (auto& b : xe[m * kP + j].t) T.H1[(size_t)a.idx * nyd + b.idx] += q * a.coef * b.coef;
            }

    // H^-1 of one axis block for every number ts of terminal segments (traj_optimizer.cpp:543-551): the dual
    // active-set solver starts from the unconstrained optimum y0 = -H^-1 g and only ever needs H^-1 products
    T.Hinv.assign((size_t)M * nyd * nyd, 0.0);
    for (int ts = 1; ts <= M; ts++) {
        std::vector<double> A((size_t)nyd * nyd), L((size_t)nyd * nyd, 0.0);
        for (int i = 0; i < nyd * nyd; i++) A[i] = T.H1[i];
        for (int m = M - ts; m < M; m++) { const int a = yid(m, 2); A[(size_t)a * nyd + a] += 2.0 * w_terminal; }
        for (int j = 0; j < nyd; j++) {                      // Cholesky A = L L'
            double d = A[(size_t)j * nyd + j];
            for (int k = 0; k < j; k++) d -= L[(size_t)j * nyd + k] * L[(size_t)j * nyd + k];
            d = std::sqrt(d);
            L[(size_t)j * nyd + j] = d;
            for (int i = j + 1; i < nyd; i++) {
                double v = A[(size_t)i * nyd + j];
                for (int k = 0; k < j; k++) v -= L[(size_t)i * nyd + k] * L[(size_t)j * nyd + k];
                L[(size_t)i * nyd + j] = v / d;
            }
        }
        double* Hi = T.Hinv.data() + (size_t)(ts - 1) * nyd * nyd;
        std::vector<double> e(nyd);
        for (int col = 0; col < nyd; col++) {                // solve A x = e_col
            for (int i = 0; i < nyd; i++) {
                double v = (i == col) ? 1.0 : 0.0;
                for (int k = 0; k < i; k++) v -= L[(size_t)i * nyd + k] * e[k];
                e[i] = v / L[(size_t)i * nyd + i];
            }
            for (int i = nyd - 1; i >= 0; i--) {
                double v = e[i];
                for (int k = i + 1; k < nyd; k++) v -= L[(size_t)k * nyd + i] * e[k];
                e[i] = v / L[(size_t)i * nyd + i];
            }
            for (int i = 0; i < nyd; i++) Hi[(size_t)i * nyd + col] = e[i];
        }
    }
    // y0 = -H^-1 g with g = T' grad f(x(y = 0)): linear in the initial-state constants c0,c1,c2 (segment 0,
    // points 0..2 through 2 w_u Q) and in the goal coordinate (terminal term -2 w_T goal on point n of the
    // last ts segments)  ->  one [nyd][4] map per ts
    T.Y0.assign((size_t)M * nyd * 4, 0.0);
    for (int ts = 1; ts <= M; ts++) {
        const double* Hi = T.Hinv.data() + (size_t)(ts - 1) * nyd * nyd;
        for (int j = 0; j < 4; j++) {
            std::vector<double> gx(T.npt, 0.0), gy(nyd, 0.0);
            if (j < 3) { for (int i = 0; i < kP; i++) gx[i] = T.Q2[i * kP + j]; }
            else { for (int m = M - ts; m < M; m++) gx[m * kP + n] += -2.0 * w_terminal; }
            for (int pt = 0; pt < T.npt; pt++)
                for (auto& t : xe[pt].t) gy[t.idx] += t.coef * gx[pt];
            for (int a = 0; a < nyd; a++) {
                double v = 0.0;
                for (int b = 0; b < nyd; b++) v += Hi[(size_t)a * nyd + b] * gy[b];
                T.Y0[((size_t)(ts - 1) * nyd + a) * 4 + j] = -v;
            }
        }
    }

    // pair rows: (axis-local expression) replicated per axis
    struct Row { int fam, pt; Expr e; };
    std::vector<Row> rows;
    auto lin = [&](std::initializer_list<std::pair<int, double>> terms) {
        Expr e;
        for (auto& tm : terms) {
            const Expr& x = xe[tm.first];
            for (auto& t : x.t) expr_add(e, t.idx, tm.second * t.coef);
            for (int i = 0; i < 3; i++) e.cc[i] += tm.second * x.cc[i];
        }
        return e;
    };
    for (int m = 0; m < M; m++)                                             // box rows
        for (int i = 0; i < kP; i++) {
            if (m == 0 && i < 3) continue;
            rows.push_back({0, m * kP + i, lin({{m * kP + i, 1.0}})});
        }
    const double scv = std::pow(dt, -1) * n, sca = std::pow(dt, -2) * n * (n - 1);
    T.scv = scv; T.sca = sca;
    // family-major order inside an axis so that row indices are closed-form (dlsc_qp.cuh: row_* helpers):
    //   box  idx = pt - 3 | vel idx = bv + 5m + i - 2 | acc idx = ba + 4m + i - 1 | comm idx = bc + mi*M - mi(mi-1)/2 + m - mi
    for (int m = 0; m < M; m++)
        for (int i = 0; i < n; i++) {                                       // velocity
            if (m == 0 && (i == 0 || i == 1)) continue;
            rows.push_back({1, m * kP + i, lin({{m * kP + i + 1, scv}, {m * kP + i, -scv}})});
        }
    for (int m = 0; m < M; m++)
        for (int i = 0; i < n - 1; i++) {                                   // acceleration
            if (m == 0 && i == 0) continue;
            rows.push_back({2, m * kP + i, lin({{m * kP + i + 2, sca}, {m * kP + i + 1, -2 * sca}, {m * kP + i, sca}})});
        }
    if (use_comm)
        for (int mi = 0; mi < M; mi++)
            for (int m = mi; m < M; m++) rows.push_back({3, (m * kP + n) | ((mi * kP) << 8), lin({{m * kP + n, 1.0}, {mi * kP + 0, -1.0}})});

    const int npl = (int)rows.size();
    T.np = npl * D;
    T.pr_fam.assign(T.np, 0); T.pr_axis.assign(T.np, 0); T.pr_nnz.assign(T.np, 0); T.pr_pt.assign(T.np, -1);
    T.pr_idx.assign((size_t)T.np * 6, 0); T.pr_val.assign((size_t)T.np * 6, 0.0); T.pr_cc.assign((size_t)T.np * 3, 0.0);
    // row order: family-major inside an axis keeps rows of one kind contiguous per axis
    for (int k = 0; k < D; k++)
        for (int r = 0; r < npl; r++) {
            const int g = k * npl + r;
            T.pr_fam[g] = (uint8_t)rows[r].fam; T.pr_axis[g] = (uint8_t)k; T.pr_pt[g] = (int16_t)(rows[r].pt & 0xff);
            T.pr_desc.push_back((uint32_t)rows[r].fam | ((uint32_t)k << 2) | ((uint32_t)(rows[r].pt & 0xff) << 4) |
                                ((uint32_t)((rows[r].pt >> 8) & 0xff) << 12));
            int nnz = 0;
            for (auto& t : rows[r].e.t) {
                if (t.coef == 0.0) continue;
                T.pr_idx[(size_t)g * 6 + nnz] = (int16_t)(k * nyd + t.idx);
                T.pr_val[(size_t)g * 6 + nnz] = t.coef;
                nnz++;
            }
            T.pr_nnz[g] = (uint8_t)nnz;
            for (int i = 0; i < 3; i++) T.pr_cc[(size_t)g * 3 + i] = rows[r].e.cc[i];
        }

    T.tri_p.assign(T.ntri, 0);
    for (int p = 0; p < T.ny; p++)
        for (int q = 0; q <= p; q++) T.tri_p[p * (p + 1) / 2 + q] = (uint8_t)p;

    // incidence of pair rows per global y, and per global lower-triangular W entry
    std::vector<std::vector<std::pair<int, double>>> yi(T.ny), wi(T.ntri);
    for (int g = 0; g < T.np; g++) {
        const int nnz = T.pr_nnz[g];
        for (int a = 0; a < nnz; a++) {
            const int pa = T.pr_idx[(size_t)g * 6 + a];
            const double va = T.pr_val[(size_t)g * 6 + a];
            yi[pa].push_back({g, va});
            for (int b = 0; b < nnz; b++) {
                const int pb = T.pr_idx[(size_t)g * 6 + b];
                if (pb > pa) continue;
                wi[pa * (pa + 1) / 2 + pb].push_back({g, va * T.pr_val[(size_t)g * 6 + b]});
            }
        }
    }
    auto flatten = [](const std::vector<std::vector<std::pair<int, double>>>& src, std::vector<int>& ptr,
                      std::vector<int16_t>& id, std::vector<double>& cf) {
        ptr.assign(src.size() + 1, 0);
        for (size_t i = 0; i < src.size(); i++) ptr[i + 1] = ptr[i] + (int)src[i].size();
        id.clear(); cf.clear();
        for (auto& l : src) for (auto& e : l) { id.push_back((int16_t)e.first); cf.push_back(e.second); }
        if (id.empty()) { id.push_back(0); cf.push_back(0.0); }
    };
    flatten(yi, T.yi_ptr, T.yi_row, T.yi_coef);
    flatten(wi, T.wi_ptr, T.wi_row, T.wi_coef);

    // incidence of control points per local y and per local (a >= b)
    std::vector<std::vector<std::pair<int, double>>> yp(nyd), wp(T.ntri_local);
    for (int pt = 0; pt < T.npt; pt++)
        for (auto& a : xe[pt].t) {
            yp[a.idx].push_back({pt, a.coef});
            for (auto& b : xe[pt].t) {
                if (b.idx > a.idx) continue;
                wp[a.idx * (a.idx + 1) / 2 + b.idx].push_back({pt, a.coef * b.coef});
            }
        }
    flatten(yp, T.yp_ptr, T.yp_pt, T.yp_coef);
    flatten(wp, T.wp_ptr, T.wp_pt, T.wp_coef);

    // structurally non-zero entries of W = H + G'DG (everything else is only fill-in of the factorisation),
    // one 16-byte header + the H value per entry: {e | k<<16 | kk<<18 | y5flag<<20, wi_off | wi_cnt<<24, wp_off | wp_cnt<<24, seg}
    T.nz_e.clear(); T.nz_hdr.clear(); T.nz_h.clear();
    for (int p = 0; p < T.ny; p++)
        for (int q = 0; q <= p; q++) {
            const int e = p * (p + 1) / 2 + q;
            const int k = p / nyd, a = p % nyd, kk = q / nyd, b = q % nyd;
            const int el = (a >= b) ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a;
            bool nz = !wp[el].empty();
            if (k == kk) nz = nz || a == b || T.H1[(size_t)a * nyd + b] != 0.0 || !wi[e].empty();
            if (nz) {
                T.nz_e.push_back((uint16_t)e);
                const int is_y5 = (k == kk && a == b && (a / 3 == M - 1 || a % 3 == 2)) ? 1 : 0;
                T.nz_hdr.push_back((uint32_t)e | ((uint32_t)k << 16) | ((uint32_t)kk << 18) | ((uint32_t)is_y5 << 20));
                T.nz_hdr.push_back((k == kk) ? ((uint32_t)T.wi_ptr[e] | ((uint32_t)(T.wi_ptr[e + 1] - T.wi_ptr[e]) << 24)) : 0u);
                T.nz_hdr.push_back((uint32_t)T.wp_ptr[el] | ((uint32_t)(T.wp_ptr[el + 1] - T.wp_ptr[el]) << 24));
                T.nz_hdr.push_back((uint32_t)(a / 3));
                T.nz_h.push_back((k == kk) ? T.H1[(size_t)a * nyd + b] : 0.0);
            }
        }
    T.nnzw = (int)T.nz_e.size();
}

}  // namespace dlsc
