/*
 * dlsc_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A serial, plain-C++ restatement of the dlsc_gc_planner replan hot path
 *   horizon shift -> LSC (GJK) -> SFC box expansion -> goal line search -> min-jerk QP
 * written from the reference's behaviour; every function cites the reference file:line
 * it follows (paths relative to the reference repository root).
 *
 * Float32 rounding points of the reference (octomap::point3d = float[3]) are reproduced
 * with the vec3f type below; compile with -ffp-contract=off (see oracle/Makefile).
 *
 * Third-party arithmetic that is NOT under /root/reference and is restated from the
 * published behaviour of the pinned packages (parity unpinned for these details):
 *   - octomath::Vector3 (octomap 1.9.x, Vector3.h): float storage; dot()/norm_sq() are
 *     float expressions returned as double; norm() = sqrt(double); normalize() divides by
 *     (float)norm; distance() subtracts in float and accumulates in double.
 *   - Eigen 3.3/3.4 Matrix3f::inverse() (cofactor formula, Inverse.h) and the 3x3 * 3x1 lazy
 *     product (redux order x0 + (x1 + x2)).
 *   - octomap key arithmetic: key = floor(x * (1/res)) + 32768, coord = (key-32768+0.5)*res.
 *   - dynamicEDT3D: getDistanceAndClosestObstacle returns dist_cells*res and the centre of
 *     the nearest occupied cell; out-of-map -> dist = -1 and the output point untouched.
 *   - IBM CPLEX 22.1.1 (QP and the 1-variable LP): replaced by a dense fp64 primal-dual
 *     interior-point solver / a closed form.  QP objective parity is unpinned; the solver is
 *     cross-checked against HiGHS in tests/ and the composed path against the golden log.
 */
#include "dlsc_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <atomic>
#include <functional>
#include <thread>

#ifndef ORC_QP_TOL_RD
#define ORC_QP_TOL_RD 1e-13
#endif
#ifndef ORC_QP_TOL_MU
#define ORC_QP_TOL_MU 1e-12
#endif

namespace {

constexpr double kEps = 1e-9;        // SP_EPSILON        include/sp_const.hpp:3
constexpr double kEpsF = 1e-5;       // SP_EPSILON_FLOAT  include/sp_const.hpp:4

// ------------------------------------------------------------------------------------------
// vec3f : stand-in for octomath::Vector3 (float storage, see header comment)
// ------------------------------------------------------------------------------------------
struct vec3f {
    float x = 0.f, y = 0.f, z = 0.f;
    vec3f() = default;
    vec3f(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3f(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
    float& operator()(int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator()(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    vec3f operator-(const vec3f& o) const { return {x - o.x, y - o.y, z - o.z}; }
    vec3f operator+(const vec3f& o) const { return {x + o.x, y + o.y, z + o.z}; }
    vec3f operator-() const { return {-x, -y, -z}; }
    vec3f operator*(float s) const { return {x * s, y * s, z * s}; }
    bool operator==(const vec3f& o) const { return x == o.x && y == o.y && z == o.z; }
    double dot(const vec3f& o) const { float r = x * o.x + y * o.y + z * o.z; return (double)r; }
    double norm_sq() const { float r = x * x + y * y + z * z; return (double)r; }
    double norm() const { return std::sqrt(norm_sq()); }
    vec3f normalized() const {
        vec3f r(*this);
        double len = norm();
        if (len > 0) { float f = (float)len; r.x /= f; r.y /= f; r.z /= f; }
        return r;
    }
    double distance(const vec3f& o) const {
        double dx = x - o.x, dy = y - o.y, dz = z - o.z;   // float subtraction, double accumulate
        return std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    vec3f cross(const vec3f& o) const {
        return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x};
    }
    void store(float* p) const { p[0] = x; p[1] = y; p[2] = z; }
};

// util.hpp:131-140
double linf_distance(const vec3f& a, const vec3f& b) {
    double dist = 0;
    for (int k = 0; k < 3; k++) {
        double c = (double)std::fabs(a(k) - b(k));
        if (dist < c) dist = c;
    }
    return dist;
}

// ------------------------------------------------------------------------------------------
// GJK : distance from the origin to the convex hull of a point set.
// Restates openGJK (src/openGJK/openGJK.cpp) for body2 = {origin}; arithmetic order of every
// expression follows the reference so that results agree bit for bit (checked against the
// reference object code in tests/test_oracle_gjk_ref.py).
// ------------------------------------------------------------------------------------------
namespace gjk {

struct Simplex {
    int n;
    double p[4][3];
};

inline double dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void cpy(double* d, const double* s) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; }
inline void cross(const double* a, const double* b, double* c) {   // openGJK.cpp:142-147
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
// openGJK.cpp:138-140
inline double det3(const double* p, const double* q, const double* r) {
    return p[0] * ((q[1] * r[2]) - (r[1] * q[2])) - p[1] * (q[0] * r[2] - r[0] * q[2]) +
           p[2] * (q[0] * r[1] - r[0] * q[1]);
}
// openGJK.cpp:149-161
inline void proj_line(const double* p, const double* q, double* v) {
    double pq[3] = {p[0] - q[0], p[1] - q[1], p[2] - q[2]};
    double t = dot(p, pq) / dot(pq, pq);
    for (int i = 0; i < 3; i++) v[i] = p[i] - pq[i] * t;
}
// openGJK.cpp:163-179
inline void proj_plane(const double* p, const double* q, const double* r, double* v) {
    double pq[3], pr[3], n[3];
    for (int i = 0; i < 3; i++) pq[i] = p[i] - q[i];
    for (int i = 0; i < 3; i++) pr[i] = p[i] - r[i];
    cross(pq, pr, n);
    double t = dot(n, p) / dot(n, n);
    for (int i = 0; i < 3; i++) v[i] = n[i] * t;
}
// openGJK.cpp:181-193 : origin projects onto the open edge p->q ?
inline int hff1(const double* p, const double* q) {
    double t = 0;
    for (int i = 0; i < 3; i++) t += (p[i] * p[i] - p[i] * q[i]);
    return t > 0 ? 1 : 0;
}
// openGJK.cpp:195-218 : 1 => r is to be discarded
inline int hff2(const double* p, const double* q, const double* r) {
    double pq[3], pr[3], nt[3], n[3];
    for (int i = 0; i < 3; i++) pq[i] = q[i] - p[i];
    for (int i = 0; i < 3; i++) pr[i] = r[i] - p[i];
    cross(pq, pr, nt);
    cross(pq, nt, n);
    double t = 0;
    for (int i = 0; i < 3; i++) t = t + (p[i] * n[i]);
    return t < 0 ? 1 : 0;
}
// openGJK.cpp:220-241
inline int hff3(const double* p, const double* q, const double* r) {
    double pq[3], pr[3], n[3];
    for (int i = 0; i < 3; i++) pq[i] = q[i] - p[i];
    for (int i = 0; i < 3; i++) pr[i] = r[i] - p[i];
    cross(pq, pr, n);
    double t = 0;
    for (int i = 0; i < 3; i++) t = t + (p[i] * n[i]);
    return t > 0 ? 0 : 1;
}

// openGJK.cpp:243-256
void sub1d(Simplex& s, double* v) {
    const double* a = s.p[1];
    const double* b = s.p[0];
    if (hff1(a, b)) {
        proj_line(a, b, v);
    } else {
        cpy(v, a);
        s.n = 1;
        cpy(s.p[0], s.p[1]);
    }
}

// openGJK.cpp:259-313
void sub2d(Simplex& s, double* v) {
    const double* a = s.p[2];
    const double* b = s.p[1];
    const double* c = s.p[0];
    const int e_ab = hff1(a, b);
    const int e_ac = hff1(a, c);
    const int f_bc = !hff2(a, b, c);
    const int f_cb = !hff2(a, c, b);
    enum { FACE, EDGE_AB, EDGE_AC, VERT } r;
    if (e_ab) {
        if (f_bc) r = (e_ac && !f_cb) ? EDGE_AC : FACE;
        else r = EDGE_AB;
    } else if (e_ac) {
        r = f_cb ? FACE : EDGE_AC;
    } else {
        r = VERT;
    }
    switch (r) {
        case FACE: proj_plane(a, b, c, v); break;
        case EDGE_AC: proj_line(a, c, v); s.n = 2; cpy(s.p[1], s.p[2]); break;     // keeps {c, a}
        case EDGE_AB: proj_line(a, b, v); s.n = 2; cpy(s.p[0], s.p[2]); break;     // keeps {a, b}
        case VERT: cpy(v, a); s.n = 1; cpy(s.p[0], s.p[2]); break;
    }
}

// openGJK.cpp:315-631
void sub3d(Simplex& s, double* v) {
    double a[3], q[3][3];                 // a = newest vertex; q[2]=s2, q[1]=s3, q[0]=s4 (slot index)
    cpy(a, s.p[3]);
    for (int t = 0; t < 3; t++) cpy(q[t], s.p[t]);
    double e2[3], e3[3], e4[3];
    for (int t = 0; t < 3; t++) { e2[t] = q[2][t] - a[t]; e3[t] = q[1][t] - a[t]; e4[t] = q[0][t] - a[t]; }

    int edge[3];                          // edge[slot] : hff1(a, q[slot])
    edge[2] = hff1(a, q[2]);
    edge[1] = hff1(a, q[1]);
    edge[0] = hff1(a, q[0]);
    const int n_edge = edge[2] + edge[1] + edge[0];
    if (n_edge == 0) {                    // vertex region
        cpy(v, a); s.n = 1; cpy(s.p[0], a);
        return;
    }
    const int sss = det3(e3, e4, e2) > 0 ? 0 : 1;
    int t2 = hff3(a, q[1], q[0]) - sss; t2 *= t2;
    int t3 = hff3(a, q[0], q[2]) - sss; t3 *= t3;
    int t4 = hff3(a, q[2], q[1]) - sss; t4 *= t4;

    // simplex rebuild helpers (slot order matters for the next iteration)
    auto keep_face = [&](const double* lo, const double* mid) {      // {lo, mid, a}
        double l[3], m[3]; cpy(l, lo); cpy(m, mid);
        s.n = 3; cpy(s.p[2], a); cpy(s.p[1], m); cpy(s.p[0], l);
    };
    auto keep_edge = [&](const double* lo) {                        // {lo, a}
        double l[3]; cpy(l, lo);
        s.n = 2; cpy(s.p[1], a); cpy(s.p[0], l);
    };

    const int n_face = t2 + t3 + t4;
    if (n_face == 3) {                    // origin enclosed
        v[0] = v[1] = v[2] = 0; s.n = 4;
        return;
    }
    if (n_face == 2) {                    // exactly one face sees the origin: drop the opposite vertex
        s.n = 3;
        if (!t2) { cpy(s.p[2], a); }                                        // {s4, s3, a}
        else if (!t3) { cpy(s.p[1], q[2]); cpy(s.p[2], a); }                // {s4, s2, a}
        else { cpy(s.p[0], q[1]); cpy(s.p[1], q[2]); cpy(s.p[2], a); }      // {s3, s2, a}
        sub2d(s, v);
        return;
    }
    int i, j, k;
    if (n_face == 1) {
        s.n = 3;
        if (t2) { k = 2; i = 1; j = 0; }
        else if (t3) { k = 1; i = 0; j = 2; }
        else { k = 0; i = 2; j = 1; }
        const double* si = q[i]; const double* sj = q[j]; const double* sk = q[k];
        if (n_edge == 1) {
            if (edge[k]) {
                if (!hff2(a, sk, si)) { keep_face(sk, si); proj_plane(a, si, sk, v); }
                else if (!hff2(a, sk, sj)) { keep_face(sk, sj); proj_plane(a, sj, sk, v); }
                else { keep_edge(sk); proj_line(a, sk, v); }
            } else if (edge[i]) {
                if (!hff2(a, si, sk)) { keep_face(sk, si); proj_plane(a, si, sk, v); }
                else { keep_edge(si); proj_line(a, si, v); }
            } else {
                if (!hff2(a, sj, sk)) { keep_face(sk, sj); proj_plane(a, sj, sk, v); }
                else { keep_edge(sj); proj_line(a, sj, v); }
            }
        } else if (n_edge == 2) {
            if (edge[i]) {
                if (!hff2(a, sk, si)) {
                    if (!hff2(a, si, sk)) { keep_face(sk, si); proj_plane(a, si, sk, v); }
                    else { keep_edge(sk); proj_line(a, sk, v); }
                } else {
                    if (!hff2(a, sk, sj)) { keep_face(sk, sj); proj_plane(a, sj, sk, v); }
                    else { keep_edge(sk); proj_line(a, sk, v); }
                }
            } else if (edge[j]) {
                if (!hff2(a, sk, sj)) {
                    if (!hff2(a, sj, sk)) { keep_face(sk, sj); proj_plane(a, sj, sk, v); }
                    else { keep_edge(sj); proj_line(a, sj, v); }
                } else {
                    if (!hff2(a, sk, si)) { keep_face(sk, si); proj_plane(a, si, sk, v); }
                    else { keep_edge(sk); proj_line(a, sk, v); }
                }
            }
            // else: reference leaves {s4,s3,s2} and v untouched (openGJK.cpp:497-499)
        } else {   // n_edge == 3
            const int d_ik = hff2(a, si, sk), d_jk = hff2(a, sj, sk);
            const int d_ki = hff2(a, sk, si), d_kj = hff2(a, sk, sj);
            if (d_ki == 1 && d_kj == 1) { keep_edge(sk); proj_line(a, sk, v); }
            else if (d_ki) {
                if (d_jk) { keep_edge(sj); proj_line(a, sj, v); }
                else { keep_face(sk, sj); proj_plane(a, sk, sj, v); }
            } else {
                if (d_ik) { keep_edge(si); proj_line(a, si, v); }
                else { keep_face(sk, si); proj_plane(a, sk, si, v); }
            }
        }
        return;
    }
    // n_face == 0 : origin outside all three faces through a
    if (n_edge == 1) {
        if (edge[1]) { k = 2; i = 1; j = 0; }
        else if (edge[0]) { k = 1; i = 0; j = 2; }
        else { k = 0; i = 2; j = 1; }
        const double* si = q[i]; const double* sj = q[j]; const double* sk = q[k];
        if (!hff2(a, si, sj)) { keep_face(sj, si); proj_plane(a, si, sj, v); }
        else if (!hff2(a, si, sk)) { keep_face(sk, si); proj_plane(a, si, sk, v); }
        else { keep_edge(si); proj_line(a, si, v); }
    } else if (n_edge == 2) {
        s.n = 3;
        if (!edge[1]) { k = 2; i = 1; j = 0; }
        else if (!edge[0]) { k = 1; i = 0; j = 2; }
        else { k = 0; i = 2; j = 1; }
        const double* si = q[i]; const double* sj = q[j]; const double* sk = q[k];
        if (!hff2(a, sj, sk)) {
            if (!hff2(a, sk, sj)) { keep_face(sk, sj); proj_plane(a, sj, sk, v); }
            else if (!hff2(a, sk, si)) { keep_face(sk, si); proj_plane(a, sk, si, v); }
            else { keep_edge(sk); proj_line(a, sk, v); }
        } else if (!hff2(a, sj, si)) { keep_face(sj, si); proj_plane(a, si, sj, v); }
        else { keep_edge(sj); proj_line(a, sj, v); }
    }
    // n_edge == 3 with n_face == 0: reference does nothing (n stays 4) -- openGJK.cpp:545-626
}

// openGJK.cpp:674-780 with bd2 = single point at the origin (geometry.hpp:289-298)
double hull_origin(const double (*c)[3], int np, double* v, int* iters, int* simplex_n) {
    const double eps_rel = 1e-10, eps_tot = 1e-12;
    const double eps_rel2 = eps_rel * eps_rel;
    Simplex s;
    double sup[3], w[3], vm[3];
    double nwmax = 0;
    int k = 0;
    cpy(v, c[0]);
    s.n = 1; cpy(s.p[0], v);
    cpy(sup, c[0]);
    do {
        k++;
        for (int t = 0; t < 3; t++) vm[t] = -v[t];
        double maxs = dot(sup, vm);                // openGJK.cpp:633-655
        int better = -1;
        for (int i = 0; i < np; i++) {
            double sv = dot(c[i], vm);
            if (sv > maxs) { maxs = sv; better = i; }
        }
        if (better != -1) cpy(sup, c[better]);
        for (int t = 0; t < 3; t++) w[t] = sup[t] - 0.0;
        const double vv = dot(v, v);
        const double ex = vv - dot(v, w);
        if (ex <= eps_rel * vv || ex < eps_tot) break;
        if (vv < eps_rel2) break;
        cpy(s.p[s.n], w);
        s.n++;
        switch (s.n) {
            case 4: sub3d(s, v); break;
            case 3: sub2d(s, v); break;
            case 2: sub1d(s, v); break;
            default: break;
        }
        for (int jj = 0; jj < s.n; jj++) {
            double tn = dot(s.p[jj], s.p[jj]);
            if (tn > nwmax) nwmax = tn;
        }
        if (dot(v, v) <= eps_tot * eps_tot * nwmax) break;
    } while (s.n != 4 && k != 25);
    if (iters) *iters = k;
    if (simplex_n) *simplex_n = s.n;
    return std::sqrt(dot(v, v));
}
}  // namespace gjk

// ------------------------------------------------------------------------------------------
// line-segment geometry (include/geometry.hpp)
// ------------------------------------------------------------------------------------------
struct Closest { double dist; vec3f p1, p2; };

// geometry.hpp:77-112
Closest closest_point_segment(const vec3f& point, const vec3f& s0, const vec3f& s1) {
    vec3f a = s0 - point, b = s1 - point, rel;
    double dmin;
    if (a == b) {
        dmin = a.norm(); rel = a;
    } else {
        dmin = a.norm(); rel = a;
        double dist = b.norm();
        if (dmin > dist) { dmin = dist; rel = b; }
        vec3f nl = (b - a).normalized();
        vec3f c = a - nl * (float)a.dot(nl);
        dist = c.norm();
        if ((c - a).dot(c - b) < 0 && dmin > dist) { dmin = dist; rel = c; }
    }
    Closest r; r.dist = dmin; r.p1 = point; r.p2 = rel + point;
    return r;
}

// Eigen::Matrix3f::inverse() * b, restated (see header comment)
void solve3f(const float m[3][3], const float b[3], float out[3]) {
    auto cof = [&](int i, int j) -> float {
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
    };
    float c0[3] = {cof(0, 0), cof(1, 0), cof(2, 0)};
    float det = c0[0] * m[0][0] + (c0[1] * m[1][0] + c0[2] * m[2][0]);
    float invdet = 1.0f / det;
    float inv[3][3];
    for (int c = 0; c < 3; c++) inv[0][c] = c0[c] * invdet;
    inv[1][0] = cof(0, 1) * invdet; inv[1][1] = cof(1, 1) * invdet; inv[1][2] = cof(2, 1) * invdet;
    inv[2][0] = cof(0, 2) * invdet; inv[2][1] = cof(1, 2) * invdet; inv[2][2] = cof(2, 2) * invdet;
    for (int r = 0; r < 3; r++) out[r] = inv[r][0] * b[0] + (inv[r][1] * b[1] + inv[r][2] * b[2]);
}

// geometry.hpp:139-182 (inputs are non-degenerate; checked by the caller)
Closest closest_lines(const vec3f& a0, const vec3f& a1, const vec3f& b0, const vec3f& b1) {
    Closest r;
    vec3f n1 = (a1 - a0).normalized();
    vec3f n2 = (b1 - b0).normalized();
    if (n1.distance(n2) < kEpsF || n1.distance(-n2) < kEpsF) {
        vec3f delta = b0 - a0;
        delta = delta - n1 * (float)(delta.dot(n1));
        r.dist = delta.norm(); r.p1 = a0; r.p2 = a0 + delta;
    } else {
        vec3f delta = b0 - a0;
        vec3f n3 = (n2.cross(n1)).normalized();
        float A[3][3] = {{n1.x, -n2.x, n3.x}, {n1.y, -n2.y, n3.y}, {n1.z, -n2.z, n3.z}};
        float bb[3] = {delta.x, delta.y, delta.z}, al[3];
        solve3f(A, bb, al);
        r.dist = (double)std::fabs(al[2]);
        r.p1 = a0 + n1 * al[0];
        r.p2 = b0 + n2 * al[1];
    }
    return r;
}

// geometry.hpp:184-274
Closest closest_segments(const vec3f& a0, const vec3f& a1, const vec3f& b0, const vec3f& b1) {
    Closest cp;
    if (a0.distance(a1) < kEpsF) {
        cp = closest_point_segment(a0, b0, b1);
    } else if (b0.distance(b1) < kEpsF) {
        cp = closest_point_segment(b0, a0, a1);
        std::swap(cp.p1, cp.p2);
    } else {
        vec3f v1 = a1 - a0, v2 = b1 - b0;
        double l1 = v1.norm(), l2 = v2.norm();
        vec3f n1 = v1 * (float)(1 / l1), n2 = v2 * (float)(1 / l2);
        if ((n1.cross(n2)).norm() < kEpsF) {
            double bmin = (b0 - a0).dot(n1), bmax = (b1 - a0).dot(n1);
            vec3f pmin = b0, pmax = b1;
            if (bmax < bmin) { std::swap(bmin, bmax); std::swap(pmin, pmax); }
            vec3f delta = b0 - a0;
            delta = delta - n1 * (float)(delta.dot(n1));
            if (l1 < bmin) { cp.p1 = a1; cp.p2 = pmin; }
            else if (bmax < 0) { cp.p1 = a0; cp.p2 = pmax; }
            else if (bmin < 0) { cp.p1 = a0; cp.p2 = a0 + delta; }
            else { cp.p1 = pmin - delta; cp.p2 = pmin; }
            cp.dist = cp.p1.distance(cp.p2);
        } else {
            cp = closest_lines(a0, a1, b0, b1);
            double al1 = (cp.p1 - a0).dot(n1) / l1;
            double al2 = (cp.p2 - b0).dot(n2) / l2;
            if (al1 < 0) cp.p1 = a0; else if (al1 > 1) cp.p1 = a1;
            if (al2 < 0) cp.p2 = b0; else if (al2 > 1) cp.p2 = b1;
            if (al1 < 0 || al1 > 1) {
                double d = n2.dot(cp.p1 - b0);
                if (d < 0) d = 0; else if (d > l2) d = l2;
                cp.p2 = b0 + n2 * (float)d;
            }
            if (al2 < 0 || al2 > 1) {
                double d = n1.dot(cp.p2 - a0);
                if (d < 0) d = 0; else if (d > l1) d = l1;
                cp.p1 = a0 + n1 * (float)d;
            }
            cp.dist = cp.p1.distance(cp.p2);
        }
    }
    return cp;
}

// ------------------------------------------------------------------------------------------
// trajectory helpers
// ------------------------------------------------------------------------------------------
inline int P_of(const orc_params* p) { return p->n + 1; }
inline size_t traj_len(const orc_params* p) { return (size_t)p->M * P_of(p) * 3; }

// trajectory.cpp:79-91
void const_vel_traj(const orc_params* p, const vec3f& pos, const vec3f& vel, float* out) {
    double time = 0;
    const int P = P_of(p);
    for (int m = 0; m < p->M; m++)
        for (int i = 0; i < P; i++) {
            vec3f q = pos + vel * (float)time;
            q.store(out + ((size_t)m * P + i) * 3);
            time += p->dt / p->n;
        }
}

// traj_planner.cpp:304-314 / 412-421
void shift_traj(const orc_params* p, const float* prev, float* out) {
    const int P = P_of(p), M = p->M;
    for (int m = 0; m < M; m++)
        for (int i = 0; i < P; i++) {
            const float* src = (m == M - 1) ? prev + ((size_t)(M - 1) * P + p->n) * 3
                                            : prev + ((size_t)(m + 1) * P + i) * 3;
            float* dst = out + ((size_t)m * P + i) * 3;
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
        }
}

// ------------------------------------------------------------------------------------------
// LSC for one (agent, neighbour) pair -- traj_planner.cpp:603-666, 1102-1127, 1150-1161
// ------------------------------------------------------------------------------------------
void lsc_pair(const orc_params* p, const float* init_traj, const float* pred_traj, const vec3f& goal_a,
              const vec3f& goal_j, double r_a, double dw_a, double r_j_in, double dw_j_in,
              float* normal /*[M][3]*/, float* anchor /*[M][P][3]*/, double* d /*[M][P]*/,
              int64_t* hist) {
    const int P = P_of(p), M = p->M;
    const double r_j = (double)(float)r_j_in;      // agent_manager.cpp:256-257
    const double dw_j = (double)(float)dw_j_in;
    const double collision_dist = r_j + r_a;                                   // :605
    const double downwash = (dw_a * r_a + dw_j * r_j) / (r_a + r_j);           // :1153-1154
    const float dwf = (float)downwash;                                         // trajectory.cpp:214
    auto tr = [&](const float* q) { vec3f v(q); v.z = v.z / dwf; return v; };  // trajectory transform
    auto trp = [&](const vec3f& q) { vec3f v = q; v.z = (float)((double)v.z / downwash); return v; };  // :1183-1187

    for (int m = 0; m < M; m++) {
        if (m < M - 1) {
            vec3f rel[8];
            double c[8][3];
            for (int i = 0; i < P; i++) {
                rel[i] = tr(init_traj + ((size_t)m * P + i) * 3) - tr(pred_traj + ((size_t)m * P + i) * 3);
                c[i][0] = rel[i].x; c[i][1] = rel[i].y; c[i][2] = rel[i].z;      // util.hpp:113-125
            }
            double v[3];
            int it = 0, sn = 0;
            gjk::hull_origin(c, P, v, &it, &sn);
            if (hist) hist[std::min(it, 31)]++;
            vec3f cp2 = vec3f(0, 0, 0) + vec3f((float)v[0], (float)v[1], (float)v[2]);   // geometry.hpp:302
            vec3f nt = cp2.normalized();                                                  // :1118
            vec3f nrm(nt.x, nt.y, (float)((double)nt.z / downwash));                      // :630-632
            nrm.store(normal + (size_t)m * 3);
            for (int i = 0; i < P; i++) {
                vec3f diff = tr(init_traj + ((size_t)m * P + i) * 3) - tr(pred_traj + ((size_t)m * P + i) * 3);
                d[(size_t)m * P + i] = 0.5 * (collision_dist + diff.dot(nt));             // :636-637
                const float* a = pred_traj + ((size_t)m * P + i) * 3;                     // anchor untransformed :638
                float* dst = anchor + ((size_t)m * P + i) * 3;
                dst[0] = a[0]; dst[1] = a[1]; dst[2] = a[2];
            }
        } else {
            vec3f o_last = tr(pred_traj + ((size_t)(M - 1) * P + p->n) * 3);
            vec3f a_last = tr(init_traj + ((size_t)(M - 1) * P + p->n) * 3);
            vec3f og = trp(goal_j), ag = trp(goal_a);                                     // :611-612
            Closest cp = closest_segments(o_last, og, a_last, ag);                        // :642-644
            vec3f nt = (cp.p2 - cp.p1).normalized();
            double dd = 0.5 * (collision_dist + cp.dist);                                  // :650
            vec3f nrm(nt.x, nt.y, (float)((double)nt.z / downwash));
            vec3f oc = cp.p1;
            oc.z = (float)((double)oc.z * downwash);                                       // :657
            nrm.store(normal + (size_t)m * 3);
            for (int i = 0; i < P; i++) {
                d[(size_t)m * P + i] = dd;
                oc.store(anchor + ((size_t)m * P + i) * 3);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Dynamic (non-agent) obstacles -- traj_planner.cpp:338-368 (size prediction), :617-627 + :1129-1148 +
// :1080-1100 (LSC), :708-735 (waypoint trap); geometry.hpp:115-137; obstacle.hpp:26-36
// ------------------------------------------------------------------------------------------
// obs_pred_sizes: radius + the Bernstein control points of 1/2 a_max t^2 over the uncertainty horizon, constant after it.
// coef * B_inv (polynomial.hpp:280-293) restated with the closed form B_inv(k, i) = C(i,k)/C(n,k) (Eigen's numeric
// inverse differs from it by rounding only: parity unpinned at the 1e-16 level).
void obstacle_sizes(const orc_params* p, int size_prediction, double uncertainty_horizon, double radius, double max_acc,
                    double* out /*[M][P]*/) {
    const int M = p->M, P = P_of(p), n = p->n;
    const int Mu = std::min((int)((uncertainty_horizon + kEps) / p->dt), M);               // :339
    for (int m = 0; m < M; m++)
        for (int i = 0; i < P; i++) {
            double v = radius;                                                              // :364 planConstVelTraj(radius, 0)
            if (size_prediction) {
                if (m < Mu) {
                    const double a0 = 0.5 * max_acc * std::pow(m * p->dt, 2);               // :349-351
                    const double a1 = max_acc * m * p->dt * p->dt;
                    const double a2 = 0.5 * max_acc * std::pow(p->dt, 2);
                    const double cp = a0 + a1 * ((double)i / n) + a2 * ((double)(i * (i - 1)) / (n * (n - 1)));
                    v = radius + cp;                                                        // :355
                } else {
                    v = radius + 0.5 * max_acc * std::pow(Mu * p->dt, 2);                   // :360-361
                }
            }
            out[(size_t)m * P + i] = v;
        }
}

// normalVectorBetweenLines(line_obs, line_agent) :1080-1100 on closestPointsBetweenLinePaths geometry.hpp:115-137
vec3f normal_between_line_paths(const vec3f& o0, const vec3f& o1, const vec3f& a0, const vec3f& a1) {
    const vec3f r0 = a0 - o0, r1 = a1 - o1;                                                // rel_path = line2 - line1
    const Closest rc = closest_point_segment(vec3f(0, 0, 0), r0, r1);
    const double len = r0.distance(r1);
    double alpha = 0;
    if (len > 0) alpha = (rc.p2 - r0).norm() / len;
    const vec3f c1 = o0 + (o1 - o0) * (float)alpha;
    const vec3f c2 = a0 + (a1 - a0) * (float)alpha;
    vec3f nv = (c2 - c1).normalized();
    if (nv.norm() == 0) {                                                                   // heuristic :1089-1098
        const vec3f a = a0 - o0, b = a1 - o1;
        if (a.norm() == 0 && b.norm() == 0) nv = vec3f(1, 0, 0);
        else nv = (b - a).cross(vec3f(0, 0, 1));
    }
    return nv;
}

void lsc_dynamic(const orc_params* p, const float* init_traj, const float* obs_traj, const double* size, double r_a,
                 double r_o, double dw_o, float* normal /*[M][3]*/, float* anchor /*[M][P][3]*/, double* d /*[M][P]*/) {
    const int P = P_of(p), M = p->M, n = p->n;
    const double downwash = (r_a + dw_o * r_o) / (r_a + r_o);                               // :1156-1157
    for (int m = 0; m < M; m++) {
        // the lines are NOT downwash-transformed (normalVectorDynamicObs :1144-1147); only z is divided afterwards
        const vec3f nt = normal_between_line_paths(vec3f(obs_traj + ((size_t)m * P) * 3), vec3f(obs_traj + ((size_t)m * P + n) * 3),
                                                   vec3f(init_traj + ((size_t)m * P) * 3), vec3f(init_traj + ((size_t)m * P + n) * 3));
        vec3f nrm(nt.x, nt.y, (float)((double)nt.z / downwash));                            // :618-620
        nrm.store(normal + (size_t)m * 3);
        for (int i = 0; i < P; i++) {
            d[(size_t)m * P + i] = size[(size_t)m * P + i] + r_a;                            // :624
            const float* a = obs_traj + ((size_t)m * P + i) * 3;                             // :625
            float* dst = anchor + ((size_t)m * P + i) * 3;
            dst[0] = a[0]; dst[1] = a[1]; dst[2] = a[2];
        }
    }
}

// Obstacle::isCollided obstacle.hpp:26-36
bool obstacle_collides(const vec3f& opos, const vec3f& ovel, double oradius, double omax_acc, const vec3f& point,
                       double agent_radius, double horizon, double uncertainty_horizon) {
    for (double t = 0; t <= horizon; t += std::min(0.1 * horizon, 0.1)) {
        const vec3f q = opos + ovel * (float)t;
        const double tm = std::min(t, uncertainty_horizon);
        if (q.distance(point) < agent_radius + oradius + 0.5 * omax_acc * tm * tm) return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------
// SFC
// ------------------------------------------------------------------------------------------
struct Box { vec3f lo, hi; };

inline bool point_in_box(const Box& b, const vec3f& q) {    // collision_constraints.cpp:109-116
    return q.x > b.lo.x - kEpsF && q.y > b.lo.y - kEpsF && q.z > b.lo.z - kEpsF &&
           q.x < b.hi.x + kEpsF && q.y < b.hi.y + kEpsF && q.z < b.hi.z + kEpsF;
}
inline bool box_includes(const Box& b, const Box& o) { return point_in_box(b, o.lo) && point_in_box(b, o.hi); }   // :204-206
inline Box box_intersection(const Box& a, const Box& b) {   // :217-224
    Box r;
    for (int i = 0; i < 3; i++) { r.lo(i) = std::max(a.lo(i), b.lo(i)); r.hi(i) = std::min(a.hi(i), b.hi(i)); }
    return r;
}
// :163-178
bool superset_of_hull(const Box& b, const vec3f* pts, int np) {
    for (int i = 0; i < 3; i++) {
        float mn = pts[0](i), mx = pts[0](i);
        for (int k = 1; k < np; k++) { mn = std::min(mn, pts[k](i)); mx = std::max(mx, pts[k](i)); }
        if (mn < b.lo(i) - kEpsF || mx > b.hi(i) + kEpsF) return false;
    }
    return true;
}

struct EdtView {
    const orc_edt* g;
    double res, inv_res;
    // DynamicEDTOctomap::getDistanceAndClosestObstacle [ext], see header comment
    void lookup(const vec3f& q, float& dist, vec3f& closest) const {
        int c[3];
        for (int k = 0; k < 3; k++) c[k] = (int)std::floor(inv_res * (double)q(k)) - g->min_key[k];
        if (c[0] >= 0 && c[0] < g->dims[0] && c[1] >= 0 && c[1] < g->dims[1] && c[2] >= 0 && c[2] < g->dims[2]) {
            size_t idx = ((size_t)c[0] * g->dims[1] + c[1]) * g->dims[2] + c[2];
            dist = g->dist[idx];
            const int32_t* o = g->obst + idx * 3;
            if (o[0] >= 0) {
                for (int k = 0; k < 3; k++) closest(k) = (float)(((double)(o[k] + g->min_key[k]) + 0.5) * res);
            }
        } else {
            dist = -1.0f;
        }
    }
};

struct SfcCtx {
    const orc_params* p;
    EdtView edt;
    int64_t lookups = 0;
};

// collision_constraints.cpp:862-892
bool obstacle_in_box(SfcCtx& c, const Box& b, double margin) {
    const double res = c.p->world_res;
    vec3f delta((float)(0.5 * res), (float)(0.5 * res), (float)(0.5 * res));
    int mi[3];
    for (int i = 0; i < 3; i++) mi[i] = (int)std::floor(((b.hi(i) - b.lo(i)) + kEpsF) / res) + 1;
    for (int ix = 0; ix < mi[0]; ix++)
        for (int iy = 0; iy < mi[1]; iy++)
            for (int iz = 0; iz < mi[2]; iz++) {
                vec3f q((float)(b.lo.x + ix * res), (float)(b.lo.y + iy * res), (float)(b.lo.z + iz * res));
                float dist;
                vec3f cl;
                c.edt.lookup(q, dist, cl);
                c.lookups++;
                Box cell{cl - delta, cl + delta};
                vec3f cq = q;                                               // Box::closestPoint :226-237
                for (int k = 0; k < 3; k++) {
                    if (q(k) < cell.lo(k)) cq(k) = cell.lo(k);
                    else if (q(k) > cell.hi(k)) cq(k) = cell.hi(k);
                }
                double dobs = linf_distance(cq, q);
                if (dist < 1 && dobs < margin + kEpsF) return true;
            }
    return false;
}

// collision_constraints.cpp:894-901 (margin = 0 at the only call site :1044)
bool box_in_boundary(const orc_params* p, const Box& b) {
    for (int k = 0; k < 3; k++) {
        if (!(b.lo(k) > (float)p->world_min[k] + 0.0 - kEpsF)) return false;
        if (!(b.hi(k) < (float)p->world_max[k] - 0.0 + kEpsF)) return false;
    }
    return true;
}

// collision_constraints.cpp:1023-1093
bool expand_incrementally(SfcCtx& c, const Box& init, double margin, double max_vel, Box& out) {
    const orc_params* p = c.p;
    const double res = p->world_res;
    if (obstacle_in_box(c, init, margin)) return false;
    int axes[6] = {0, 1, 2, 3, 4, 5};
    int n_axes = 6;
    int iters[6] = {0, 0, 0, 0, 0, 0};
    const int max_iter = (int)std::round(std::max(2 * p->grid_res, max_vel * p->dt) / res) + 1;
    int i = -1;
    Box sfc = init, cand, upd;
    while (n_axes > 0) {
        cand = sfc; upd = sfc;
        while (box_in_boundary(p, upd) && !obstacle_in_box(c, upd, margin)) {
            i++;
            if (i >= n_axes) i = 0;
            int ax = axes[i];
            sfc = cand; upd = cand;
            if (ax < 3) {
                upd.hi(ax) = cand.lo(ax);
                cand.lo(ax) = (float)(cand.lo(ax) - res);
                upd.lo(ax) = cand.lo(ax);
            } else {
                upd.lo(ax - 3) = cand.hi(ax - 3);
                cand.hi(ax - 3) = (float)(cand.hi(ax - 3) + res);
                upd.hi(ax - 3) = cand.hi(ax - 3);
            }
            iters[ax]++;
            if (iters[ax] > max_iter) break;
        }
        if (i < 0) return false;   // reference would erase begin()-1 (UB) when the start box is outside the world
        for (int t = i; t < n_axes - 1; t++) axes[t] = axes[t + 1];
        n_axes--;
        if (i > 0) i--; else i = n_axes - 1;
    }
    double delta = margin - ((int)(margin / res) * res);     // :1081
    for (int k = 0; k < 3; k++) {
        if (sfc.lo(k) > (float)p->world_min[k] + kEpsF) sfc.lo(k) = (float)(sfc.lo(k) - delta);
        if (sfc.hi(k) < (float)p->world_max[k] - kEpsF) sfc.hi(k) = (float)(sfc.hi(k) + delta);
    }
    out = sfc;
    return true;
}

Box hull_aabb(const vec3f* pts, int np) {
    Box b{pts[0], pts[0]};
    for (int t = 0; t < np; t++)
        for (int k = 0; k < 3; k++) {
            if (pts[t](k) < b.lo(k)) b.lo(k) = pts[t](k);
            if (pts[t](k) > b.hi(k)) b.hi(k) = pts[t](k);
        }
    return b;
}

// collision_constraints.cpp:781-815
bool expand_from_hull(SfcCtx& c, const vec3f* pts, int np, double margin, double max_vel, Box& out) {
    const double res = c.p->world_res;
    Box b = hull_aabb(pts, np);
    for (int k = 0; k < 3; k++) {
        b.lo(k) = (float)(std::round(b.lo(k) / res) * res);
        b.hi(k) = (float)(std::round(b.hi(k) / res) * res);
    }
    bool ok = expand_incrementally(c, b, margin, max_vel, out);
    if (ok && !superset_of_hull(out, pts, np)) ok = false;
    return ok;
}

// collision_constraints.cpp:817-860
bool expand_from_hull_prev(SfcCtx& c, const vec3f* pts, int np, const Box& prev, double margin,
                           double max_vel, Box& out) {
    const double res = c.p->world_res;
    Box b = hull_aabb(pts, np);
    for (int k = 0; k < 3; k++) {
        b.lo(k) = (float)(std::floor(b.lo(k) / res) * res);
        b.hi(k) = (float)(std::ceil(b.hi(k) / res) * res);
    }
    if (!box_includes(prev, b)) {
        b = box_intersection(prev, b);
        for (int k = 0; k < 3; k++) {
            b.lo(k) = (float)(std::ceil((b.lo(k) - kEpsF) / res) * res);
            b.hi(k) = (float)(std::floor((b.hi(k) + kEpsF) / res) * res);
        }
    }
    return expand_incrementally(c, b, margin, max_vel, out);
}

inline Box load_box(const float* s) { return Box{vec3f(s), vec3f(s + 3)}; }
inline void store_box(const Box& b, float* s) { b.lo.store(s); b.hi.store(s + 3); }

// One agent's SFC update for this replan.
// init: collision_constraints.cpp:435-452 ; else :502-536 (traj_planner.cpp:692-706)
int sfc_agent(const orc_params* p, const orc_edt* edt, bool init, const vec3f& pos, const float* init_traj,
              const vec3f& goal, const vec3f& wp, double radius, double max_vel, float* sfc, int64_t* lookups) {
    SfcCtx c{p, EdtView{edt, edt->res, 1.0 / edt->res}};
    const int M = p->M, P = P_of(p);
    const double res = p->world_res;
    int status = ORC_OK;
    if (init) {
        Box b;
        for (int k = 0; k < 3; k++) {
            b.lo(k) = (float)(std::floor(pos(k) / res) * res);
            b.hi(k) = (float)(std::ceil(pos(k) / res) * res);
        }
        Box out;
        if (!expand_incrementally(c, b, radius, max_vel, out)) {
            status = ORC_SFC_INIT_FAILED;
            out = b;
        }
        for (int m = 0; m < M; m++) store_box(out, sfc + (size_t)m * 6);
    } else {
        for (int m = 0; m < M - 1; m++) std::memcpy(sfc + (size_t)m * 6, sfc + (size_t)(m + 1) * 6, 6 * sizeof(float));
        for (int m = 0; m < M - 2; m++) {                                           // :511-516
            vec3f cps[8];
            for (int i = 0; i < P; i++) cps[i] = vec3f(init_traj + ((size_t)m * P + i) * 3);
            Box nxt = load_box(sfc + (size_t)(m + 1) * 6);
            if (superset_of_hull(nxt, cps, P)) store_box(nxt, sfc + (size_t)m * 6);
        }
        vec3f hull[3] = {vec3f(init_traj + ((size_t)(M - 1) * P + p->n) * 3), goal, wp};
        Box upd;
        bool ok = expand_from_hull(c, hull, 3, radius, max_vel, upd);
        if (!ok) {
            Box prev = load_box(sfc + (size_t)(M - 1) * 6);
            ok = expand_from_hull_prev(c, hull, 2, prev, radius, max_vel, upd);
            if (!ok) { upd = prev; status |= ORC_SFC_REUSED; }
        }
        store_box(upd, sfc + (size_t)(M - 1) * 6);
    }
    if (lookups) *lookups += c.lookups;
    return status;
}

// ------------------------------------------------------------------------------------------
// goal line search: closed form of the 1-variable LP (goal_optimizer.cpp:7-136, 138-198)
//   min t in [0, 1+1e-5]  s.t.  a_r t + b_r >= 0
// CPLEX feasibility tolerance is emulated with 1e-6 on the row activity (parity unpinned).
// ------------------------------------------------------------------------------------------
int goal_agent(const orc_params* p, bool disturbed, const vec3f& pos, const vec3f& wp, const float* sfc_last,
               int K, const float* normal, const float* anchor, const double* d, size_t pair_stride_n,
               size_t pair_stride_a, size_t pair_stride_d, vec3f& goal) {
    const int M = p->M, P = P_of(p), D = p->dim;
    if (disturbed) { goal = pos; return ORC_OK; }                     // traj_planner.cpp:447-450
    if (goal.distance(wp) < kEpsF) { goal = wp; return ORC_OK; }     // goal_optimizer.cpp:12-14
    vec3f gw = goal - wp;                                             // float coefficients :165
    std::vector<double> A, B;
    auto add_row = [&](const float nrm[3], const float anc[3], double dd) {
        double a = 0, b = 0;
        for (int k = 0; k < D; k++) {
            a += (double)nrm[k] * (double)gw(k);
            b += (double)nrm[k] * ((double)wp(k) - (double)anc[k]);
        }
        A.push_back(a); B.push_back(b - dd);
    };
    if (p->use_sfc) {                                                 // Box::convertToLSCs :66-87
        for (int i = 0; i < D; i++) {
            float nmin[3] = {0, 0, 0}, nmax[3] = {0, 0, 0}, zero[3] = {0, 0, 0};
            nmin[i] = 1; nmax[i] = -1;
            add_row(nmin, zero, (double)sfc_last[i]);
            add_row(nmax, zero, -(double)sfc_last[3 + i]);
        }
    }
    for (int oi = 0; oi < K; oi++) {
        const float* nrm = normal + oi * pair_stride_n + (size_t)(M - 1) * 3;
        if (vec3f(nrm).norm() < kEpsF) continue;                      // :182-184
        const float* anc = anchor + oi * pair_stride_a + ((size_t)(M - 1) * P + p->n) * 3;
        double dd = d[oi * pair_stride_d + (size_t)(M - 1) * P + p->n];
        add_row(nrm, anc, dd);
    }
    double tlo = 0.0, thi = 1.0 + kEpsF;
    bool feasible = true;
    const double tol = 1e-6;
    for (size_t r = 0; r < A.size(); r++) {
        if (A[r] > 0) tlo = std::max(tlo, -B[r] / A[r]);
        else if (A[r] < 0) thi = std::min(thi, -B[r] / A[r]);
    }
    double t = std::min(tlo, 1.0 + kEpsF);
    for (size_t r = 0; r < A.size(); r++)
        if (A[r] * t + B[r] < -tol) feasible = false;
    (void)thi;
    if (feasible) {
        goal = gw * (float)t + wp;                                    // :51
        return ORC_OK;
    }
    // infeasible: "numerical error" rule :55-81
    bool numerical_error = true;
    if (p->use_sfc) {
        Box b = load_box(sfc_last);
        if (!point_in_box(b, goal)) numerical_error = false;
    }
    for (int oi = 0; oi < K; oi++) {
        const float* nrm = normal + oi * pair_stride_n + (size_t)(M - 1) * 3;
        vec3f nv(nrm);
        if (nv.norm() < kEpsF) continue;
        const float* anc = anchor + oi * pair_stride_a + ((size_t)(M - 1) * P + p->n) * 3;
        double dd = d[oi * pair_stride_d + (size_t)(M - 1) * P + p->n];
        double delta = nv.dot(goal - vec3f(anc)) - dd;                // :73
        if (delta < -kEpsF) numerical_error = false;
    }
    if (numerical_error) return ORC_OK;                               // keep goal :80
    return ORC_GOAL_INFEASIBLE;
}

// ------------------------------------------------------------------------------------------
// checkWaypointTrap (traj_planner.cpp:708-735): when the current goal or the next waypoint lies outside the region
// the AGENT LSCs of (M-1, n), the last SFC box and the communication box leave (isPointInFeasibleRegion,
// collision_constraints.cpp:586-598), the LSCs of every dynamic obstacle that can reach the waypoint are dropped.
// Slots [0, n_dyn) are the dynamic obstacles.  Returns 1 when trapped.
// ------------------------------------------------------------------------------------------
int waypoint_trap(const orc_params* p, const vec3f& goal, const vec3f& wp, const float* sfc_last, const float* comm_box,
                  int K, int n_dyn, float* normal, const float* anchor, double* d, size_t sn, size_t sa, size_t sd,
                  const float* dyn_pos, const float* dyn_vel, const double* dyn_radius, const double* dyn_max_acc,
                  double agent_radius, double uncertainty_horizon) {
    const int M = p->M, P = P_of(p);
    if (K == 0) return 0;                                                                   // obstacles.empty() :709
    auto feasible = [&](const vec3f& q) {
        for (int oi = n_dyn; oi < K; oi++) {
            const vec3f nv(normal + oi * sn + (size_t)(M - 1) * 3);
            const vec3f an(anchor + oi * sa + ((size_t)(M - 1) * P + p->n) * 3);
            const double dd = d[oi * sd + (size_t)(M - 1) * P + p->n];
            if (!((q - an).dot(nv) - dd > -kEps)) return false;                             // LSC::isPointInLSC :35-37
        }
        if (p->use_sfc && !point_in_box(load_box(sfc_last), q)) return false;
        return point_in_box(load_box(comm_box), q);
    };
    const bool trapped = !(feasible(goal) && feasible(wp));
    if (!trapped) return 0;
    for (int oi = 0; oi < n_dyn; oi++) {
        if (!obstacle_collides(vec3f(dyn_pos + 3 * oi), vec3f(dyn_vel + 3 * oi), dyn_radius[oi], dyn_max_acc[oi], wp,
                               agent_radius, M * p->dt, uncertainty_horizon)) continue;
        for (int m = 0; m < M; m++) {
            float* nr = normal + oi * sn + (size_t)m * 3;
            nr[0] = nr[1] = nr[2] = 0.f;                                                     // default LSC: skipped by the QP :731
            for (int i = 0; i < P; i++) d[oi * sd + (size_t)m * P + i] = 0.0;
        }
    }
    return 1;
}

// ------------------------------------------------------------------------------------------
// QP
// ------------------------------------------------------------------------------------------
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
// traj_optimizer.cpp:172-187 with phi_n = 1, polynomial.hpp:280-293
void q_base(const orc_params* p, double* Q) {
    const int P = P_of(p), n = p->n, k = p->phi;
    std::vector<double> B(P * P, 0.0), Z(P * P, 0.0), T(P * P, 0.0);
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
    const double sc = std::pow(p->dt, -2 * k + 1);
    for (int i = 0; i < P; i++)
        for (int j = 0; j < P; j++) {
            double s = 0;
            for (int a = 0; a < P; a++) s += T[i * P + a] * B[j * P + a];
            Q[i * P + j] = s * sc;
        }
}

// x-space variable expressed in the reduced (null-space) coordinates y:
//   x = c0 + sum coef[t] * y[idx[t]]
// Free coordinates per axis: control points 3,4,5 of every segment (the last segment's three
// coincide, traj_optimizer.cpp:515-524); points 0,1,2 follow from the C0/C1/C2 continuity rows
// (:338-381) and, for segment 0, from the initial state rows (:335-352).
struct XExpr { int nv; int idx[3]; double coef[3]; double c0; };

struct SparseRow { int nnz; int idx[10]; double val[10]; double rhs; };   // sum val*y <= rhs (9 trajectory terms + 1 slack)

struct QpWork {
    int D, M, P, nx, ny;
    std::vector<XExpr> xe;       // [D][M][P]
    std::vector<double> Px;      // quadratic form blocks per (k,m): [D][M][P][P]  (objective x'Px)
    std::vector<double> qx;      // linear term [nx]
    double cx;                   // constant
    std::vector<SparseRow> rows;
};

inline int xid(const QpWork& w, int k, int m, int i) { return (k * w.M + m) * w.P + i; }

void add_row(QpWork& w, const int* xi, const double* xc, int nxv, double rhs) {
    // a'x <= rhs  ->  sparse row in y
    SparseRow r; r.nnz = 0; r.rhs = rhs;
    for (int t = 0; t < nxv; t++) {
        const XExpr& e = w.xe[xi[t]];
        r.rhs -= xc[t] * e.c0;
        for (int u = 0; u < e.nv; u++) {
            double cv = xc[t] * e.coef[u];
            int found = -1;
            for (int z = 0; z < r.nnz; z++) if (r.idx[z] == e.idx[u]) { found = z; break; }
            if (found >= 0) r.val[found] += cv;
            else { r.idx[r.nnz] = e.idx[u]; r.val[r.nnz] = cv; r.nnz++; }
        }
    }
    w.rows.push_back(r);
}

// Dense primal-dual interior point (Mehrotra predictor-corrector) on
//   min 1/2 y'Hy + g'y   s.t.  G y <= h     (H positive definite)
int ipm_solve(int ny, const std::vector<double>& H, const std::vector<double>& g,
              const std::vector<SparseRow>& rows, std::vector<double>& y, int* iters_out) {
    const int m = (int)rows.size();
    std::vector<double> s(m), z(m), Gy(m), rp(m), ds(m), dz(m), dsa(m), dza(m);
    std::vector<double> W(ny * ny), L(ny * ny), rhs(ny), dy(ny), rd(ny), tmp(m);
    auto mulG = [&](const std::vector<double>& v, std::vector<double>& out) {
        for (int r = 0; r < m; r++) {
            double a = 0;
            for (int t = 0; t < rows[r].nnz; t++) a += rows[r].val[t] * v[rows[r].idx[t]];
            out[r] = a;
        }
    };
    mulG(y, Gy);
    double hscale = 0;
    for (int i = 0; i < ny; i++) hscale = std::max(hscale, std::fabs(H[i * ny + i]));
    for (int r = 0; r < m; r++) { s[r] = std::max(rows[r].rhs - Gy[r], 1e-2); z[r] = 1.0 / s[r] * 1e-2; }
    int status = ORC_QP_MAXITER;
    int it = 0;
    const int max_it = 80;
    bool acceptable = false;
    for (it = 0; it < max_it; it++) {
        mulG(y, Gy);
        // residuals
        for (int i = 0; i < ny; i++) {
            double a = g[i];
            for (int j = 0; j < ny; j++) a += H[i * ny + j] * y[j];
            rd[i] = a;
        }
        for (int r = 0; r < m; r++)
            for (int t = 0; t < rows[r].nnz; t++) rd[rows[r].idx[t]] += rows[r].val[t] * z[r];
        double rp_inf = 0, rd_inf = 0, mu = 0, g_inf = 0;
        for (int r = 0; r < m; r++) { rp[r] = Gy[r] + s[r] - rows[r].rhs; rp_inf = std::max(rp_inf, std::fabs(rp[r])); mu += s[r] * z[r]; }
        for (int i = 0; i < ny; i++) { rd_inf = std::max(rd_inf, std::fabs(rd[i])); g_inf = std::max(g_inf, std::fabs(g[i])); }
        if (m > 0) mu /= m;
        if (rp_inf <= 1e-10 && rd_inf <= ORC_QP_TOL_RD * (1.0 + g_inf) && mu <= ORC_QP_TOL_MU) { status = ORC_OK; break; }
        // "acceptable" level: if the factorisation breaks down (or the cap is hit) after this level was reached,
        // the iterate is returned as converged (the tight target above is at the edge of what fp64 Cholesky of
        // the barrier-scaled system can deliver)
        if (rp_inf <= 1e-10 && rd_inf <= 1e-9 * (1.0 + g_inf) && mu <= 1e-11) acceptable = true;
        // W = H + G' diag(z/s) G
        W = H;
        for (int r = 0; r < m; r++) {
            double dd = z[r] / s[r];
            const SparseRow& R = rows[r];
            for (int a = 0; a < R.nnz; a++)
                for (int b = 0; b < R.nnz; b++) W[R.idx[a] * ny + R.idx[b]] += dd * R.val[a] * R.val[b];
        }
        // Cholesky
        L = W;
        bool ok = true;
        for (int j = 0; j < ny && ok; j++) {
            double dj = L[j * ny + j];
            for (int k = 0; k < j; k++) dj -= L[j * ny + k] * L[j * ny + k];
            if (!(dj > 0)) { ok = false; break; }
            dj = std::sqrt(dj);
            L[j * ny + j] = dj;
            for (int i = j + 1; i < ny; i++) {
                double a = L[i * ny + j];
                for (int k = 0; k < j; k++) a -= L[i * ny + k] * L[j * ny + k];
                L[i * ny + j] = a / dj;
            }
        }
        if (!ok) { status = acceptable ? ORC_OK : ORC_QP_NUMERIC; break; }
        auto solve = [&](std::vector<double>& b) {
            for (int i = 0; i < ny; i++) {
                double a = b[i];
                for (int k = 0; k < i; k++) a -= L[i * ny + k] * b[k];
                b[i] = a / L[i * ny + i];
            }
            for (int i = ny - 1; i >= 0; i--) {
                double a = b[i];
                for (int k = i + 1; k < ny; k++) a -= L[k * ny + i] * b[k];
                b[i] = a / L[i * ny + i];
            }
        };
        // direction for a given complementarity residual rc (vector in tmp): returns dy, ds, dz
        auto direction = [&](const std::vector<double>& rc, std::vector<double>& dyv, std::vector<double>& dsv,
                             std::vector<double>& dzv) {
            for (int i = 0; i < ny; i++) dyv[i] = -rd[i];
            for (int r = 0; r < m; r++) {
                double c = (-rc[r] + z[r] * rp[r]) / s[r];
                for (int t = 0; t < rows[r].nnz; t++) dyv[rows[r].idx[t]] -= rows[r].val[t] * c;
            }
            solve(dyv);
            std::vector<double>& Gd = Gy;   // reuse buffer
            mulG(dyv, Gd);
            for (int r = 0; r < m; r++) {
                dsv[r] = -rp[r] - Gd[r];
                dzv[r] = (-rc[r] - z[r] * dsv[r]) / s[r];
            }
        };
        auto max_step = [&](const std::vector<double>& dsv, const std::vector<double>& dzv) {
            double a = 1.0;
            for (int r = 0; r < m; r++) {
                if (dsv[r] < 0) a = std::min(a, -s[r] / dsv[r]);
                if (dzv[r] < 0) a = std::min(a, -z[r] / dzv[r]);
            }
            return a;
        };
        for (int r = 0; r < m; r++) tmp[r] = s[r] * z[r];
        direction(tmp, dy, dsa, dza);
        double a_aff = max_step(dsa, dza);
        double mu_aff = 0;
        for (int r = 0; r < m; r++) mu_aff += (s[r] + a_aff * dsa[r]) * (z[r] + a_aff * dza[r]);
        if (m > 0) mu_aff /= m;
        double sigma = mu > 0 ? std::pow(mu_aff / mu, 3.0) : 0.0;
        for (int r = 0; r < m; r++) tmp[r] = s[r] * z[r] + dsa[r] * dza[r] - sigma * mu;
        direction(tmp, dy, ds, dz);
        double a = max_step(ds, dz);
        a = std::min(1.0, 0.995 * a);
        for (int i = 0; i < ny; i++) y[i] += a * dy[i];
        for (int r = 0; r < m; r++) { s[r] += a * ds[r]; z[r] += a * dz[r]; }
    }
    (void)hscale;
    if (status == ORC_QP_MAXITER && acceptable) status = ORC_OK;
    if (iters_out) *iters_out = it;
    return status;
}

int qp_solve(const orc_params* p, const vec3f& pos, const vec3f& vel, const vec3f& acc, const vec3f& goal,
             const vec3f& wp, double radius, double max_vel, double max_acc, double nominal_vel,
             const float* sfc, int K, const float* normal, const float* anchor, const double* d,
             size_t sn, size_t sa, size_t sd, const float* init_traj, float* traj_out, double* x_out,
             double* cost, double* max_violation, int* iters,
             int n_dyn = 0, double slack_weight = 0.0, double* slack_out = nullptr) {
    // The first n_dyn obstacle slots are dynamic (non-agent) obstacles: their LSC rows carry one slack variable per
    // (obstacle, segment), epsilon <= 0, with cost slack_weight (M - m)/M epsilon^2 (traj_optimizer.cpp:272-283, 317-331, 436-448).
    const int M = p->M, P = P_of(p), D = p->dim, n = p->n, phi = p->phi;
    const double dt = p->dt;
    QpWork w;
    w.D = D; w.M = M; w.P = P; w.nx = D * M * P;
    const int nyd = 3 * M - 2;     // per axis
    w.ny = D * nyd;
    w.xe.assign(w.nx, XExpr{0, {0, 0, 0}, {0, 0, 0}, 0.0});
    auto yid = [&](int k, int m, int j) {       // j = 0,1,2 -> control point 3+j
        if (m == M - 1) return k * nyd + 3 * (M - 1);
        return k * nyd + 3 * m + j;
    };
    for (int k = 0; k < D; k++) {
        // initial state rows traj_optimizer.cpp:335-352
        double c0 = (double)pos(k);
        double c1 = c0 + (double)vel(k) * dt / n;
        double c2 = (double)acc(k) * dt * dt / (n * (n - 1)) + 2 * c1 - c0;
        w.xe[xid(w, k, 0, 0)].c0 = c0;
        w.xe[xid(w, k, 0, 1)].c0 = c1;
        w.xe[xid(w, k, 0, 2)].c0 = c2;
        for (int m = 0; m < M; m++) {
            for (int j = 0; j < 3; j++) {
                XExpr& e = w.xe[xid(w, k, m, 3 + j)];
                e.nv = 1; e.idx[0] = yid(k, m, j); e.coef[0] = 1.0; e.c0 = 0;
            }
            if (m + 1 < M) {
                int y3 = yid(k, m, 0), y4 = yid(k, m, 1), y5 = yid(k, m, 2);
                XExpr& e0 = w.xe[xid(w, k, m + 1, 0)];
                e0.nv = 1; e0.idx[0] = y5; e0.coef[0] = 1;
                XExpr& e1 = w.xe[xid(w, k, m + 1, 1)];
                e1.nv = 2; e1.idx[0] = y5; e1.coef[0] = 2; e1.idx[1] = y4; e1.coef[1] = -1;
                XExpr& e2 = w.xe[xid(w, k, m + 1, 2)];
                e2.nv = 3; e2.idx[0] = y5; e2.coef[0] = 4; e2.idx[1] = y4; e2.coef[1] = -4; e2.idx[2] = y3; e2.coef[2] = 1;
            }
        }
    }
    // objective x'Px + q'x + c : traj_optimizer.cpp:285-315
    std::vector<double> Q(P * P);
    q_base(p, Q.data());
    w.Px.assign((size_t)D * M * P * P, 0.0);
    w.qx.assign(w.nx, 0.0);
    w.cx = 0;
    for (int k = 0; k < D; k++)
        for (int m = 0; m < M; m++)
            for (int i = 0; i < P; i++)
                for (int j = 0; j < P; j++)
                    w.Px[((size_t)(k * M + m) * P + i) * P + j] = p->w_control * Q[i * P + j];
    double ideal_time = (goal - pos).norm() / nominal_vel;                                   // :543-551
    int ts = std::max((int)((M * dt - ideal_time + kEps) / dt), 1);
    for (int m = M - ts; m < M; m++)
        for (int k = 0; k < D; k++) {
            double gk = (double)goal(k);
            w.Px[((size_t)(k * M + m) * P + n) * P + n] += p->w_terminal;
            w.qx[xid(w, k, m, n)] += -2.0 * p->w_terminal * gk;
            w.cx += p->w_terminal * gk * gk;
        }
    // inequality rows
    auto skip0 = [&](int m, int i) { return m == 0 && i < phi; };
    for (int k = 0; k < D; k++)                                                              // bounds :251-265
        for (int m = 0; m < M; m++)
            for (int i = 0; i < P; i++) {
                if (m == 0 && i < 3) continue;
                int xi = xid(w, k, m, i);
                double one = 1, neg = -1;
                add_row(w, &xi, &one, 1, (double)(float)p->world_max[k]);
                add_row(w, &xi, &neg, 1, -(double)(float)p->world_min[k]);
            }
    if (p->use_sfc) {                                                                         // SFC :384-410
        for (int m = 0; m < M; m++)
            for (int f = 0; f < D; f++)
                for (int j = 0; j < P; j++) {
                    if (skip0(m, j)) continue;
                    int xi = xid(w, f, m, j);
                    double one = 1, neg = -1;
                    add_row(w, &xi, &neg, 1, -(double)sfc[(size_t)m * 6 + f]);         //  x >= min
                    add_row(w, &xi, &one, 1, (double)sfc[(size_t)m * 6 + 3 + f]);      //  x <= max
                }
    }
    for (int oi = 0; oi < K; oi++)                                                            // LSC :412-450
        for (int m = 0; m < M; m++) {
            const float* nrm = normal + oi * sn + (size_t)m * 3;
            if (vec3f(nrm).norm() < kEpsF) continue;
            for (int i = 0; i < P; i++) {
                if (skip0(m, i)) continue;
                const float* anc = anchor + oi * sa + ((size_t)m * P + i) * 3;
                double dd = d[oi * sd + (size_t)m * P + i];
                int xi[3]; double xc[3]; double rhs = -dd;
                for (int k = 0; k < D; k++) {
                    xi[k] = xid(w, k, m, i);
                    xc[k] = -(double)nrm[k];
                    rhs -= (double)nrm[k] * (double)anc[k];
                }
                add_row(w, xi, xc, D, rhs);       // -n.x <= -(n.anchor + d)
                if (oi < n_dyn) {                 // n.(x - anchor) - (d + eps) >= 0   :443-444
                    SparseRow& r = w.rows.back();
                    r.idx[r.nnz] = w.ny + M * oi + m; r.val[r.nnz] = 1.0; r.nnz++;
                }
            }
        }
    for (int s = 0; s < n_dyn * M; s++) {                                                    // eps <= 0 :274
        SparseRow r; r.nnz = 1; r.idx[0] = w.ny + s; r.val[0] = 1.0; r.rhs = 0.0;
        w.rows.push_back(r);
    }
    for (int k = 0; k < D; k++)                                                               // dynamics :452-487
        for (int m = 0; m < M; m++) {
            for (int i = 0; i < n; i++) {
                if (m == 0 && (i == 0 || i == 1)) continue;
                int xi[2] = {xid(w, k, m, i + 1), xid(w, k, m, i)};
                double sc = std::pow(dt, -1) * n;
                double c1[2] = {sc, -sc}, c2[2] = {-sc, sc};
                add_row(w, xi, c1, 2, max_vel);
                add_row(w, xi, c2, 2, max_vel);
            }
            for (int i = 0; i < n - 1; i++) {
                if (m == 0 && i == 0) continue;
                int xi[3] = {xid(w, k, m, i + 2), xid(w, k, m, i + 1), xid(w, k, m, i)};
                double sc = std::pow(dt, -2) * n * (n - 1);
                double c1[3] = {sc, -2 * sc, sc}, c2[3] = {-sc, 2 * sc, -sc};
                add_row(w, xi, c1, 3, max_acc);
                add_row(w, xi, c2, 3, max_acc);
            }
        }
    if (p->comm_range > 0) {                                                                  // :490-513
        for (int k = 0; k < D; k++)
            for (int mi = 0; mi < M; mi++)
                for (int m = mi; m < M; m++) {
                    int xi[2] = {xid(w, k, m, n), xid(w, k, mi, 0)};
                    double c1[2] = {1, -1}, c2[2] = {-1, 1};
                    double rhs = 0.5 * p->comm_range - radius;
                    add_row(w, xi, c1, 2, rhs);
                    add_row(w, xi, c2, 2, rhs);
                }
        for (int k = 0; k < D; k++)
            for (int m = 0; m < M; m++) {
                int xi = xid(w, k, m, n);
                double one = 1, neg = -1;
                double rhs = 0.5 * p->comm_range - kEpsF;
                add_row(w, &xi, &one, 1, rhs + (double)wp(k));
                add_row(w, &xi, &neg, 1, rhs - (double)wp(k));
            }
    }
    // reduced objective: 1/2 y'Hy + g'y + c0
    const int ny = w.ny + n_dyn * M;                 // trajectory unknowns, then the slack variables
    std::vector<double> H((size_t)ny * ny, 0.0), g(ny, 0.0);
    for (int oi = 0; oi < n_dyn; oi++)                                                       // :317-331
        for (int m = 0; m < M; m++) {
            const int e = w.ny + M * oi + m;
            H[(size_t)e * ny + e] = 2.0 * slack_weight * ((double)(M - m) / M);
        }
    double c0 = w.cx;
    for (int k = 0; k < D; k++)
        for (int m = 0; m < M; m++) {
            const double* Pb = &w.Px[(size_t)(k * M + m) * P * P];
            for (int i = 0; i < P; i++) {
                const XExpr& ei = w.xe[xid(w, k, m, i)];
                for (int j = 0; j < P; j++) {
                    const XExpr& ej = w.xe[xid(w, k, m, j)];
                    double pij = Pb[i * P + j];
                    if (pij == 0) continue;
                    c0 += pij * ei.c0 * ej.c0;
                    for (int a = 0; a < ei.nv; a++) g[ei.idx[a]] += pij * ei.coef[a] * ej.c0;
                    for (int b = 0; b < ej.nv; b++) g[ej.idx[b]] += pij * ej.coef[b] * ei.c0;
                    for (int a = 0; a < ei.nv; a++)
                        for (int b = 0; b < ej.nv; b++)
                            H[(size_t)ei.idx[a] * ny + ej.idx[b]] += 2.0 * pij * ei.coef[a] * ej.coef[b];
                }
            }
        }
    for (int xi = 0; xi < w.nx; xi++) {
        if (w.qx[xi] == 0) continue;
        const XExpr& e = w.xe[xi];
        c0 += w.qx[xi] * e.c0;
        for (int a = 0; a < e.nv; a++) g[e.idx[a]] += w.qx[xi] * e.coef[a];
    }
    // start from the initial trajectory's free control points
    std::vector<double> y(ny, 0.0);
    for (int k = 0; k < D; k++)
        for (int m = 0; m < M; m++)
            for (int j = 0; j < 3; j++) {
                if (m == M - 1 && j < 2) continue;
                y[yid(k, m, j)] = (double)init_traj[((size_t)m * P + 3 + j) * 3 + k];
            }
    int it = 0;
    int status = ipm_solve(ny, H, g, w.rows, y, &it);
    if (iters) *iters = it;
    // recover x
    std::vector<double> x(w.nx);
    for (int xi = 0; xi < w.nx; xi++) {
        const XExpr& e = w.xe[xi];
        double a = e.c0;
        for (int t = 0; t < e.nv; t++) a += e.coef[t] * y[e.idx[t]];
        x[xi] = a;
    }
    if (x_out) std::memcpy(x_out, x.data(), sizeof(double) * w.nx);
    // objective in x-space (constant included, like IloCplex::getObjValue, :109)
    double obj = w.cx;
    for (int k = 0; k < D; k++)
        for (int m = 0; m < M; m++) {
            const double* Pb = &w.Px[(size_t)(k * M + m) * P * P];
            for (int i = 0; i < P; i++)
                for (int j = 0; j < P; j++) obj += Pb[i * P + j] * x[xid(w, k, m, i)] * x[xid(w, k, m, j)];
        }
    for (int xi = 0; xi < w.nx; xi++) obj += w.qx[xi] * x[xi];
    for (int oi = 0; oi < n_dyn; oi++)
        for (int m = 0; m < M; m++) {
            const double e = y[w.ny + M * oi + m];
            obj += slack_weight * ((double)(M - m) / M) * e * e;
            if (slack_out) slack_out[oi * M + m] = e;
        }
    (void)c0;
    if (cost) *cost = obj;
    double viol = 0;
    for (const SparseRow& r : w.rows) {
        double a = -r.rhs;
        for (int t = 0; t < r.nnz; t++) a += r.val[t] * y[r.idx[t]];
        viol = std::max(viol, a);
    }
    if (max_violation) *max_violation = viol;
    for (int m = 0; m < M; m++)                                                               // :71-83
        for (int i = 0; i < P; i++) {
            float* o = traj_out + ((size_t)m * P + i) * 3;
            o[0] = (float)x[xid(w, 0, m, i)];
            o[1] = (float)x[xid(w, 1, m, i)];
            o[2] = (D == 3) ? (float)x[xid(w, 2, m, i)] : (float)p->z_2d;
        }
    return status;
}

// one agent per task, dynamic schedule, plain std::thread (no OpenMP dependency)
void parallel_for(int n, int n_threads, const std::function<void(int)>& body) {
    if (n_threads <= 1 || n <= 1) { for (int i = 0; i < n; i++) body(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([&]() { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) body(i); });
    for (auto& t : th) t.join();
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

double orc_gjk_hull_origin(const double* pts, int npts, double v[3], int* iters, int* simplex_n) {
    return gjk::hull_origin(reinterpret_cast<const double(*)[3]>(pts), npts, v, iters, simplex_n);
}

void orc_closest_segments(const float a0[3], const float a1[3], const float b0[3], const float b1[3],
                          float p1[3], float p2[3], double* dist) {
    Closest c = closest_segments(vec3f(a0), vec3f(a1), vec3f(b0), vec3f(b1));
    c.p1.store(p1); c.p2.store(p2); *dist = c.dist;
}

void orc_predict(const orc_params* p, int N, int seq, const float* pos, const float* vel,
                 const float* prev_traj, const uint8_t* disturbed, float* init_traj, float* pred_traj) {
    const size_t L = traj_len(p);
    for (int a = 0; a < N; a++) {
        vec3f ps(pos + 3 * a), vl(vel + 3 * a);
        float* it = init_traj + L * a;
        float* pt = pred_traj + L * a;
        if (seq < 2) {                                   // traj_planner.cpp:293-296, 410-411
            const_vel_traj(p, ps, vl, it);
        } else {
            shift_traj(p, prev_traj + L * a, it);
        }
        std::memcpy(pt, it, L * sizeof(float));
        // others' view: checkObstacleDisturbance :329-336
        if ((vec3f(pt) - ps).norm() > p->reset_threshold) const_vel_traj(p, ps, vec3f(0, 0, 0), pt);
        // own view: initialTrajPlanningCheck :435-441
        if (disturbed && disturbed[a]) const_vel_traj(p, ps, vec3f(0, 0, 0), it);
    }
}

// slots [0, first) of every list are left to the caller (dynamic obstacles); nbr_cnt counts them
static int neighbours_from(const orc_params* p, int N, const float* pos, int max_nbr, int first, int32_t* nbr_idx, int32_t* nbr_cnt) {
    int overflow = 0;
    for (int a = 0; a < N; a++) {
        int c = first;
        vec3f pa(pos + 3 * a);
        for (int j = 0; j < N; j++) {
            if (j == a) continue;
            double dist = linf_distance(pa, vec3f(pos + 3 * j));       // multi_sync_simulator.cpp:484-491
            if (p->comm_range > 0 && dist > p->comm_range) continue;
            if (c < max_nbr) nbr_idx[(size_t)a * max_nbr + c] = j;
            c++;
        }
        if (c > max_nbr) { overflow = 1; c = max_nbr; }
        nbr_cnt[a] = c;
    }
    return overflow;
}
int orc_neighbours(const orc_params* p, int N, const float* pos, int max_nbr, int32_t* nbr_idx, int32_t* nbr_cnt) {
    return neighbours_from(p, N, pos, max_nbr, 0, nbr_idx, nbr_cnt);
}

void orc_lsc_batch(const orc_params* p, int N, const float* init_traj, const float* pred_traj,
                   const float* goal_cur, const double* radius, const double* downwash, int max_nbr,
                   const int32_t* nbr_idx, const int32_t* nbr_cnt, float* normal, float* anchor,
                   double* d, int64_t* hist) {
    const int M = p->M, P = P_of(p);
    const size_t L = traj_len(p);
    for (int a = 0; a < N; a++)
        for (int c = 0; c < nbr_cnt[a]; c++) {
            int j = nbr_idx[(size_t)a * max_nbr + c];
            size_t pr = (size_t)a * max_nbr + c;
            lsc_pair(p, init_traj + L * a, pred_traj + L * j, vec3f(goal_cur + 3 * a), vec3f(goal_cur + 3 * j),
                     radius[a], downwash[a], radius[j], downwash[j], normal + pr * M * 3,
                     anchor + pr * M * P * 3, d + pr * M * P, hist);
        }
}

int orc_sfc_expand(const orc_params* p, const orc_edt* edt, const float box_in[6], double margin,
                   double max_vel, float box_out[6], int64_t* n_lookups) {
    SfcCtx c{p, EdtView{edt, edt->res, 1.0 / edt->res}};
    Box out, in = load_box(box_in);
    bool ok = expand_incrementally(c, in, margin, max_vel, out);
    if (ok) store_box(out, box_out);
    if (n_lookups) *n_lookups += c.lookups;
    return ok ? 1 : 0;
}

void orc_sfc_batch(const orc_params* p, const orc_edt* edt, int N, const uint8_t* init_flag,
                   const float* pos, const float* init_traj, const float* goal_cur,
                   const float* waypoint, const double* radius, const double* max_vel,
                   float* sfc, int32_t* status, int64_t* n_lookups) {
    const size_t L = traj_len(p);
    for (int a = 0; a < N; a++) {
        int st = sfc_agent(p, edt, init_flag[a] != 0, vec3f(pos + 3 * a), init_traj + L * a, vec3f(goal_cur + 3 * a),
                           vec3f(waypoint + 3 * a), radius[a], max_vel[a], sfc + (size_t)a * p->M * 6, n_lookups);
        if (status) status[a] |= st;
    }
}

void orc_goal_batch(const orc_params* p, int N, const uint8_t* disturbed, const float* pos,
                    const float* waypoint, const float* sfc, int max_nbr, const int32_t* nbr_cnt,
                    const float* normal, const float* anchor, const double* d, float* goal_cur, int32_t* status) {
    const int M = p->M, P = P_of(p);
    for (int a = 0; a < N; a++) {
        vec3f g(goal_cur + 3 * a);
        size_t pr = (size_t)a * max_nbr;
        int st = goal_agent(p, disturbed && disturbed[a], vec3f(pos + 3 * a), vec3f(waypoint + 3 * a),
                            sfc ? sfc + ((size_t)a * M + (M - 1)) * 6 : nullptr, nbr_cnt[a], normal + pr * M * 3,
                            anchor + pr * M * P * 3, d + pr * M * P, (size_t)M * 3, (size_t)M * P * 3,
                            (size_t)M * P, g);
        g.store(goal_cur + 3 * a);
        if (status) status[a] |= st;
    }
}

int orc_qp_solve(const orc_params* p, const float pos[3], const float vel[3], const float acc[3],
                 const float goal[3], const float waypoint[3], double radius, double max_vel,
                 double max_acc, double nominal_vel, const float* sfc, int K, const float* normal,
                 const float* anchor, const double* d, const float* init_traj, float* traj_out,
                 double* x_out, double* cost, double* max_violation, int* iters) {
    const int M = p->M, P = P_of(p);
    return qp_solve(p, vec3f(pos), vec3f(vel), vec3f(acc), vec3f(goal), vec3f(waypoint), radius, max_vel, max_acc,
                    nominal_vel, sfc, K, normal, anchor, d, (size_t)M * 3, (size_t)M * P * 3, (size_t)M * P,
                    init_traj, traj_out, x_out, cost, max_violation, iters);
}

// traj_planner.cpp:108-133 for every agent, agents in index order (multi_sync_simulator.cpp:516-524)
void orc_step(const orc_params* p, orc_step_io* io) { orc_step_range(p, io, 0, io->N); }

// Same as orc_step, but only the agents [a_begin, a_end) are replanned (prediction and neighbour lists are
// still built for the whole swarm, they are inputs of the replanned agents).  Used for bounded CPU-baseline
// samples of large swarms.
void orc_step_range(const orc_params* p, orc_step_io* io, int a_begin, int a_end) {
    const int N = io->N, M = p->M, P = P_of(p), K = io->max_nbr;
    const int NR = a_end - a_begin;
    const size_t L = traj_len(p);
    double t0 = now_s();
    for (int a = 0; a < N; a++) io->status[a] = ORC_OK;
    orc_predict(p, N, io->seq, io->pos, io->vel, io->prev_traj, io->disturbed, io->init_traj, io->pred_traj);
    // dynamic obstacles come first in every agent's obstacle list (multi_sync_simulator.cpp:476-480): slots [0, nd),
    // list entries N + o; constant-velocity prediction (traj_planner.cpp:303-305) and predicted sizes (:338-368)
    const int nd = io->n_dyn > 0 ? io->n_dyn : 0;
    std::vector<float> dyn_pred((size_t)nd * L);
    std::vector<double> dyn_size((size_t)nd * M * P);
    for (int o = 0; o < nd; o++) {
        const_vel_traj(p, vec3f(io->dyn_pos + 3 * o), vec3f(io->dyn_vel + 3 * o), dyn_pred.data() + L * o);
        obstacle_sizes(p, io->dyn_size_prediction, io->dyn_uncertainty_horizon, io->dyn_radius[o], io->dyn_max_acc[o],
                       dyn_size.data() + (size_t)o * M * P);
    }
    if (neighbours_from(p, N, io->pos, K, nd, io->nbr_idx, io->nbr_cnt))
        for (int a = 0; a < N; a++) io->status[a] |= ORC_NBR_OVERFLOW;
    for (int a = 0; a < N; a++)
        for (int o = 0; o < nd; o++) io->nbr_idx[(size_t)a * K + o] = N + o;
    double t1 = now_s();
    const int nt = io->n_threads > 1 ? io->n_threads : 1;
    parallel_for(NR, nt, [&](int ar) {
        const int a = a_begin + ar;
        for (int c = 0; c < io->nbr_cnt[a]; c++) {
            int j = io->nbr_idx[(size_t)a * K + c];
            size_t pr = (size_t)a * K + c;
            if (c < nd) {
                lsc_dynamic(p, io->init_traj + L * a, dyn_pred.data() + L * c, dyn_size.data() + (size_t)c * M * P,
                            io->radius[a], io->dyn_radius[c], io->dyn_downwash[c], io->lsc_normal + pr * M * 3,
                            io->lsc_anchor + pr * M * P * 3, io->lsc_d + pr * M * P);
                continue;
            }
            lsc_pair(p, io->init_traj + L * a, io->pred_traj + L * j, vec3f(io->goal_cur + 3 * a),
                     vec3f(io->goal_cur + 3 * j), io->radius[a], io->downwash[a], io->radius[j], io->downwash[j],
                     io->lsc_normal + pr * M * 3, io->lsc_anchor + pr * M * P * 3, io->lsc_d + pr * M * P, nullptr);
        }
    });
    double t2 = now_s();
    if (p->use_sfc) {
        parallel_for(NR, nt, [&](int ar) {
            const int a = a_begin + ar;
            bool init = io->sfc_init_flag[a] != 0 || (io->disturbed && io->disturbed[a]);   // traj_planner.cpp:439, 693-695
            int st = sfc_agent(p, io->edt, init, vec3f(io->pos + 3 * a), io->init_traj + L * a,
                               vec3f(io->goal_cur + 3 * a), vec3f(io->waypoint + 3 * a), io->radius[a],
                               io->max_vel[a], io->sfc + (size_t)a * M * 6, nullptr);
            io->sfc_init_flag[a] = 0;
            io->status[a] |= st;
            if (!init && p->comm_range > 0 && io->comm_box) {                               // constructCommunicationRange :538-546
                const vec3f wp(io->waypoint + 3 * a), dl((float)(0.5 * p->comm_range), (float)(0.5 * p->comm_range), (float)(0.5 * p->comm_range));
                (wp - dl).store(io->comm_box + 6 * a); (wp + dl).store(io->comm_box + 6 * a + 3);
            }
        });
    }
    if (nd > 0) {
        static const float zero_box[6] = {0, 0, 0, 0, 0, 0};
        for (int a = a_begin; a < a_end; a++) {
            size_t pr = (size_t)a * K;
            const int tr = waypoint_trap(p, vec3f(io->goal_cur + 3 * a), vec3f(io->waypoint + 3 * a),
                                         p->use_sfc ? io->sfc + ((size_t)a * M + (M - 1)) * 6 : nullptr,
                                         io->comm_box ? io->comm_box + 6 * a : zero_box, io->nbr_cnt[a], nd,
                                         io->lsc_normal + pr * M * 3, io->lsc_anchor + pr * M * P * 3, io->lsc_d + pr * M * P,
                                         (size_t)M * 3, (size_t)M * P * 3, (size_t)M * P, io->dyn_pos, io->dyn_vel,
                                         io->dyn_radius, io->dyn_max_acc, io->radius[a], io->dyn_uncertainty_horizon);
            if (io->trap) io->trap[a] = (uint8_t)tr;
        }
    }
    double t3 = now_s();
    // goal planning must see every neighbour's PREVIOUS goal in the LSC stage above; update after it.
    std::vector<float> new_goal(io->goal_cur, io->goal_cur + (size_t)3 * N);
    for (int a = a_begin; a < a_end; a++) {
        vec3f g(io->goal_cur + 3 * a);
        size_t pr = (size_t)a * K;
        int st = goal_agent(p, io->disturbed && io->disturbed[a], vec3f(io->pos + 3 * a), vec3f(io->waypoint + 3 * a),
                            p->use_sfc ? io->sfc + ((size_t)a * M + (M - 1)) * 6 : nullptr, io->nbr_cnt[a] - nd,
                            io->lsc_normal + (pr + nd) * M * 3, io->lsc_anchor + (pr + nd) * M * P * 3, io->lsc_d + (pr + nd) * M * P,
                            (size_t)M * 3, (size_t)M * P * 3, (size_t)M * P, g);          // dynamic obstacles skipped: goal_optimizer.cpp:176-178
        g.store(new_goal.data() + 3 * a);
        io->status[a] |= st;
    }
    std::memcpy(io->goal_cur, new_goal.data(), sizeof(float) * 3 * N);
    double t4 = now_s();
    const int nxa = p->dim * M * P;
    parallel_for(NR, nt, [&](int ar) {
        const int a = a_begin + ar;
        size_t pr = (size_t)a * K;
        std::vector<float> out(L);
        double cost = 0, viol = 0;
        int it = 0;
        int st = qp_solve(p, vec3f(io->pos + 3 * a), vec3f(io->vel + 3 * a), vec3f(io->acc + 3 * a),
                          vec3f(io->goal_cur + 3 * a), vec3f(io->waypoint + 3 * a), io->radius[a], io->max_vel[a],
                          io->max_acc[a], io->nominal_vel[a], io->sfc + (size_t)a * M * 6, io->nbr_cnt[a],
                          io->lsc_normal + pr * M * 3, io->lsc_anchor + pr * M * P * 3, io->lsc_d + pr * M * P,
                          (size_t)M * 3, (size_t)M * P * 3, (size_t)M * P, io->init_traj + L * a, out.data(),
                          io->qp_x ? io->qp_x + (size_t)a * nxa : nullptr, &cost, &viol, &it, nd, io->slack_collision_weight,
                          (nd && io->qp_slack) ? io->qp_slack + (size_t)a * nd * M : nullptr);
        if (st != ORC_OK) std::memcpy(out.data(), io->init_traj + L * a, L * sizeof(float));   // failsafe traj_planner.cpp:775-776
        std::memcpy(io->prev_traj + L * a, out.data(), L * sizeof(float));                       // :57
        io->status[a] |= st;
        if (io->cost) io->cost[a] = cost;
        if (io->max_violation) io->max_violation[a] = viol;
        if (io->qp_iters) io->qp_iters[a] = it;
    });
    double t5 = now_s();
    if (io->stage_seconds) {
        io->stage_seconds[0] = t1 - t0; io->stage_seconds[1] = t2 - t1; io->stage_seconds[2] = t3 - t2;
        io->stage_seconds[3] = t4 - t3; io->stage_seconds[4] = t5 - t4;
    }
}

// trajectory.cpp:111-170, 183-199 ; polynomial.hpp:22-24
void orc_state_at(const orc_params* p, const float* traj, double time, float state[9]) {
    const int M = p->M;
    int n = p->n;
    std::vector<std::vector<vec3f>> cur(M, std::vector<vec3f>(p->n + 1));
    for (int m = 0; m < M; m++)
        for (int i = 0; i <= p->n; i++) cur[m][i] = vec3f(traj + ((size_t)m * (p->n + 1) + i) * 3);
    for (int order = 0; order < 3; order++) {
        // getPointAt
        vec3f point;
        int mm = -1;
        double tn = 0, seg_end = 0;
        for (int idx = 0; idx < M; idx++) {
            seg_end += p->dt;
            if (time < seg_end) { mm = idx; tn = 1 - (seg_end - time) / p->dt; break; }
        }
        if (mm == -1 && time < seg_end + kEpsF) { mm = M - 1; tn = 1.0; }
        if (mm >= 0) {
            for (int i = 0; i < n + 1; i++) {
                double b = n_choose_k(n, i) * std::pow(tn, i) * std::pow(1 - tn, n - i);
                point = point + cur[mm][i] * (float)b;
            }
        }
        point.store(state + 3 * order);
        // derivative :183-199 (control point array keeps its length; only n shrinks)
        for (int m = 0; m < M; m++) {
            for (int i = 0; i < n; i++) cur[m][i] = (cur[m][i + 1] - cur[m][i]) * (float)(n / p->dt);
            cur[m][n] = vec3f();
        }
        n -= 1;
    }
}

void orc_edt_dims(const orc_params* p, int32_t dims[3], int32_t min_key[3]) {
    const double inv = 1.0 / p->world_res;
    for (int k = 0; k < 3; k++) {
        int lo = (int)std::floor(inv * (double)(float)p->world_min[k]);
        int hi = (int)std::floor(inv * (double)(float)p->world_max[k]);
        min_key[k] = lo;
        dims[k] = hi - lo + 1;
    }
}

void orc_edt_build(const orc_params* p, int nb, const float* boxes, float* dist, int32_t* obst) {
    int32_t dims[3], mk[3];
    orc_edt_dims(p, dims, mk);
    const double res = p->world_res, inv = 1.0 / res;
    const int nx = dims[0], ny = dims[1], nz = dims[2];
    const size_t nc = (size_t)nx * ny * nz;
    const int R = 10;                 // only cells with sq dist < 100 can matter (dist < 1 test)
    const int maxd = (int)(1.0 / res + 1);          // DynamicEDTOctomap(maxdist=1.0): int(maxdist/res+1) cells
    const int INF = 1 << 28;
    std::vector<uint8_t> occ(nc, 0);
    auto lin = [&](int x, int y, int z) { return ((size_t)x * ny + y) * nz + z; };
    for (int b = 0; b < nb; b++) {   // map_manager.cpp:285-311
        const float* r = boxes + 6 * b;
        int s[3], e[3];
        for (int k = 0; k < 3; k++) {
            s[k] = (int)std::round((r[k] - 0.5 * r[3 + k]) / res);
            e[k] = (int)std::round((r[k] + 0.5 * r[3 + k]) / res);
        }
        for (int i = s[0]; i < e[0]; i++)
            for (int j = s[1]; j < e[1]; j++)
                for (int k = s[2]; k < e[2]; k++) {
                    float c[3] = {(float)((i + 0.5) * res), (float)((j + 0.5) * res), (float)((k + 0.5) * res)};
                    int m[3];
                    for (int t = 0; t < 3; t++) m[t] = (int)std::floor(inv * (double)c[t]) - mk[t];
                    if (m[0] < 0 || m[0] >= nx || m[1] < 0 || m[1] >= ny || m[2] < 0 || m[2] >= nz) continue;
                    occ[lin(m[0], m[1], m[2])] = 1;
                }
    }
    // exact separable EDT with feature transform; ties -> lowest linear index
    std::vector<int> d1(nc, INF), f1(nc, -1), d2(nc, INF), f2(nc, -1);
    for (int x = 0; x < nx; x++)
        for (int y = 0; y < ny; y++)
            for (int z = 0; z < nz; z++) {
                int best = INF, bf = -1;
                for (int zz = std::max(0, z - R); zz <= std::min(nz - 1, z + R); zz++)
                    if (occ[lin(x, y, zz)]) {
                        int dd = (zz - z) * (zz - z);
                        if (dd < best) { best = dd; bf = (int)lin(x, y, zz); }
                    }
                d1[lin(x, y, z)] = best; f1[lin(x, y, z)] = bf;
            }
    for (int x = 0; x < nx; x++)
        for (int y = 0; y < ny; y++)
            for (int z = 0; z < nz; z++) {
                int best = INF, bf = -1;
                for (int yy = std::max(0, y - R); yy <= std::min(ny - 1, y + R); yy++) {
                    size_t q = lin(x, yy, z);
                    if (f1[q] < 0) continue;
                    int dd = d1[q] + (yy - y) * (yy - y);
                    if (dd < best || (dd == best && f1[q] < bf)) { best = dd; bf = f1[q]; }
                }
                d2[lin(x, y, z)] = best; f2[lin(x, y, z)] = bf;
            }
    for (int x = 0; x < nx; x++)
        for (int y = 0; y < ny; y++)
            for (int z = 0; z < nz; z++) {
                int best = INF, bf = -1;
                for (int xx = std::max(0, x - R); xx <= std::min(nx - 1, x + R); xx++) {
                    size_t q = lin(xx, y, z);
                    if (f2[q] < 0) continue;
                    int dd = d2[q] + (xx - x) * (xx - x);
                    if (dd < best || (dd == best && f2[q] < bf)) { best = dd; bf = f2[q]; }
                }
                size_t c = lin(x, y, z);
                if (bf >= 0 && best < maxd * maxd) {
                    dist[c] = (float)((double)(float)std::sqrt((double)best) * res);
                    int fz = bf % nz, fy = (bf / nz) % ny, fx = bf / (nz * ny);
                    obst[3 * c] = fx; obst[3 * c + 1] = fy; obst[3 * c + 2] = fz;
                } else {
                    dist[c] = (float)((double)(float)maxd * res);
                    obst[3 * c] = obst[3 * c + 1] = obst[3 * c + 2] = -1;
                }
            }
}

void orc_q_base(const orc_params* p, double* Q) { q_base(p, Q); }

}  // extern "C"
