// dlsc_math.cuh -- float32/float64 geometry primitives of the replan hot path.
//
// Every function reproduces the rounding points of the reference (octomap::point3d is float[3];
// GJK and scalars are double) so that LSC/SFC results are bit-identical to the CPU reference path.
// The translation units that include this header for the exact stages are compiled with
// -fmad=false (no FMA contraction), matching the reference's baseline x86-64 build.
//
// Reference being followed (not copied): src/openGJK/openGJK.cpp:138-780 (distance sub-algorithm
// and main loop), include/geometry.hpp:77-112, 139-306 (segment/line closest points),
// octomath::Vector3 semantics (float storage, float dot/norm_sq widened to double).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DLSC_HD __host__ __device__ __forceinline__
#define DLSC_HDN static __host__ __device__ __noinline__
#else
#define DLSC_HD inline
#define DLSC_HDN static inline
#endif

namespace dlsc {

constexpr double kEps = 1e-9;    // SP_EPSILON        include/sp_const.hpp:3
constexpr double kEpsF = 1e-5;   // SP_EPSILON_FLOAT  include/sp_const.hpp:4

struct V3 {
    float x, y, z;
};

DLSC_HD V3 v3(float a, float b, float c) { V3 r; r.x = a; r.y = b; r.z = c; return r; }
DLSC_HD V3 v3_load(const float* p) { return v3(p[0], p[1], p[2]); }
DLSC_HD void v3_store(float* p, const V3& a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
DLSC_HD float v3_get(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
DLSC_HD void v3_set(V3& a, int i, float v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; }
DLSC_HD V3 operator-(const V3& a, const V3& b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
DLSC_HD V3 operator+(const V3& a, const V3& b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
DLSC_HD V3 operator-(const V3& a) { return v3(-a.x, -a.y, -a.z); }
DLSC_HD V3 operator*(const V3& a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
DLSC_HD bool v3_eq(const V3& a, const V3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// float expression widened to double (octomath::Vector3::dot / norm_sq)
DLSC_HD double v3_dot(const V3& a, const V3& b) { float r = a.x * b.x + a.y * b.y + a.z * b.z; return (double)r; }
DLSC_HD double v3_norm_sq(const V3& a) { float r = a.x * a.x + a.y * a.y + a.z * a.z; return (double)r; }
DLSC_HD double v3_norm(const V3& a) { return sqrt(v3_norm_sq(a)); }
// octomath::Vector3::normalized(): len = sqrt((double)norm_sq_float), components divided by (float)len.
// (float)sqrt((double)r) == sqrtf(r) for every float r >= 0 (a correctly rounded double square root rounds to the
// correctly rounded float one: 53 >= 2 * 24 + 2 bits), so the float instruction is used; len > 0 <=> r > 0.
DLSC_HD V3 v3_normalized(const V3& a) {
    V3 r = a;
    const float nsq = a.x * a.x + a.y * a.y + a.z * a.z;
    if (nsq > 0.f) { const float f = sqrtf(nsq); r.x /= f; r.y /= f; r.z /= f; }
    return r;
}
DLSC_HD double v3_distance(const V3& a, const V3& b) {
    double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;   // float subtraction, double accumulate
    return sqrt(dx * dx + dy * dy + dz * dz);
}
DLSC_HD V3 v3_cross(const V3& a, const V3& b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// include/util.hpp:131-140
DLSC_HD double linf_distance(const V3& a, const V3& b) {
    double dist = 0;
    double c = (double)fabsf(a.x - b.x); if (dist < c) dist = c;
    c = (double)fabsf(a.y - b.y); if (dist < c) dist = c;
    c = (double)fabsf(a.z - b.z); if (dist < c) dist = c;
    return dist;
}

// ------------------------------------------------------------------------------------------------
// GJK: closest point of conv{c_0..c_{np-1}} to the origin (openGJK with body 2 = {origin}).
// ------------------------------------------------------------------------------------------------
namespace gjk {

struct D3 { double x, y, z; };
DLSC_HD D3 d3(double a, double b, double c) { D3 r; r.x = a; r.y = b; r.z = c; return r; }
DLSC_HD double dot(const D3& a, const D3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DLSC_HD D3 sub(const D3& a, const D3& b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
DLSC_HD D3 cross(const D3& a, const D3& b) {                 // openGJK.cpp:142-147
    return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
DLSC_HD double det3(const D3& p, const D3& q, const D3& r) {  // openGJK.cpp:138-140
    return p.x * ((q.y * r.z) - (r.y * q.z)) - p.y * (q.x * r.z - r.x * q.z) + p.z * (q.x * r.y - r.x * q.y);
}
DLSC_HD D3 proj_line(const D3& p, const D3& q) {              // openGJK.cpp:149-161
    D3 pq = sub(p, q);
    double t = dot(p, pq) / dot(pq, pq);
    return d3(p.x - pq.x * t, p.y - pq.y * t, p.z - pq.z * t);
}
DLSC_HD D3 proj_plane(const D3& p, const D3& q, const D3& r) {  // openGJK.cpp:163-179
    D3 n = cross(sub(p, q), sub(p, r));
    double t = dot(n, p) / dot(n, n);
    return d3(n.x * t, n.y * t, n.z * t);
}
DLSC_HD int hff1(const D3& p, const D3& q) {                  // openGJK.cpp:181-193
    double t = 0;
    t += (p.x * p.x - p.x * q.x);
    t += (p.y * p.y - p.y * q.y);
    t += (p.z * p.z - p.z * q.z);
    return t > 0 ? 1 : 0;
}
DLSC_HD int hff2(const D3& p, const D3& q, const D3& r) {     // openGJK.cpp:195-218
    D3 pq = sub(q, p), pr = sub(r, p);
    D3 n = cross(pq, cross(pq, pr));
    double t = 0;
    t = t + (p.x * n.x); t = t + (p.y * n.y); t = t + (p.z * n.z);
    return t < 0 ? 1 : 0;
}
DLSC_HD int hff3(const D3& p, const D3& q, const D3& r) {     // openGJK.cpp:220-241
    D3 n = cross(sub(q, p), sub(r, p));
    double t = 0;
    t = t + (p.x * n.x); t = t + (p.y * n.y); t = t + (p.z * n.z);
    return t > 0 ? 0 : 1;
}

// Branch tracer of the distance sub-algorithm: the product kernels pass NoTrace (no code); the per-kernel parity
// entry point dlsc_gjk_batch passes MaskTrace and returns, per hull, the set of decision-tree leaves it went
// through, so the tests can show that every leaf is pinned against the reference's object code.
//   0-1   sub1d: edge region, vertex          2-5   sub2d: face, edge ac, edge ab, vertex (55-56: face / edge ac
//                                                   reached through the "ab not an edge region" side)
//   6     sub3d: vertex a                     7     origin inside the tetrahedron
//   8-10  two visible faces (which one is hidden)                 11-13 one visible face: rotation
//   14-20 one face, one edge      21-29 one face, two edges       30-34 one face, three edges
//   35-37 no face, one edge: rotation, 38-40 its leaves           41-43 no face, two edges: rotation, 44-48 leaves
//   49    no face, three edges (simplex kept)
//   50-54 main loop exits: relative/absolute gap, |v|^2 tiny, |v|^2 vs simplex scale, 4 vertices, 25 iterations
struct NoTrace { DLSC_HD void hit(int) const {} };
struct MaskTrace { unsigned long long m; DLSC_HD void hit(int id) { m |= 1ull << id; } };
constexpr int kGjkLeaves = 57;

// simplex: slots s0..s3 kept as named members so they stay in registers
struct Simplex {
    int n;
    D3 s0, s1, s2, s3;
};

DLSC_HD void set_edge(Simplex& s, const D3& lo, const D3& a) { s.n = 2; s.s0 = lo; s.s1 = a; }
DLSC_HD void set_face(Simplex& s, const D3& lo, const D3& mid, const D3& a) { s.n = 3; s.s0 = lo; s.s1 = mid; s.s2 = a; }

// openGJK.cpp:243-256
template <class Tr>
DLSC_HD D3 sub1d(Simplex& s, Tr& tr) {
    const D3 a = s.s1, b = s.s0;
    if (hff1(a, b)) { tr.hit(0); return proj_line(a, b); }
    tr.hit(1);
    s.n = 1; s.s0 = a;
    return a;
}

// openGJK.cpp:259-313
template <class Tr>
DLSC_HD D3 sub2d(Simplex& s, Tr& tr) {
    const D3 a = s.s2, b = s.s1, c = s.s0;
    const int e_ab = hff1(a, b);
    const int e_ac = hff1(a, c);
    const int f_bc = !hff2(a, b, c);
    const int f_cb = !hff2(a, c, b);
    int r;   // 0 face, 1 edge ab, 2 edge ac, 3 vertex
    if (e_ab) {
        if (f_bc) { r = (e_ac && !f_cb) ? 2 : 0; tr.hit(r ? 3 : 2); }
        else { r = 1; tr.hit(4); }
    } else if (e_ac) {
        r = f_cb ? 0 : 2; tr.hit(r ? 56 : 55);
    } else {
        r = 3; tr.hit(5);
    }
    if (r == 0) return proj_plane(a, b, c);
    if (r == 2) { s.n = 2; s.s1 = a; return proj_line(a, c); }            // keeps {c, a}
    if (r == 1) { s.n = 2; s.s0 = a; return proj_line(a, b); }            // keeps {a, b}
    s.n = 1; s.s0 = a;
    return a;
}

// openGJK.cpp:315-631.  v is the previous search vector (left untouched on the do-nothing paths).
template <class Tr>
DLSC_HDN D3 sub3d(Simplex& s, const D3& v_in, Tr& tr) {
    const D3 a = s.s3;
    const D3 q2 = s.s2, q1 = s.s1, q0 = s.s0;      // q2 = s2, q1 = s3, q0 = s4 of the reference
    const D3 e2 = sub(q2, a), e3 = sub(q1, a), e4 = sub(q0, a);
    const int ed2 = hff1(a, q2), ed1 = hff1(a, q1), ed0 = hff1(a, q0);
    const int n_edge = ed2 + ed1 + ed0;
    if (n_edge == 0) { tr.hit(6); s.n = 1; s.s0 = a; return a; }
    const int sss = det3(e3, e4, e2) > 0 ? 0 : 1;
    int t2 = hff3(a, q1, q0) - sss; t2 *= t2;
    int t3 = hff3(a, q0, q2) - sss; t3 *= t3;
    int t4 = hff3(a, q2, q1) - sss; t4 *= t4;
    const int n_face = t2 + t3 + t4;
    if (n_face == 3) { tr.hit(7); s.n = 4; return d3(0, 0, 0); }
    if (n_face == 2) {
        s.n = 3;
        if (!t2) { tr.hit(8); s.s2 = a; }                               // {s4, s3, a}
        else if (!t3) { tr.hit(9); s.s1 = q2; s.s2 = a; }               // {s4, s2, a}
        else { tr.hit(10); s.s0 = q1; s.s1 = q2; s.s2 = a; }            // {s3, s2, a}
        return sub2d(s, tr);
    }
    // rotation (k, i, j) of the slots (2,1,0)
    D3 si, sj, sk;
    int ei, ej, ek;
    if (n_face == 1) {
        s.n = 3;
        if (t2) { tr.hit(11); sk = q2; si = q1; sj = q0; ek = ed2; ei = ed1; ej = ed0; }
        else if (t3) { tr.hit(12); sk = q1; si = q0; sj = q2; ek = ed1; ei = ed0; ej = ed2; }
        else { tr.hit(13); sk = q0; si = q2; sj = q1; ek = ed0; ei = ed2; ej = ed1; }
        if (n_edge == 1) {
            if (ek) {
                if (!hff2(a, sk, si)) { tr.hit(14); set_face(s, sk, si, a); return proj_plane(a, si, sk); }
                if (!hff2(a, sk, sj)) { tr.hit(15); set_face(s, sk, sj, a); return proj_plane(a, sj, sk); }
                tr.hit(16); set_edge(s, sk, a); return proj_line(a, sk);
            } else if (ei) {
                if (!hff2(a, si, sk)) { tr.hit(17); set_face(s, sk, si, a); return proj_plane(a, si, sk); }
                tr.hit(18); set_edge(s, si, a); return proj_line(a, si);
            } else {
                if (!hff2(a, sj, sk)) { tr.hit(19); set_face(s, sk, sj, a); return proj_plane(a, sj, sk); }
                tr.hit(20); set_edge(s, sj, a); return proj_line(a, sj);
            }
        } else if (n_edge == 2) {
            if (ei) {
                if (!hff2(a, sk, si)) {
                    if (!hff2(a, si, sk)) { tr.hit(21); set_face(s, sk, si, a); return proj_plane(a, si, sk); }
                    tr.hit(22); set_edge(s, sk, a); return proj_line(a, sk);
                } else {
                    if (!hff2(a, sk, sj)) { tr.hit(23); set_face(s, sk, sj, a); return proj_plane(a, sj, sk); }
                    tr.hit(24); set_edge(s, sk, a); return proj_line(a, sk);
                }
            } else if (ej) {
                if (!hff2(a, sk, sj)) {
                    if (!hff2(a, sj, sk)) { tr.hit(25); set_face(s, sk, sj, a); return proj_plane(a, sj, sk); }
                    tr.hit(26); set_edge(s, sj, a); return proj_line(a, sj);
                } else {
                    if (!hff2(a, sk, si)) { tr.hit(27); set_face(s, sk, si, a); return proj_plane(a, si, sk); }
                    tr.hit(28); set_edge(s, sk, a); return proj_line(a, sk);
                }
            }
            tr.hit(29);
            return v_in;    // reference leaves {s4,s3,s2} (n = 3) and v untouched (openGJK.cpp:497-499)
        } else {
            const int d_ik = hff2(a, si, sk), d_jk = hff2(a, sj, sk);
            const int d_ki = hff2(a, sk, si), d_kj = hff2(a, sk, sj);
            if (d_ki == 1 && d_kj == 1) { tr.hit(30); set_edge(s, sk, a); return proj_line(a, sk); }
            if (d_ki) {
                if (d_jk) { tr.hit(31); set_edge(s, sj, a); return proj_line(a, sj); }
                tr.hit(32); set_face(s, sk, sj, a); return proj_plane(a, sk, sj);
            }
            if (d_ik) { tr.hit(33); set_edge(s, si, a); return proj_line(a, si); }
            tr.hit(34); set_face(s, sk, si, a); return proj_plane(a, sk, si);
        }
    }
    // n_face == 0
    if (n_edge == 1) {
        if (ed1) { tr.hit(35); sk = q2; si = q1; sj = q0; }
        else if (ed0) { tr.hit(36); sk = q1; si = q0; sj = q2; }
        else { tr.hit(37); sk = q0; si = q2; sj = q1; }
        if (!hff2(a, si, sj)) { tr.hit(38); set_face(s, sj, si, a); return proj_plane(a, si, sj); }
        if (!hff2(a, si, sk)) { tr.hit(39); set_face(s, sk, si, a); return proj_plane(a, si, sk); }
        tr.hit(40); set_edge(s, si, a); return proj_line(a, si);
    }
    if (n_edge == 2) {
        s.n = 3;
        if (!ed1) { tr.hit(41); sk = q2; si = q1; sj = q0; }
        else if (!ed0) { tr.hit(42); sk = q1; si = q0; sj = q2; }
        else { tr.hit(43); sk = q0; si = q2; sj = q1; }
        if (!hff2(a, sj, sk)) {
            if (!hff2(a, sk, sj)) { tr.hit(44); set_face(s, sk, sj, a); return proj_plane(a, sj, sk); }
            if (!hff2(a, sk, si)) { tr.hit(45); set_face(s, sk, si, a); return proj_plane(a, sk, si); }
            tr.hit(46); set_edge(s, sk, a); return proj_line(a, sk);
        }
        if (!hff2(a, sj, si)) { tr.hit(47); set_face(s, sj, si, a); return proj_plane(a, si, sj); }
        tr.hit(48); set_edge(s, sj, a); return proj_line(a, sj);
    }
    tr.hit(49);
    return v_in;   // n_edge == 3 with no visible face: reference does nothing, n stays 4 (openGJK.cpp:545-626)
}

// openGJK.cpp:674-780 with bd2 = {origin} (include/geometry.hpp:289-298).  NP points c[0..NP).
template <int NP, class Tr>
DLSC_HD D3 hull_origin(const D3 (&c)[NP], int* iters_out, Tr& tr, int* simplex_out = nullptr) {
    const double eps_rel = 1e-10, eps_tot = 1e-12;
    const double eps_rel2 = eps_rel * eps_rel;
    Simplex s;
    D3 v = c[0], sup = c[0];
    s.n = 1; s.s0 = v; s.s1 = v; s.s2 = v; s.s3 = v;
    double nwmax = 0;
    int k = 0;
    do {
        k++;
        const D3 vm = d3(-v.x, -v.y, -v.z);
        double maxs = dot(sup, vm);                         // support, openGJK.cpp:633-655
        int better = -1;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            double sv = dot(c[i], vm);
            if (sv > maxs) { maxs = sv; better = i; }
        }
        if (better != -1) {
#pragma unroll
            for (int i = 0; i < NP; i++) if (i == better) sup = c[i];
        }
        const D3 w = d3(sup.x - 0.0, sup.y - 0.0, sup.z - 0.0);
        const double vv = dot(v, v);
        const double ex = vv - dot(v, w);
        if (ex <= eps_rel * vv || ex < eps_tot) { tr.hit(50); break; }
        if (vv < eps_rel2) { tr.hit(51); break; }
        if (s.n == 1) { s.s1 = w; s.n = 2; v = sub1d(s, tr); }
        else if (s.n == 2) { s.s2 = w; s.n = 3; v = sub2d(s, tr); }
        else { s.s3 = w; s.n = 4; v = sub3d(s, v, tr); }
        double tn = dot(s.s0, s.s0); if (tn > nwmax) nwmax = tn;
        if (s.n > 1) { tn = dot(s.s1, s.s1); if (tn > nwmax) nwmax = tn; }
        if (s.n > 2) { tn = dot(s.s2, s.s2); if (tn > nwmax) nwmax = tn; }
        if (s.n > 3) { tn = dot(s.s3, s.s3); if (tn > nwmax) nwmax = tn; }
        if (dot(v, v) <= eps_tot * eps_tot * nwmax) { tr.hit(52); break; }
        if (s.n == 4) tr.hit(53);
        else if (k == 25) tr.hit(54);
    } while (s.n != 4 && k != 25);
    if (iters_out) *iters_out = k;
    if (simplex_out) *simplex_out = s.n;
    return v;
}
template <int NP>
DLSC_HD D3 hull_origin(const D3 (&c)[NP], int* iters_out) {
    NoTrace tr;
    return hull_origin<NP, NoTrace>(c, iters_out, tr);
}
// The same iteration restricted to simplices of one or two vertices: what the LSC kernel runs for EVERY hull
// (k_lsc phase A).  The swarm's hulls are Bernstein control polygons of short trajectory pieces relative to a
// neighbour -- nearly a point or a needle -- so the closest feature is a vertex or an edge for ~90 % of them and the
// loop below ends through one of openGJK's exit tests after one to three support evaluations, without ever entering
// S2D / S3D.  Operation for operation the prefix of hull_origin (same support order, same exit tests in the same
// order, same S1D), so the returned v is bit-identical.  Returns false, with nothing to rely on, as soon as the
// triangle sub-algorithm would be needed: the caller then runs hull_origin from scratch (phase B).
template <int NP>
DLSC_HD bool hull_origin_short(const D3 (&c)[NP], D3& v_out, int* iters_out) {
    const double eps_rel = 1e-10, eps_tot = 1e-12;
    const double eps_rel2 = eps_rel * eps_rel;
    D3 v = c[0], sup = c[0], s0 = c[0];
    double nwmax = 0;
    int k = 0;
    bool edge = false;                     // the simplex holds two vertices {s0, s1}: only the exit tests are left
    for (;;) {
        k++;
        const D3 vm = d3(-v.x, -v.y, -v.z);
        double maxs = dot(sup, vm);
        int better = -1;
#pragma unroll
        for (int i = 0; i < NP; i++) {
            const double sv = dot(c[i], vm);
            if (sv > maxs) { maxs = sv; better = i; }
        }
        if (better != -1) {
#pragma unroll
            for (int i = 0; i < NP; i++) if (i == better) sup = c[i];
        }
        const D3 w = d3(sup.x - 0.0, sup.y - 0.0, sup.z - 0.0);
        const double vv = dot(v, v);
        const double ex = vv - dot(v, w);
        if (ex <= eps_rel * vv || ex < eps_tot) break;
        if (vv < eps_rel2) break;
        if (edge) return false;            // hull_origin would call sub2d here
        // sub1d on {s0, w}
        double tn;
        if (hff1(w, s0)) {
            v = proj_line(w, s0);
            edge = true;
            tn = dot(s0, s0); if (tn > nwmax) nwmax = tn;
            tn = dot(w, w); if (tn > nwmax) nwmax = tn;
        } else {
            s0 = w; v = w;
            tn = dot(s0, s0); if (tn > nwmax) nwmax = tn;
        }
        if (dot(v, v) <= eps_tot * eps_tot * nwmax) break;
        if (k == 25) break;
    }
    v_out = v;
    if (iters_out) *iters_out = k;
    return true;
}
}  // namespace gjk

// ------------------------------------------------------------------------------------------------
// closest points between two line segments (include/geometry.hpp:77-112, 139-274)
// ------------------------------------------------------------------------------------------------
struct Closest { double dist; V3 p1, p2; };

// geometry.hpp:77-112
DLSC_HD Closest closest_point_segment(const V3& point, const V3& s0, const V3& s1) {
    V3 a = s0 - point, b = s1 - point, rel;
    double dmin;
    if (v3_eq(a, b)) {
        dmin = v3_norm(a); rel = a;
    } else {
        dmin = v3_norm(a); rel = a;
        double dist = v3_norm(b);
        if (dmin > dist) { dmin = dist; rel = b; }
        V3 nl = v3_normalized(b - a);
        V3 c = a - nl * (float)v3_dot(a, nl);
        dist = v3_norm(c);
        if (v3_dot(c - a, c - b) < 0 && dmin > dist) { dmin = dist; rel = c; }
    }
    Closest r; r.dist = dmin; r.p1 = point; r.p2 = rel + point;
    return r;
}

// Eigen::Matrix3f::inverse() * b (cofactor formula; 3x3 * 3x1 product sums x0 + (x1 + x2))
DLSC_HD void solve3f(const float (&m)[3][3], const float (&b)[3], float (&out)[3]) {
#define DLSC_COF(i, j) (m[((i) + 1) % 3][((j) + 1) % 3] * m[((i) + 2) % 3][((j) + 2) % 3] - \
                        m[((i) + 1) % 3][((j) + 2) % 3] * m[((i) + 2) % 3][((j) + 1) % 3])
    const float c00 = DLSC_COF(0, 0), c10 = DLSC_COF(1, 0), c20 = DLSC_COF(2, 0);
    const float det = c00 * m[0][0] + (c10 * m[1][0] + c20 * m[2][0]);
    const float invdet = 1.0f / det;
    float inv[3][3];
    inv[0][0] = c00 * invdet; inv[0][1] = c10 * invdet; inv[0][2] = c20 * invdet;
    inv[1][0] = DLSC_COF(0, 1) * invdet; inv[1][1] = DLSC_COF(1, 1) * invdet; inv[1][2] = DLSC_COF(2, 1) * invdet;
    inv[2][0] = DLSC_COF(0, 2) * invdet; inv[2][1] = DLSC_COF(1, 2) * invdet; inv[2][2] = DLSC_COF(2, 2) * invdet;
#undef DLSC_COF
    for (int r = 0; r < 3; r++) out[r] = inv[r][0] * b[0] + (inv[r][1] * b[1] + inv[r][2] * b[2]);
}

// geometry.hpp:139-182
DLSC_HD Closest closest_lines(const V3& a0, const V3& a1, const V3& b0, const V3& b1) {
    Closest r;
    V3 n1 = v3_normalized(a1 - a0);
    V3 n2 = v3_normalized(b1 - b0);
    if (v3_distance(n1, n2) < kEpsF || v3_distance(n1, -n2) < kEpsF) {
        V3 delta = b0 - a0;
        delta = delta - n1 * (float)(v3_dot(delta, n1));
        r.dist = v3_norm(delta); r.p1 = a0; r.p2 = a0 + delta;
    } else {
        V3 delta = b0 - a0;
        V3 n3 = v3_normalized(v3_cross(n2, n1));
        float A[3][3] = {{n1.x, -n2.x, n3.x}, {n1.y, -n2.y, n3.y}, {n1.z, -n2.z, n3.z}};
        float bb[3] = {delta.x, delta.y, delta.z}, al[3];
        solve3f(A, bb, al);
        r.dist = (double)fabsf(al[2]);
        r.p1 = a0 + n1 * al[0];
        r.p2 = b0 + n2 * al[1];
    }
    return r;
}

// geometry.hpp:184-274
DLSC_HD Closest closest_segments(const V3& a0, const V3& a1, const V3& b0, const V3& b1) {
    Closest cp;
    if (v3_distance(a0, a1) < kEpsF) {
        cp = closest_point_segment(a0, b0, b1);
    } else if (v3_distance(b0, b1) < kEpsF) {
        cp = closest_point_segment(b0, a0, a1);
        V3 t = cp.p1; cp.p1 = cp.p2; cp.p2 = t;
    } else {
        V3 v1 = a1 - a0, v2 = b1 - b0;
        double l1 = v3_norm(v1), l2 = v3_norm(v2);
        V3 n1 = v1 * (float)(1 / l1), n2 = v2 * (float)(1 / l2);
        if (v3_norm(v3_cross(n1, n2)) < kEpsF) {
            double bmin = v3_dot(b0 - a0, n1), bmax = v3_dot(b1 - a0, n1);
            V3 pmin = b0, pmax = b1;
            if (bmax < bmin) { double t = bmin; bmin = bmax; bmax = t; V3 tp = pmin; pmin = pmax; pmax = tp; }
            V3 delta = b0 - a0;
            delta = delta - n1 * (float)(v3_dot(delta, n1));
            if (l1 < bmin) { cp.p1 = a1; cp.p2 = pmin; }
            else if (bmax < 0) { cp.p1 = a0; cp.p2 = pmax; }
            else if (bmin < 0) { cp.p1 = a0; cp.p2 = a0 + delta; }
            else { cp.p1 = pmin - delta; cp.p2 = pmin; }
            cp.dist = v3_distance(cp.p1, cp.p2);
        } else {
            cp = closest_lines(a0, a1, b0, b1);
            double al1 = v3_dot(cp.p1 - a0, n1) / l1;
            double al2 = v3_dot(cp.p2 - b0, n2) / l2;
            if (al1 < 0) cp.p1 = a0; else if (al1 > 1) cp.p1 = a1;
            if (al2 < 0) cp.p2 = b0; else if (al2 > 1) cp.p2 = b1;
            if (al1 < 0 || al1 > 1) {
                double d = v3_dot(n2, cp.p1 - b0);
                if (d < 0) d = 0; else if (d > l2) d = l2;
                cp.p2 = b0 + n2 * (float)d;
            }
            if (al2 < 0 || al2 > 1) {
                double d = v3_dot(n1, cp.p2 - a0);
                if (d < 0) d = 0; else if (d > l1) d = l1;
                cp.p1 = a0 + n1 * (float)d;
            }
            cp.dist = v3_distance(cp.p1, cp.p2);
        }
    }
    return cp;
}

}  // namespace dlsc
