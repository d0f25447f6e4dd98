// dlsc_stages.cuh -- per-agent / per-pair cores of the "exact" stages of the replan hot path:
// horizon shift, neighbour list, LSC, SFC box expansion, goal line search, state step.
//
// The cores are __host__ __device__ so that the same source runs (a) inside the sm_100a kernels
// of dlsc_kernels_exact.cu and (b) in the test-only host simulator (tests/hostsim), which lets
// the arithmetic be checked against the oracle without a GPU.  Cooperative cores take a Group
// (one warp on the device, a single lane on the host).
//
// Must be compiled with FMA contraction off (-fmad=false / -ffp-contract=off).
#pragma once
#include <string.h>

#include "dlsc_math.cuh"
#include "dlsc_types.h"

namespace dlsc {

// ------------------------------------------------------------------------------------------------
// cooperative group abstraction: a warp on the device, one lane in the host simulator
// ------------------------------------------------------------------------------------------------
struct Group {
    int lane, width;
    bool block;      // device: true = the whole CTA cooperates (barrier votes), false = one warp
    DLSC_HD bool any(bool p) const {
#ifdef __CUDA_ARCH__
        return block ? (__syncthreads_or(p ? 1 : 0) != 0) : (__any_sync(0xffffffffu, p) != 0);
#else
        return p;
#endif
    }
    DLSC_HD void sync() const {
#ifdef __CUDA_ARCH__
        if (block) __syncthreads(); else __syncwarp();
#endif
    }
    DLSC_HD unsigned reduce_or(unsigned v) const {   // OR over the group, every member gets the result (+ barrier)
#ifdef __CUDA_ARCH__
        if (!block) return __reduce_or_sync(0xffffffffu, v);
        unsigned r = 0;
        for (int bit = 0; bit < 8; bit++) if (__syncthreads_or((v >> bit) & 1u)) r |= 1u << bit;
        return r;
#else
        return v;
#endif
    }
    DLSC_HD double max(double v) const {         // warp mode only; order-independent, exact
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, v, o); v = (v < t) ? t : v; }
#endif
        return v;
    }
    DLSC_HD unsigned ballot(bool p) const {     // warp mode only
#ifdef __CUDA_ARCH__
        return __ballot_sync(0xffffffffu, p);
#else
        return p ? 1u : 0u;
#endif
    }
};
DLSC_HD int popc_u32(unsigned v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// ------------------------------------------------------------------------------------------------
// horizon shift (traj_planner.cpp:290-336, 409-441 ; trajectory.cpp:79-91)
// one call = one control point `pt` (= m*P + i) of one agent
// ------------------------------------------------------------------------------------------------
DLSC_HD V3 shifted_point(const DevParams& P, const float* rec, int seq, int pt) {
    const V3 pos = v3_load(rec + P.M * kP * 3), vel = v3_load(rec + P.M * kP * 3 + 3);
    if (seq < 2) return pos + vel * P.tk[pt];                                  // planConstVelTraj
    const int m = pt / kP, i = pt - m * kP;
    const int src = (m == P.M - 1) ? ((P.M - 1) * kP + (kP - 1)) : ((m + 1) * kP + i);   // :304-314
    return v3_load(rec + src * 3);
}

// pred: how the other agents see this one (checkObstacleDisturbance :329-336)
// init: the agent's own initial trajectory (initialTrajPlanningCheck :435-441)
DLSC_HD void predict_point(const DevParams& P, const float* rec, int seq, int pt, bool disturbed,
                           float* pred_out, float* init_out) {
    const V3 pos = v3_load(rec + P.M * kP * 3);
    const V3 q = shifted_point(P, rec, seq, pt);
    const V3 first = shifted_point(P, rec, seq, 0);
    const bool reset = v3_norm(first - pos) > P.reset_threshold;
    if (pred_out) v3_store(pred_out + pt * 3, reset ? (pos + v3(0.f, 0.f, 0.f) * P.tk[pt]) : q);
    if (init_out) v3_store(init_out + pt * 3, disturbed ? (pos + v3(0.f, 0.f, 0.f) * P.tk[pt]) : q);
}

// ------------------------------------------------------------------------------------------------
// neighbour list by Chebyshev range, ascending index (multi_sync_simulator.cpp:481-503)
// ------------------------------------------------------------------------------------------------
DLSC_HD bool in_comm_range(const DevParams& P, const V3& pa, const V3& pj) {
    const double dist = linf_distance(pa, pj);
    return !(P.comm_range > 0 && dist > P.comm_range);
}
DLSC_HD int neighbours_agent(const Group& g, const DevParams& P, const float* rec, int a_global,
                             int32_t* idx_out) {
    const int off = P.M * kP * 3;
    const V3 pa = v3_load(rec + (size_t)a_global * P.rec + off);
    int count = 0;
    for (int base = 0; base < P.N; base += g.width) {
        const int j = base + g.lane;
        bool in = false;
        if (j < P.N && j != a_global && rec[(size_t)j * P.rec + off + 11] == rec[(size_t)a_global * P.rec + off + 11]) {
            in = in_comm_range(P, pa, v3_load(rec + (size_t)j * P.rec + off));
        }
        const unsigned mask = g.ballot(in);
        const int pos = count + popc_u32(mask & ((1u << g.lane) - 1u));
        if (in && pos < P.K - P.n_dyn) idx_out[pos] = j;                      // idx_out points past the dynamic slots
        count += popc_u32(mask);
    }
    return count;
}

// smallest normalised slack of an item as stored for the QP row screen (float, +inf for a zero normal: such rows
// are not constraints, traj_optimizer.cpp:422-424)
DLSC_HD float lsc_item_slack(double smin, double nn) {
    if (!(nn > 0.0)) return 3.0e38f;
    // a screen value, not a result: float arithmetic (1e-7 relative) is far inside the screen's 1e-3 m margin
    const float q = (float)smin / (float)nn;
    return q > 3.0e38f ? 3.0e38f : (q < -3.0e38f ? -3.0e38f : q);
}
// the 6 control points of one segment (18 floats, 8-byte aligned: 72-byte segments of 16-byte aligned trajectories)
DLSC_HD void load_segment(const float* p, V3 (&out)[kP]) {
#ifdef __CUDA_ARCH__
    const float2* q = reinterpret_cast<const float2*>(p);
    float f[kP * 3];
#pragma unroll
    for (int i = 0; i < kP * 3 / 2; i++) { const float2 t = q[i]; f[2 * i] = t.x; f[2 * i + 1] = t.y; }
#pragma unroll
    for (int i = 0; i < kP; i++) out[i] = v3(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
#else
    for (int i = 0; i < kP; i++) out[i] = v3_load(p + 3 * i);
#endif
}
DLSC_HD bool is_pow2_f32(float f) {           // normal, positive, zero mantissa
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    memcpy(&u, &f, 4);
#endif
    return (u & 0x807fffffu) == 0 && ((u >> 23) - 1u) < 253u;
}
// ------------------------------------------------------------------------------------------------
// LSC for one (agent, neighbour, segment) -- traj_planner.cpp:603-666, 1102-1127, 1150-1161
//   init_a : own initial trajectory [M][P][3], pred_j : neighbour's predicted trajectory
//   r_j / dw_j arrive as float (agent_manager.cpp:256-258)
//   normal_out[3], d_out[P]; anchor_last_out[3] written only for m == M-1
// A GJK item (m < M-1) is load -> hull vs origin -> finish; the kernel runs the cheap two-vertex GJK for every item
// and queues the few that need the triangle / tetrahedron sub-algorithms (k_lsc / k_lsc_rest), lsc_segment does both
// in place (host simulator, subset stepping).
// ------------------------------------------------------------------------------------------------
// per (agent, neighbour) constants; inv_* are the exact reciprocals when the value is a power of two (every mission with
// equal downwash coefficients: 2.0), else 0: x / 2^k == x * 2^-k bit for bit, the IEEE division otherwise
struct LscPair { double collision_dist, downwash, inv_downwash; float dwf, inv_dwf; };
DLSC_HD bool is_pow2_f64(double d) {           // normal, positive, zero mantissa, reciprocal representable
    unsigned long long u;
#ifdef __CUDA_ARCH__
    u = (unsigned long long)__double_as_longlong(d);
#else
    memcpy(&u, &d, 8);
#endif
    return (u & 0x800fffffffffffffull) == 0 && ((u >> 52) - 2ull) < 2043ull;
}
DLSC_HD LscPair lsc_pair_consts(double r_a, double dw_a, float r_j_f, float dw_j_f) {
    const double r_j = (double)r_j_f, dw_j = (double)dw_j_f;
    LscPair q;
    q.collision_dist = r_j + r_a;                                              // :605
    q.downwash = (dw_a * r_a + dw_j * r_j) / (r_a + r_j);                      // :1153-1154
    q.dwf = (float)q.downwash;                                                 // trajectory.cpp:214
    q.inv_dwf = is_pow2_f32(q.dwf) ? 1.0f / q.dwf : 0.0f;
    q.inv_downwash = is_pow2_f64(q.downwash) ? 1.0 / q.downwash : 0.0;
    return q;
}
// (float)((double)z / downwash)  (:630-632, 1183-1187)
DLSC_HD float lsc_div_downwash(float z, const LscPair& q) {
    return (float)(q.inv_downwash != 0.0 ? (double)z * q.inv_downwash : (double)z / q.downwash);
}
// hull points c[i] = (double)(a_i - b_i) in the downwash frame; (float)c[i] recovers the float difference exactly
DLSC_HD void lsc_gjk_load(const float* init_seg, const float* pred_seg, const LscPair& q, gjk::D3 (&c)[kP]) {
    V3 wb[kP];
    load_segment(pred_seg, wb);
    const bool pow2 = q.inv_dwf != 0.0f;
    const float inv_dwf = q.inv_dwf;
#pragma unroll
    for (int i = 0; i < kP; i++) {
        V3 a = v3_load(init_seg + i * 3); a.z = pow2 ? a.z * inv_dwf : a.z / q.dwf;
        V3 b = wb[i]; b.z = pow2 ? b.z * inv_dwf : b.z / q.dwf;
        const V3 rel = a - b;
        c[i] = gjk::d3((double)rel.x, (double)rel.y, (double)rel.z);   // util.hpp:113-125
    }
}
DLSC_HD void lsc_gjk_finish(const gjk::D3 (&c)[kP], const gjk::D3& v, const LscPair& q, int m, float* normal_out,
                            double* d_out, float* slack_out) {
    const V3 cp2 = v3(0.f, 0.f, 0.f) + v3((float)v.x, (float)v.y, (float)v.z);   // geometry.hpp:302
    const V3 nt = v3_normalized(cp2);                                             // :1118
    const float nz_w = lsc_div_downwash(nt.z, q);                                 // :630-632
    normal_out[0] = nt.x; normal_out[1] = nt.y; normal_out[2] = nz_w;
    double pmin = 1e300;
#pragma unroll
    for (int i = 0; i < kP; i++) {                                                // :636-637
        const double pi = v3_dot(v3((float)c[i].x, (float)c[i].y, (float)c[i].z), nt);
        d_out[i] = 0.5 * (q.collision_dist + pi);
        if (m > 0 || i >= 3) pmin = (pi < pmin) ? pi : pmin;                      // rows of (m = 0, i < 3) do not exist
    }
    if (slack_out) {
        // QP row screen (dlsc_qp_gi.cuh): smallest slack of the rows of this item at the agent's own initial
        // trajectory, normalised by the world-frame |normal|: a row with normalised slack s cannot be violated
        // by any x with |x_pt - init_pt| < s.  The slack of row i at the initial trajectory is
        // n_w.(a_i - b_i) - d_i = p_i - (collision_dist + p_i) / 2 with p_i the dot product above (the world-frame
        // normal against world-frame points equals the scaled normal against scaled points), so it costs nothing;
        // float rounding (1e-6 m) is far inside the screen's margin (1e-3 m).
        const float nn = sqrtf(nt.x * nt.x + nt.y * nt.y + nz_w * nz_w);
        *slack_out = lsc_item_slack(0.5 * (pmin - q.collision_dist), (double)nn);
    }
}
// last segment: goal-directed line segments, one LSC for all control points (:640-660)
DLSC_HD void lsc_last_segment(const DevParams& P, const float* init_a, const float* pred_j, const V3& goal_a, const V3& goal_j,
                              const LscPair& q, float* normal_out, double* d_out, float* anchor_last_out, float* slack_out) {
    const float dwf = q.dwf;
    const double downwash = q.downwash;
    const int last = (P.M - 1) * kP + (kP - 1);
    V3 o_last = v3_load(pred_j + last * 3); o_last.z = o_last.z / dwf;
    V3 a_last = v3_load(init_a + last * 3); a_last.z = a_last.z / dwf;
    V3 og = goal_j; og.z = (float)((double)og.z / downwash);                      // :1183-1187
    V3 ag = goal_a; ag.z = (float)((double)ag.z / downwash);
    const Closest cp = closest_segments(o_last, og, a_last, ag);                  // :642-644
    const V3 nt = v3_normalized(cp.p2 - cp.p1);
    const double dd = 0.5 * (q.collision_dist + cp.dist);                         // :650
    const float nz_w = (float)((double)nt.z / downwash);
    const float al_z = (float)((double)cp.p1.z * downwash);                       // :657
    normal_out[0] = nt.x; normal_out[1] = nt.y; normal_out[2] = nz_w;
    anchor_last_out[0] = cp.p1.x; anchor_last_out[1] = cp.p1.y; anchor_last_out[2] = al_z;
#pragma unroll
    for (int i = 0; i < kP; i++) d_out[i] = dd;
    if (slack_out) {      // same screen; rows n_w.(a_i - anchor) >= dd, evaluated in the scaled frame in float
        const float inv = 1.0f / dwf;
        float smin = 3.0e38f;
#pragma unroll
        for (int i = 0; i < kP; i++) {
            const float* a = init_a + ((P.M - 1) * kP + i) * 3;
            const float sl = (a[0] - cp.p1.x) * nt.x + (a[1] - cp.p1.y) * nt.y + (a[2] * inv - cp.p1.z) * nt.z;
            smin = (sl < smin) ? sl : smin;
        }
        const float nn = sqrtf(nt.x * nt.x + nt.y * nt.y + nz_w * nz_w);
        *slack_out = lsc_item_slack((double)smin - dd, (double)nn);
    }
}
DLSC_HD void lsc_segment(const DevParams& P, const float* init_a, const float* pred_j, const V3& goal_a,
                         const V3& goal_j, double r_a, double dw_a, float r_j_f, float dw_j_f, int m,
                         float* normal_out, double* d_out, float* anchor_last_out, int* gjk_iters,
                         float* slack_out = nullptr) {
    const LscPair q = lsc_pair_consts(r_a, dw_a, r_j_f, dw_j_f);
    if (m < P.M - 1) {
        gjk::D3 c[kP], v;
        lsc_gjk_load(init_a + m * kP * 3, pred_j + m * kP * 3, q, c);
        int it = 0;
        if (!gjk::hull_origin_short<kP>(c, v, &it)) v = gjk::hull_origin<kP>(c, &it);
        if (gjk_iters) *gjk_iters = it;
        lsc_gjk_finish(c, v, q, m, normal_out, d_out, slack_out);
    } else {
        lsc_last_segment(P, init_a, pred_j, goal_a, goal_j, q, normal_out, d_out, anchor_last_out, slack_out);
        if (gjk_iters) *gjk_iters = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// SFC -- collision_constraints.cpp:435-452, 502-536, 781-901, 1023-1093
// ------------------------------------------------------------------------------------------------
struct Box { V3 lo, hi; };
DLSC_HD Box box_load(const float* s) { Box b; b.lo = v3_load(s); b.hi = v3_load(s + 3); return b; }
DLSC_HD void box_store(float* s, const Box& b) { v3_store(s, b.lo); v3_store(s + 3, b.hi); }

DLSC_HD bool point_in_box(const Box& b, const V3& q) {     // collision_constraints.cpp:109-116
    return q.x > b.lo.x - kEpsF && q.y > b.lo.y - kEpsF && q.z > b.lo.z - kEpsF &&
           q.x < b.hi.x + kEpsF && q.y < b.hi.y + kEpsF && q.z < b.hi.z + kEpsF;
}
DLSC_HD bool box_includes(const Box& b, const Box& o) { return point_in_box(b, o.lo) && point_in_box(b, o.hi); }
DLSC_HD bool superset_of_hull(const Box& b, const V3* pts, int np) {   // :163-178
    for (int i = 0; i < 3; i++) {
        float mn = v3_get(pts[0], i), mx = mn;
        for (int k = 1; k < np; k++) {
            const float c = v3_get(pts[k], i);
            mn = (c < mn) ? c : mn;
            mx = (mx < c) ? c : mx;
        }
        if (mn < v3_get(b.lo, i) - kEpsF || mx > v3_get(b.hi, i) + kEpsF) return false;
    }
    return true;
}

// DynamicEDTOctomap::getDistanceAndClosestObstacle + the per-vertex test of isObstacleInSFC (:876-886)
DLSC_HD bool vertex_blocked(const EdtDev& E, const V3& q, double margin, float half_res) {
    const int cx = (int)floor(E.inv_res * (double)q.x) - E.min_key[0];
    const int cy = (int)floor(E.inv_res * (double)q.y) - E.min_key[1];
    const int cz = (int)floor(E.inv_res * (double)q.z) - E.min_key[2];
    float dist = -1.0f;
    V3 cl = v3(0.f, 0.f, 0.f);
    if (cx >= 0 && cx < E.dims[0] && cy >= 0 && cy < E.dims[1] && cz >= 0 && cz < E.dims[2]) {
        const size_t idx = ((size_t)cx * E.dims[1] + cy) * E.dims[2] + cz;
#ifdef __CUDA_ARCH__
        const int4 rec = __ldg(E.cells + idx);
#else
        const int4 rec = E.cells[idx];
#endif
#ifdef __CUDA_ARCH__
        dist = __int_as_float(rec.x);
#else
        { union { int i; float f; } u; u.i = rec.x; dist = u.f; }
#endif
        if (rec.y >= 0) {
            cl.x = (float)(((double)(rec.y + E.min_key[0]) + 0.5) * E.res);
            cl.y = (float)(((double)(rec.z + E.min_key[1]) + 0.5) * E.res);
            cl.z = (float)(((double)(rec.w + E.min_key[2]) + 0.5) * E.res);
        }
    }
    if (!(dist < 1)) return false;
    const V3 d3v = v3(half_res, half_res, half_res);
    const V3 lo = cl - d3v, hi = cl + d3v;
    V3 cq = q;                                                  // Box::closestPoint :226-237
    if (q.x < lo.x) cq.x = lo.x; else if (q.x > hi.x) cq.x = hi.x;
    if (q.y < lo.y) cq.y = lo.y; else if (q.y > hi.y) cq.y = hi.y;
    if (q.z < lo.z) cq.z = lo.z; else if (q.z > hi.z) cq.z = hi.z;
    return linf_distance(cq, q) < margin + kEpsF;
}

// isObstacleInSFC (:862-892): any lattice vertex of the box blocked?
// Per call the cooperating threads first tabulate, per axis, the float lattice coordinates of the box and
// the grid cell each falls into (exactly the reference's float -> double -> floor arithmetic, done once per
// coordinate instead of once per vertex); the vertex loop is then integer indexing + one 16-byte load.
// Note: no vertex may be skipped because "it was tested before" -- coordinates regenerated from a moved
// box corner differ in the last float bit about half the time and sit exactly on cell boundaries, so the
// reference's decision depends on that bit; the whole sequence of tests is reproduced.
constexpr int kSfcTabMax = 96;
constexpr int kSfcUnroll = 2;
constexpr int kSfcZsMax = 256;
struct alignas(16) SfcTab {
    float q[3][kSfcTabMax];            // record path: lattice coordinates / cells per axis
    int c[3][kSfcTabMax];
    // mask path: per-axis entries, cached across the box tests of one expansion (keyed by the float bits of the
    // axis' lo / hi: consecutive tests differ in one axis only)
    alignas(16) uint32_t zm[kSfcZsMax / 4];     // per z-vertex nibble select (0x0F / 0xF0), 0 outside the box
    int mc[3][kSfcTabMax];             // v << 3 | near << 2 | oob << 1 | s
    uint32_t klo[3], khi[3];           // keys of the per-axis scalars km / kv
    int km[3];                         // vertices per axis (or <= 0), -1000 = slot empty
    int kv[3];                         // lattice index of the first vertex (relative to min_key); kNoLattice = off-lattice
    uint32_t tlo[3], thi[3];           // keys of the table content mc[ax] / zm
    int tm[3];                         // entries in mc[ax], -1000 = table empty
    int kflag[3];                      // bit 0: axis unusable (off-lattice / ambiguous), z only: bit 1 near(any), bit 2 near & oob (any)
};
constexpr int kNoLattice = -(1 << 30);
DLSC_HD void sfc_tab_reset(const Group& g, SfcTab* t) {
    if (g.lane < 3) { t->km[g.lane] = -1000; t->tm[g.lane] = -1000; }
    g.sync();
}
DLSC_HD uint32_t f32_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
// exact n / d for n < 2^16, 2 <= d < 2^16 with magic = floor(2^32 / d) + 1 (d == 1: magic = 0 -> n)
// (32-bit arithmetic: floor(2^32 / d) = floor((2^32 - 1) / d) unless d is a power of two, where the smaller
// magic 2^32 / d is exact as well)
DLSC_HD uint32_t fastdiv_magic(uint32_t d) {
    if (d <= 1) return 0u;
    const uint32_t q = 0xffffffffu / d;
    return q + 1u;
}
DLSC_HD uint32_t fastdiv(uint32_t n, uint32_t magic) {
#ifdef __CUDA_ARCH__
    return magic ? __umulhi(n, magic) : n;
#else
    return magic ? (uint32_t)(((uint64_t)n * magic) >> 32) : n;
#endif
}

DLSC_HD bool vertex_blocked_tab(const EdtDev& E, float qx, float qy, float qz, int cx, int cy, int cz,
                                double margin, float half_res) {
    float dist = -1.0f;
    V3 cl = v3(0.f, 0.f, 0.f);
    if (cx >= 0 && cy >= 0 && cz >= 0) {
        const size_t idx = ((size_t)cx * E.dims[1] + cy) * E.dims[2] + cz;
#ifdef __CUDA_ARCH__
        const int4 rec = __ldg(E.cells + idx);
        dist = __int_as_float(rec.x);
#else
        const int4 rec = E.cells[idx];
        { union { int i; float f; } u; u.i = rec.x; dist = u.f; }
#endif
        if (!(dist < 1)) return false;
        if (rec.y >= 0) { cl.x = E.centre[0][rec.y]; cl.y = E.centre[1][rec.z]; cl.z = E.centre[2][rec.w]; }
    }
    const V3 q = v3(qx, qy, qz);
    const V3 d3v = v3(half_res, half_res, half_res);
    const V3 lo = cl - d3v, hi = cl + d3v;
    V3 cq = q;                                                  // Box::closestPoint :226-237
    if (q.x < lo.x) cq.x = lo.x; else if (q.x > hi.x) cq.x = hi.x;
    if (q.y < lo.y) cq.y = lo.y; else if (q.y > hi.y) cq.y = hi.y;
    if (q.z < lo.z) cq.z = lo.z; else if (q.z > hi.z) cq.z = hi.z;
    return linf_distance(cq, q) < margin + kEpsF;
}

// ---- lattice-vertex mask -------------------------------------------------------------------------
// Every vertex the reference tests is a point of the res-lattice (box corners are grid-aligned, :440-441,
// :803-806, :840-852, :1058-1063) whose float coordinate sits within a few ulp of a cell boundary, so
// floor(coord / res) picks either cell v or cell v-1 on each axis -- that bit is reproduced exactly per
// coordinate (per-axis tables below).  Given the cell, the vertex test (:876-886) compares an L-infinity
// distance that is a multiple of res (+- float noise) with margin + 1e-5; when no such multiple is within
// kMaskGuard of the threshold the outcome is a function of (lattice vertex, cell choice) only and is
// tabulated once per grid: 8 outcomes = one byte per vertex instead of a 16-byte record per test.
// edt_vertex_mask reports any decision closer than kMaskGuard to the threshold; the mask is then not used.
constexpr double kMaskGuard = 1e-3;
DLSC_HD int edt_mask_zs(int dims2) { return ((dims2 + 1) + 15) / 16 * 16; }

DLSC_HD uint8_t edt_vertex_mask(const EdtDev& E, int vx, int vy, int vz, double margin, bool* unsafe) {
    const float half_res = (float)(0.5 * E.res);
    const float qx = (float)((double)(vx + E.min_key[0]) * E.res);
    const float qy = (float)((double)(vy + E.min_key[1]) * E.res);
    const float qz = (float)((double)(vz + E.min_key[2]) * E.res);
    const V3 q = v3(qx, qy, qz);
    const V3 d3v = v3(half_res, half_res, half_res);
    uint8_t bits = 0;
    for (int s = 0; s < 8; s++) {
        const int cx = vx - (s & 1), cy = vy - ((s >> 1) & 1), cz = vz - (s >> 2);
        if (cx < 0 || cx >= E.dims[0] || cy < 0 || cy >= E.dims[1] || cz < 0 || cz >= E.dims[2]) continue;
        const size_t idx = ((size_t)cx * E.dims[1] + cy) * E.dims[2] + cz;
        const int4 rec = E.cells[idx];
        float dist;
#ifdef __CUDA_ARCH__
        dist = __int_as_float(rec.x);
#else
        { union { int i; float f; } u; u.i = rec.x; dist = u.f; }
#endif
        if (!(dist < 1)) continue;
        V3 cl = v3(0.f, 0.f, 0.f);
        if (rec.y >= 0) { cl.x = E.centre[0][rec.y]; cl.y = E.centre[1][rec.z]; cl.z = E.centre[2][rec.w]; }
        const V3 lo = cl - d3v, hi = cl + d3v;
        V3 cq = q;
        if (q.x < lo.x) cq.x = lo.x; else if (q.x > hi.x) cq.x = hi.x;
        if (q.y < lo.y) cq.y = lo.y; else if (q.y > hi.y) cq.y = hi.y;
        if (q.z < lo.z) cq.z = lo.z; else if (q.z > hi.z) cq.z = hi.z;
        const double li = linf_distance(cq, q), thr = margin + kEpsF;
        if (li < thr) bits |= (uint8_t)(1u << s);
        if (fabs(li - thr) < kMaskGuard) *unsafe = true;
    }
    return bits;
}

struct U4 { uint32_t x, y, z, w; };

// ---- summed-area table over "mask byte != 0" ---------------------------------------------------
// sat[i][j][k] = number of flagged lattice vertices with vx < i, vy < j, vz < k   (dims (nv+1) per axis,
// nv = dims+1 vertices).  A vertex is flagged when any of its 8 cell choices is blocked, or when it lies on
// the outermost vertex layer of the grid close to the world origin: there a cell choice can fall outside
// the grid, where the accessor's error return makes (0,0,0) the "closest obstacle" (see obstacle_in_box_mask).
DLSC_HD size_t sat_index(const EdtDev& E, int i, int j, int k) {
    return ((size_t)i * (E.dims[1] + 2) + j) * (E.dims[2] + 2) + k;
}
DLSC_HD int sat_indicator(const EdtDev& E, int vx, int vy, int vz) {
    if (E.vmask[((size_t)vx * (E.dims[1] + 1) + vy) * E.zs + vz]) return 1;
    const bool edge = vx == 0 || vx == E.dims[0] || vy == 0 || vy == E.dims[1] || vz == 0 || vz == E.dims[2];
    if (!edge) return 0;
    const double lim = 0.5 + 3.0 * E.res;                  // generous: the predicate needs |coordinate| < res/2 + margin
    const double x = (double)(vx + E.min_key[0]) * E.res, y = (double)(vy + E.min_key[1]) * E.res,
                 z = (double)(vz + E.min_key[2]) * E.res;
    return (fabs(x) <= lim && fabs(y) <= lim && fabs(z) <= lim) ? 1 : 0;
}
// flagged vertices in the inclusive lattice range; lanes 0..7 fetch one corner each
DLSC_HD int sat_count(const Group& g, const EdtDev& E, int x0, int x1, int y0, int y1, int z0, int z1) {
#ifdef __CUDA_ARCH__
    if (!g.block) {
        int v = 0;
        if (g.lane < 8) {
            const int i = (g.lane & 1) ? x0 : x1 + 1, j = (g.lane & 2) ? y0 : y1 + 1, k = (g.lane & 4) ? z0 : z1 + 1;
            const int t = __ldg(E.sat + sat_index(E, i, j, k));
            v = (__popc(g.lane) & 1) ? -t : t;
        }
        return __reduce_add_sync(0xffffffffu, v);
    }
#endif
    int s = 0;
    for (int c = 0; c < 8; c++) {
        const int i = (c & 1) ? x0 : x1 + 1, j = (c & 2) ? y0 : y1 + 1, k = (c & 4) ? z0 : z1 + 1;
        const int t = E.sat[sat_index(E, i, j, k)];
        s += (popc_u32((unsigned)c) & 1) ? -t : t;
    }
    return s;
}

// Lattice description of one axis of a box [lo, hi]: vertex count mm = floor((hi - lo + eps) / res) + 1 (the
// reference's loop bound, collision_constraints.cpp:868-872) and the lattice index vv of lo relative to min_key
// (kNoLattice when lo is not within 1e-2 cells of a lattice vertex).
DLSC_HD void lattice_axis(const EdtDev& E, double res, int ax, float lo, float hi, int& mm, int& vv) {
    // floor(x / res): the quotient sits ~eps/res = 1e-4 above an integer, so the product with 1/res has
    // the same floor unless it lands within 1e-9 of an integer; only then the IEEE division decides
    const double x = (hi - lo) + kEpsF, q = x * E.inv_res, f = floor(q);
    mm = ((E.res == res && q - f > 1e-9 && q - f < 1.0 - 1e-9) ? (int)f : (int)floor(x / res)) + 1;
    const double t = E.inv_res * (double)lo, vr = rint(t);
    vv = (fabs(t - vr) > 1e-2) ? kNoLattice : (int)vr - E.min_key[ax];
}

// isObstacleInSFC through the vertex mask.  Returns 0 / 1, or -1 when this box cannot use the mask (a corner
// off the lattice, ambiguous cell choice, oversized) -> the caller runs the record path.
// One work item = 16 consecutive z-vertices of one (x, y) lattice column: one 16-byte load.
// A coordinate outside the grid ("oob", e.g. z = -1e-9 under the world floor) makes the accessor return
// dist = -1 and obstacle (0,0,0) (vertex_blocked_tab): the test is then the separable predicate
// near_x & near_y & near_z with near = (|q - clamp(q, +-res/2)| < margin + 1e-5), evaluated per axis here.
// Per-axis table entry: v << 3 | near << 2 | oob << 1 | s   (v = lattice index, s = v - cell in {0,1}).
DLSC_HD int obstacle_in_box_mask(const Group& g, const DevParams& P, const EdtDev& E, const Box& b, double margin,
                                 SfcTab* tab, long long* lookups) {
    const double res = P.world_res;
    const float half_res = (float)(0.5 * res);
    const double thr = margin + kEpsF;
    // ---- per-axis scalars (vertex count, first lattice index), cached: consecutive tests differ in one axis ----
    int mm[3], vv[3];
    bool smiss = false;
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
        const float lo = v3_get(b.lo, ax), hi = v3_get(b.hi, ax);
        if (tab->km[ax] != -1000 && tab->klo[ax] == f32_bits(lo) && tab->khi[ax] == f32_bits(hi)) {
            mm[ax] = tab->km[ax]; vv[ax] = tab->kv[ax];
        } else {
            smiss = true;
            lattice_axis(E, res, ax, lo, hi, mm[ax], vv[ax]);
        }
    }
    if (smiss) {
        g.sync();                                       // every lane has read the old keys
        if (g.lane < 3) {
            const int ax = g.lane;
            tab->klo[ax] = f32_bits(v3_get(b.lo, ax)); tab->khi[ax] = f32_bits(v3_get(b.hi, ax));
            tab->km[ax] = (ax == 0) ? mm[0] : (ax == 1 ? mm[1] : mm[2]);
            tab->kv[ax] = (ax == 0) ? vv[0] : (ax == 1 ? vv[1] : vv[2]);
        }
        g.sync();
    }
    if (mm[0] <= 0 || mm[1] <= 0 || mm[2] <= 0) return 0;
    if (mm[0] > kSfcTabMax || mm[1] > kSfcTabMax || mm[2] > kSfcTabMax) return -1;
    if (lookups && g.lane == 0) {
        lookups[0] += (long long)mm[0] * mm[1] * mm[2];
        if (lookups[5]) lookups[4] += (long long)mm[0] * mm[1] * mm[2];
    }
    // ---- O(1) emptiness query: the summed-area table counts the lattice vertices whose mask byte is non-zero
    //      (any cell choice blocked); a box whose vertex range holds none is free whatever the float bits say ----
    if (E.sat && vv[0] >= 0 && vv[1] >= 0 && vv[2] >= 0 && vv[0] + mm[0] - 1 <= E.dims[0] &&
        vv[1] + mm[1] - 1 <= E.dims[1] && vv[2] + mm[2] - 1 <= E.dims[2]) {
        if (sat_count(g, E, vv[0], vv[0] + mm[0] - 1, vv[1], vv[1] + mm[1] - 1, vv[2], vv[2] + mm[2] - 1) == 0) {
            if (lookups && g.lane == 0) lookups[3] += 1;
            return 0;
        }
    }
    // ---- exact path: per-axis entry tables (rebuilt for the axes whose table is stale) ----
    bool miss[3];
#pragma unroll
    for (int ax = 0; ax < 3; ax++)
        miss[ax] = !(tab->tm[ax] != -1000 && tab->tlo[ax] == f32_bits(v3_get(b.lo, ax)) && tab->thi[ax] == f32_bits(v3_get(b.hi, ax)));
    if (miss[0] || miss[1] || miss[2]) {
        g.sync();
        const int n0 = miss[0] ? mm[0] : 0, n1 = miss[1] ? mm[1] : 0, n2 = miss[2] ? mm[2] : 0;
        bool bad[3] = {false, false, false};
        bool nz_all = false, nz_oob = false;
        for (int e = g.lane; e < n0 + n1 + n2; e += g.width) {
            const int ax = (e < n0) ? 0 : (e < n0 + n1 ? 1 : 2);
            const int i = e - (ax == 0 ? 0 : (ax == 1 ? n0 : n0 + n1));
            const float lo = (ax == 0) ? b.lo.x : (ax == 1 ? b.lo.y : b.lo.z);
            const float q = (float)(lo + i * res);
            const double t = E.inv_res * (double)q;
            const int c = (int)floor(t) - E.min_key[ax];
            const double vr = rint(t);
            int v = (int)vr - E.min_key[ax];
            int s = v - c;
            const bool oob = c < 0 || c >= E.dims[ax];
            if (fabs(t - vr) > 1e-2 || (!oob && (s < 0 || s > 1))) bad[ax] = true;
            if (oob) { v = 0; s = 0; }
            const float olo = 0.f - half_res, ohi = 0.f + half_res;          // obstacle (0,0,0) +- res/2
            const float cq = (q < olo) ? olo : (q > ohi ? ohi : q);
            const bool near = (double)fabsf(cq - q) < thr;
            if (ax == 2 && near) { nz_all = true; if (oob) nz_oob = true; }
            tab->mc[ax][i] = (v << 3) | (near ? 4 : 0) | (oob ? 2 : 0) | (s & 1);
        }
        const unsigned fl = (bad[0] ? 1u : 0u) | (bad[1] ? 2u : 0u) | (bad[2] ? 4u : 0u) | (nz_all ? 8u : 0u) | (nz_oob ? 16u : 0u);
        const unsigned all = g.reduce_or(fl);           // OR over the lanes (also the barrier that completes the tables)
        if (miss[2]) {
            const int zs = E.zs;
            for (int e = g.lane; e < (zs >> 2); e += g.width) tab->zm[e] = 0u;
            g.sync();
            uint8_t* zb = reinterpret_cast<uint8_t*>(tab->zm);
            for (int i = g.lane; i < mm[2]; i += g.width) {
                const int ez = tab->mc[2][i];
                if (!(ez & 2)) zb[ez >> 3] = (ez & 1) ? 0xF0 : 0x0F;
            }
        }
        if (g.lane == 0) {
#pragma unroll
            for (int ax = 0; ax < 3; ax++)
                if (miss[ax]) {
                    tab->tlo[ax] = f32_bits(v3_get(b.lo, ax)); tab->thi[ax] = f32_bits(v3_get(b.hi, ax)); tab->tm[ax] = mm[ax];
                    tab->kflag[ax] = (int)((all >> ax) & 1u) | (ax == 2 ? (int)((all >> 3) & 3u) << 1 : 0);
                }
        }
        g.sync();
    }
    if ((tab->kflag[0] | tab->kflag[1] | tab->kflag[2]) & 1) return -1;
    const bool nz_all = (tab->kflag[2] & 2) != 0, nz_oob = (tab->kflag[2] & 4) != 0;
    const int m0 = mm[0], m1 = mm[1], m2 = mm[2];
    int vz0 = tab->mc[2][0] >> 3, vz1 = tab->mc[2][m2 - 1] >> 3;
    if (tab->mc[2][0] & 2) vz0 = 0;                      // oob below: the in-grid entries start at 0 or later
    if (tab->mc[2][m2 - 1] & 2) vz1 = E.dims[2];         // oob above
    const int ch0 = vz0 >> 4, nch = (vz1 >> 4) - ch0 + 1;
    const int per_x = m1 * nch, items = m0 * per_x;
    const uint32_t mg_x = fastdiv_magic((uint32_t)per_x), mg_c = fastdiv_magic((uint32_t)nch);
    const size_t sy = (size_t)E.zs, sx = (size_t)(E.dims[1] + 1) * E.zs;
    const uint32_t* zm = tab->zm;
    // Most tests find nothing, so there is no early exit inside a batch: the loads of a batch are independent and
    // stay in flight together (the serial chain of ~70 box tests per expansion is latency-bound); one vote per
    // kVoteItems work items.
    constexpr int U = 8, kVoteItems = 2048;
    for (int vbase = 0; vbase < items; vbase += kVoteItems) {
        const int vend = (vbase + kVoteItems < items) ? vbase + kVoteItems : items;
        uint32_t hit = 0;
        for (int base = vbase; base < vend; base += g.width * U) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int idx = base + u * g.width + g.lane;
                if (idx < vend) {
                    const int ix = (int)fastdiv((uint32_t)idx, mg_x), r = idx - ix * per_x;
                    const int iy = (int)fastdiv((uint32_t)r, mg_c), ch = ch0 + (r - iy * nch);
                    const int ex = tab->mc[0][ix], ey = tab->mc[1][iy];
                    const bool near_xy = (ex & ey & 4) != 0;
                    if ((ex | ey) & 2) {
                        if (near_xy && nz_all) hit |= 1u;       // whole column outside the grid
                    } else {
                        const uint8_t* p = E.vmask + (size_t)(ex >> 3) * sx + (size_t)(ey >> 3) * sy + (size_t)ch * 16;
#ifdef __CUDA_ARCH__
                        const uint4 B = __ldg(reinterpret_cast<const uint4*>(p));
#else
                        U4 B; memcpy(&B, p, 16);
#endif
                        const uint32_t* z4 = zm + ch * 4;
                        const uint32_t sel = 0x11111111u << ((ex & 1) | ((ey & 1) << 1));
                        hit |= ((B.x & z4[0]) | (B.y & z4[1]) | (B.z & z4[2]) | (B.w & z4[3])) & sel;
                        if (near_xy && nz_oob) hit |= 1u;       // the column's out-of-grid z entries
                    }
                }
            }
        }
        if (g.any(hit != 0)) return 1;
    }
    return 0;
}

DLSC_HD bool obstacle_in_box(const Group& g, const DevParams& P, const EdtDev& E, const Box& b, double margin,
                             SfcTab* tab, long long* lookups) {
    // lookups: [0] lattice vertices of the tested boxes, [1] box tests through the vertex mask, [2] through the records,
    // [3] answered by the summed-area query, [4] vertices of the non-redundant tests ([5] = flag set by the caller)
    if (E.vmask && margin == E.mask_margin && E.zs <= kSfcZsMax) {
        const int r = obstacle_in_box_mask(g, P, E, b, margin, tab, lookups);
        if (r >= 0) {
            if (lookups && g.lane == 0) lookups[1] += 1;
            return r != 0;
        }
    }
    const double res = P.world_res;
    const float half_res = (float)(0.5 * res);
    const int m0 = (int)floor(((b.hi.x - b.lo.x) + kEpsF) / res) + 1;
    const int m1 = (int)floor(((b.hi.y - b.lo.y) + kEpsF) / res) + 1;
    const int m2 = (int)floor(((b.hi.z - b.lo.z) + kEpsF) / res) + 1;
    if (m0 <= 0 || m1 <= 0 || m2 <= 0) return false;
    const int total = m0 * m1 * m2;
    if (lookups && g.lane == 0) lookups[2] += 1;
    if (m0 > kSfcTabMax || m1 > kSfcTabMax || m2 > kSfcTabMax) {
        // oversized box: direct evaluation
        for (int base = 0; base < total; base += g.width) {
            const int idx = base + g.lane;
            bool hit = false;
            if (idx < total) {
                const int iz = idx % m2, t = idx / m2;
                const int iy = t % m1, ix = t / m1;
                const V3 q = v3((float)(b.lo.x + ix * res), (float)(b.lo.y + iy * res), (float)(b.lo.z + iz * res));
                hit = vertex_blocked(E, q, margin, half_res);
                if (lookups) { lookups[0] += 1; if (lookups[5]) lookups[4] += 1; }
            }
            if (g.any(hit)) return true;
        }
        return false;
    }
    // per-axis tables: coordinate (float)(lo + i*res) and cell floor(inv_res * coord) - min_key (or -1)
    for (int e = g.lane; e < m0 + m1 + m2; e += g.width) {
        const int ax = (e < m0) ? 0 : (e < m0 + m1 ? 1 : 2);
        const int i = e - (ax == 0 ? 0 : (ax == 1 ? m0 : m0 + m1));
        const float lo = (ax == 0) ? b.lo.x : (ax == 1 ? b.lo.y : b.lo.z);
        const float q = (float)(lo + i * res);
        int c = (int)floor(E.inv_res * (double)q) - E.min_key[ax];
        if (c < 0 || c >= E.dims[ax]) c = -1;
        tab->q[ax][i] = q;
        tab->c[ax][i] = c;
    }
    g.sync();
    const int m12 = m1 * m2;
    for (int base = 0; base < total; base += g.width * kSfcUnroll) {
        bool hit = false;
#pragma unroll
        for (int u = 0; u < kSfcUnroll; u++) {
            const int idx = base + u * g.width + g.lane;
            if (idx < total) {
                const int ix = idx / m12, r = idx - ix * m12;
                const int iy = r / m2, iz = r - iy * m2;
                hit = vertex_blocked_tab(E, tab->q[0][ix], tab->q[1][iy], tab->q[2][iz], tab->c[0][ix], tab->c[1][iy],
                                         tab->c[2][iz], margin, half_res) || hit;
                if (lookups) { lookups[0] += 1; if (lookups[5]) lookups[4] += 1; }
            }
        }
        if (g.any(hit)) return true;
    }
    g.sync();
    return false;
}

// isSFCInBoundary (:894-901), margin 0 at the only call site
DLSC_HD bool box_in_boundary(const DevParams& P, const Box& b) {
    for (int k = 0; k < 3; k++) {
        if (!(v3_get(b.lo, k) > P.world_min[k] + 0.0 - kEpsF)) return false;
        if (!(v3_get(b.hi, k) < P.world_max[k] - 0.0 + kEpsF)) return false;
    }
    return true;
}

// expandSFCIncrementally (:1023-1093)
// Box test of the greedy growth with the lattice description of the box kept in registers.  Consecutive tests of
// expand_incrementally differ in one face, so only the moved axis is re-derived from the floats (lattice_axis, the
// same formulas obstacle_in_box_mask applies to all three axes); when the summed-area query over that vertex range
// is empty the box is free -- exactly the first exit of obstacle_in_box_mask, with the same work counters.  Anything
// else (flagged vertices, off-lattice corner, oversized or out-of-grid range, no mask) goes through obstacle_in_box.
struct LatticeBox { int vv[3], mm[3]; };
DLSC_HD bool box_test(const Group& g, const DevParams& P, const EdtDev& E, const Box& b, const LatticeBox& L, bool fast_ok,
                      double margin, SfcTab* memo, long long* lookups) {
    if (fast_ok && L.mm[0] > 0 && L.mm[1] > 0 && L.mm[2] > 0 && L.mm[0] <= kSfcTabMax && L.mm[1] <= kSfcTabMax &&
        L.mm[2] <= kSfcTabMax && L.vv[0] >= 0 && L.vv[1] >= 0 && L.vv[2] >= 0 && L.vv[0] + L.mm[0] - 1 <= E.dims[0] &&
        L.vv[1] + L.mm[1] - 1 <= E.dims[1] && L.vv[2] + L.mm[2] - 1 <= E.dims[2]) {
        if (sat_count(g, E, L.vv[0], L.vv[0] + L.mm[0] - 1, L.vv[1], L.vv[1] + L.mm[1] - 1, L.vv[2],
                      L.vv[2] + L.mm[2] - 1) == 0) {
            if (lookups && g.lane == 0) {
                const long long nvert = (long long)L.mm[0] * L.mm[1] * L.mm[2];
                lookups[0] += nvert;
                if (lookups[5]) lookups[4] += nvert;
                lookups[3] += 1;
                lookups[1] += 1;
            }
            return false;
        }
    }
    return obstacle_in_box(g, P, E, b, margin, memo, lookups);
}

template <int K> DLSC_HD float& v3_comp(V3& v) { if (K == 0) return v.x; if (K == 1) return v.y; return v.z; }
// grow face AX (0..2: low faces, 3..5: high faces) of `cand` by one cell and make `upd` the new slab (:1058-1063);
// returns isSFCInBoundary of the slab (it shares four faces with the verified `cand`: only the growth axis is new)
template <int AX>
DLSC_HD bool sfc_grow_face(const DevParams& P, const EdtDev& E, double res, Box& cand, Box& upd, LatticeBox& Lc, LatticeBox& Lu) {
    constexpr int A3 = (AX < 3) ? AX : AX - 3;
    if (AX < 3) {
        v3_comp<A3>(upd.hi) = v3_comp<A3>(cand.lo);
        v3_comp<A3>(cand.lo) = (float)(v3_comp<A3>(cand.lo) - res);
        v3_comp<A3>(upd.lo) = v3_comp<A3>(cand.lo);
    } else {
        v3_comp<A3>(upd.lo) = v3_comp<A3>(cand.hi);
        v3_comp<A3>(cand.hi) = (float)(v3_comp<A3>(cand.hi) + res);
        v3_comp<A3>(upd.hi) = v3_comp<A3>(cand.hi);
    }
    lattice_axis(E, res, A3, v3_comp<A3>(cand.lo), v3_comp<A3>(cand.hi), Lc.mm[A3], Lc.vv[A3]);
    lattice_axis(E, res, A3, v3_comp<A3>(upd.lo), v3_comp<A3>(upd.hi), Lu.mm[A3], Lu.vv[A3]);
    return (v3_comp<A3>(upd.lo) > P.world_min[A3] + 0.0 - kEpsF) && (v3_comp<A3>(upd.hi) < P.world_max[A3] - 0.0 + kEpsF);
}
DLSC_HD bool expand_incrementally(const Group& g, const DevParams& P, const EdtDev& E, const Box& init,
                                  double margin, double max_vel, Box& out, SfcTab* memo, long long* lookups) {
    const double res = P.world_res;
    const bool fast_ok = E.vmask && E.sat && margin == E.mask_margin && E.zs <= kSfcZsMax;
    if (lookups) lookups[5] = 1;                  // the initial box and the slabs are the algorithmic tests
    if (obstacle_in_box(g, P, E, init, margin, memo, lookups)) return false;
    int axes = 0x543210;              // packed axis list, 4 bits each: -x -y -z +x +y +z
    int n_axes = 6;
    unsigned long long iters = 0;     // growth steps per axis, 8 bits each
    const double span = (2 * P.grid_res < max_vel * P.dt) ? max_vel * P.dt : 2 * P.grid_res;
    const int max_iter = (int)round(span / res) + 1;
    int i = -1;
    Box sfc = init, cand = init, upd = init;
    // lattice descriptions: Ls of `sfc`, Lc of `cand`, Lu of `upd` (Lu = Lc with the growth axis replaced by the slab)
    LatticeBox Ls, Lc, Lu;
#pragma unroll
    for (int k = 0; k < 3; k++) lattice_axis(E, res, k, v3_get(init.lo, k), v3_get(init.hi, k), Ls.mm[k], Ls.vv[k]);
    Lc = Ls; Lu = Ls;
    while (n_axes > 0) {
        cand = sfc; upd = sfc;
        Lc = Ls; Lu = Ls;
        if (lookups) lookups[5] = 0;              // the whole-box recheck at the top of a pass is the reference's redundancy
        // isSFCInBoundary on `upd`: the whole box at the top of a pass; afterwards `upd` is a slab that shares four
        // faces with the already verified `cand`, so only its two faces along the growth axis are new
        bool inside = box_in_boundary(P, upd);
        while (inside && !box_test(g, P, E, upd, Lu, fast_ok, margin, memo, lookups)) {
            if (lookups) lookups[5] = 1;
            i++;
            if (i >= n_axes) i = 0;
            const int ax = (axes >> (4 * i)) & 0xf;
            sfc = cand; upd = cand;
            Ls = Lc; Lu = Lc;
            // one growth step of face `ax`, specialised per face: the control flow is uniform in the warp, so the switch
            // costs one branch and every component / lattice slot below is addressed at compile time
            switch (ax) {
                case 0: inside = sfc_grow_face<0>(P, E, res, cand, upd, Lc, Lu); break;
                case 1: inside = sfc_grow_face<1>(P, E, res, cand, upd, Lc, Lu); break;
                case 2: inside = sfc_grow_face<2>(P, E, res, cand, upd, Lc, Lu); break;
                case 3: inside = sfc_grow_face<3>(P, E, res, cand, upd, Lc, Lu); break;
                case 4: inside = sfc_grow_face<4>(P, E, res, cand, upd, Lc, Lu); break;
                default: inside = sfc_grow_face<5>(P, E, res, cand, upd, Lc, Lu); break;
            }
            iters += 1ull << (8 * ax);
            if ((int)((iters >> (8 * ax)) & 0xffull) > max_iter) break;
        }
        if (i < 0) return false;      // start box outside the world (the reference would erase begin()-1)
        // erase axes[i]
        const int low = axes & ((1 << (4 * i)) - 1);
        const int high = (axes >> (4 * (i + 1))) << (4 * i);
        axes = low | high;
        n_axes--;
        if (i > 0) i--; else i = n_axes - 1;
    }
    const double delta = margin - ((int)(margin / res) * res);     // :1081
    for (int k = 0; k < 3; k++) {
        if (v3_get(sfc.lo, k) > P.world_min[k] + kEpsF) v3_set(sfc.lo, k, (float)(v3_get(sfc.lo, k) - delta));
        if (v3_get(sfc.hi, k) < P.world_max[k] - kEpsF) v3_set(sfc.hi, k, (float)(v3_get(sfc.hi, k) + delta));
    }
    out = sfc;
    return true;
}

DLSC_HD Box hull_aabb(const V3* pts, int np) {
    Box b; b.lo = pts[0]; b.hi = pts[0];
    for (int t = 0; t < np; t++)
        for (int k = 0; k < 3; k++) {
            const float c = v3_get(pts[t], k);
            if (c < v3_get(b.lo, k)) v3_set(b.lo, k, c);
            if (c > v3_get(b.hi, k)) v3_set(b.hi, k, c);
        }
    return b;
}

// One agent's SFC update (init: :435-452 ; else :502-536 with the two expandSFCFromConvexHull overloads
// :781-860).  sfc [M][6] in/out.  Returns status bits.
DLSC_HD int sfc_agent(const Group& g, const DevParams& P, const EdtDev& E, bool init, const V3& pos,
                      const float* init_traj, const V3& goal, const V3& wp, double radius, double max_vel,
                      float* sfc, SfcTab* memo, long long* lookups) {
    const int M = P.M;
    const double res = P.world_res;
    int status = 0;
    sfc_tab_reset(g, memo);
    if (init) {
        Box b;
        for (int k = 0; k < 3; k++) {
            v3_set(b.lo, k, (float)(floor(v3_get(pos, k) / res) * res));
            v3_set(b.hi, k, (float)(ceil(v3_get(pos, k) / res) * res));
        }
        Box out;
        if (!expand_incrementally(g, P, E, b, radius, max_vel, out, memo, lookups)) {
            status = kStSfcInitFailed;
            out = b;
        }
        for (int m = g.lane; m < M; m += g.width) box_store(sfc + m * 6, out);
        return status;
    }
    // shift + refinement (:506-516) -- serial dependency along m, done redundantly by every lane
    Box cur[kMaxM];
    for (int m = 0; m < M - 1; m++) cur[m] = box_load(sfc + (m + 1) * 6);
    const Box prev = box_load(sfc + (M - 1) * 6);
    for (int m = 0; m < M - 2; m++) {
        V3 cps[kP];
        for (int i = 0; i < kP; i++) cps[i] = v3_load(init_traj + (m * kP + i) * 3);
        if (superset_of_hull(cur[m + 1], cps, kP)) cur[m] = cur[m + 1];
    }
    V3 hull[3];
    hull[0] = v3_load(init_traj + ((M - 1) * kP + (kP - 1)) * 3); hull[1] = goal; hull[2] = wp;
    Box upd;
    // expandSFCFromConvexHull(convex_hull) :781-815
    Box b = hull_aabb(hull, 3);
    for (int k = 0; k < 3; k++) {
        v3_set(b.lo, k, (float)(round(v3_get(b.lo, k) / res) * res));
        v3_set(b.hi, k, (float)(round(v3_get(b.hi, k) / res) * res));
    }
    bool ok = expand_incrementally(g, P, E, b, radius, max_vel, upd, memo, lookups);
    if (ok && !superset_of_hull(upd, hull, 3)) ok = false;
    if (!ok) {
        // expandSFCFromConvexHull(convex_hull, sfc_prev) :817-860
        b = hull_aabb(hull, 2);
        for (int k = 0; k < 3; k++) {
            v3_set(b.lo, k, (float)(floor(v3_get(b.lo, k) / res) * res));
            v3_set(b.hi, k, (float)(ceil(v3_get(b.hi, k) / res) * res));
        }
        if (!box_includes(prev, b)) {
            Box r;
            for (int k = 0; k < 3; k++) {
                const float l0 = v3_get(prev.lo, k), l1 = v3_get(b.lo, k);
                const float h0 = v3_get(prev.hi, k), h1 = v3_get(b.hi, k);
                v3_set(r.lo, k, (l0 < l1) ? l1 : l0);
                v3_set(r.hi, k, (h1 < h0) ? h1 : h0);
            }
            for (int k = 0; k < 3; k++) {
                v3_set(b.lo, k, (float)(ceil((v3_get(r.lo, k) - kEpsF) / res) * res));
                v3_set(b.hi, k, (float)(floor((v3_get(r.hi, k) + kEpsF) / res) * res));
            }
        }
        ok = expand_incrementally(g, P, E, b, radius, max_vel, upd, memo, lookups);
        if (!ok) { upd = prev; status |= kStSfcReused; }
    }
    cur[M - 1] = upd;
    g.sync();
    for (int m = g.lane; m < M; m += g.width) box_store(sfc + m * 6, cur[m]);
    return status;
}

// ------------------------------------------------------------------------------------------------
// goal line search: closed form of the 1-variable LP (goal_optimizer.cpp:7-136, 138-198)
//   min t in [0, 1+1e-5]  s.t.  a_r t + b_r >= 0 ; row activity tolerance 1e-6 (CPLEX default)
// LSC arrays of this agent: normal [K][M][3], d [K][M][P], anchor_last [K][3]
// ------------------------------------------------------------------------------------------------
struct GoalRows {
    const DevParams* P; const float* sfc_last; int K;
    const float* normal; const double* d; const float* anchor_last;
};
DLSC_HD int goal_num_rows(const GoalRows& R) { return (R.P->use_sfc ? 2 * R.P->D : 0) + R.K; }
// row r -> (a, b); returns false for skipped rows (zero normal, goal_optimizer.cpp:182-184)
DLSC_HD bool goal_row(const GoalRows& R, int r, const V3& gw, const V3& wp, double& a, double& b) {
    const DevParams& P = *R.P;
    float nrm[3] = {0.f, 0.f, 0.f}, anc[3] = {0.f, 0.f, 0.f};
    double dd;
    const int nb = P.use_sfc ? 2 * P.D : 0;
    if (r < nb) {                                               // Box::convertToLSCs :66-87
        const int i = r >> 1;
        if ((r & 1) == 0) { nrm[i] = 1.f; dd = (double)R.sfc_last[i]; }
        else { nrm[i] = -1.f; dd = -(double)R.sfc_last[3 + i]; }
    } else {
        const int oi = r - nb;
        const float* n = R.normal + ((size_t)oi * P.M + (P.M - 1)) * 3;
        nrm[0] = n[0]; nrm[1] = n[1]; nrm[2] = n[2];
        if (v3_norm(v3(nrm[0], nrm[1], nrm[2])) < kEpsF) return false;
        anc[0] = R.anchor_last[oi * 3]; anc[1] = R.anchor_last[oi * 3 + 1]; anc[2] = R.anchor_last[oi * 3 + 2];
        dd = R.d[((size_t)oi * P.M + (P.M - 1)) * kP + (kP - 1)];
    }
    a = 0; b = 0;
    for (int k = 0; k < P.D; k++) {
        a += (double)nrm[k] * (double)v3_get(gw, k);
        b += (double)nrm[k] * ((double)v3_get(wp, k) - (double)anc[k]);
    }
    b = b - dd;
    return true;
}

DLSC_HD int goal_agent(const Group& g, const DevParams& P, bool disturbed, const V3& pos, const V3& wp,
                       const float* sfc_last, int K, const float* normal, const double* d, const float* anchor_last,
                       V3& goal) {
    if (disturbed) { goal = pos; return 0; }                          // traj_planner.cpp:447-450
    if (v3_distance(goal, wp) < kEpsF) { goal = wp; return 0; }       // goal_optimizer.cpp:12-14
    const V3 gw = goal - wp;                                          // float coefficients :165
    GoalRows R; R.P = &P; R.sfc_last = sfc_last; R.K = K; R.normal = normal; R.d = d; R.anchor_last = anchor_last;
    const int nr = goal_num_rows(R);
    // each lane keeps its rows' (a, b): one pass over memory for both the bound and the feasibility check
    double tlo = 0.0;
    for (int r = g.lane; r < nr; r += g.width) {
        double a, b;
        if (!goal_row(R, r, gw, wp, a, b)) continue;
        if (a > 0) { const double c = -b / a; tlo = (tlo < c) ? c : tlo; }
    }
    tlo = g.max(tlo);
    const double t = (1.0 + kEpsF < tlo) ? 1.0 + kEpsF : tlo;
    bool infeasible = false;
    const double tol = 1e-6;
    for (int r = g.lane; r < nr; r += g.width) {
        double a, b;
        if (!goal_row(R, r, gw, wp, a, b)) continue;
        if (a * t + b < -tol) infeasible = true;
    }
    if (!g.any(infeasible)) { goal = gw * (float)t + wp; return 0; }  // :51
    // infeasible: "numerical error" rule :55-81
    bool not_numerical = false;
    if (P.use_sfc) {
        const Box b = box_load(sfc_last);
        if (!point_in_box(b, goal)) not_numerical = true;
    }
    for (int oi = g.lane; oi < K; oi += g.width) {
        const V3 nv = v3_load(normal + ((size_t)oi * P.M + (P.M - 1)) * 3);
        if (v3_norm(nv) < kEpsF) continue;
        const V3 anc = v3_load(anchor_last + oi * 3);
        const double dd = d[((size_t)oi * P.M + (P.M - 1)) * kP + (kP - 1)];
        const double delta = v3_dot(nv, goal - anc) - dd;             // :73
        if (delta < -kEpsF) not_numerical = true;
    }
    return g.any(not_numerical) ? kStGoalInfeasible : 0;              // keep goal :80
}

// ------------------------------------------------------------------------------------------------
// dynamic (non-agent) obstacles: traj_planner.cpp:617-627 + :1129-1148 + :1080-1100 (LSC), :708-735 (waypoint
// trap), geometry.hpp:115-137, obstacle.hpp:26-36, collision_constraints.cpp:538-546, 586-598.
// They take the first P.n_dyn obstacle slots of every agent (MultiSyncSimulator::broadcastMsgs lists them first,
// multi_sync_simulator.cpp:476-480); their constant-velocity predictions are rows N.. of pred_traj.
// ------------------------------------------------------------------------------------------------
struct DynObs { const float *pos, *vel; const double *radius, *downwash, *max_acc, *size; };

// Obstacle sizes over the horizon (obstacleSizePredictionWithConstAcc, traj_planner.cpp:338-368): radius + the Bernstein
// control points of 1/2 a_max t^2 up to the uncertainty horizon, constant after it.  coef * B_inv (polynomial.hpp:280-293)
// with B_inv(k, i) = C(i,k)/C(n,k) in closed form.  Host side: runs once per dlsc_set_obstacles.
inline void dyn_obstacle_sizes(const DevParams& P, bool size_prediction, double horizon, double radius, double max_acc,
                               double* out /*[M][P]*/) {
    const int M = P.M, n = kP - 1;
    int Mu = (int)((horizon + kEps) / P.dt);                                   // :339
    if (Mu > M) Mu = M;
    for (int m = 0; m < M; m++)
        for (int i = 0; i < kP; i++) {
            double v = radius;                                                 // :364
            if (size_prediction) {
                if (m < Mu) {
                    const double a0 = 0.5 * max_acc * ((m * P.dt) * (m * P.dt));   // :349-351 (pow(x, 2) = x * x exactly)
                    const double a1 = max_acc * m * P.dt * P.dt;
                    const double a2 = 0.5 * max_acc * (P.dt * P.dt);
                    v = radius + (a0 + a1 * ((double)i / n) + a2 * ((double)(i * (i - 1)) / (n * (n - 1))));
                } else {
                    v = radius + 0.5 * max_acc * ((Mu * P.dt) * (Mu * P.dt));   // :360-361
                }
            }
            out[(size_t)m * kP + i] = v;
        }
}

// normalVectorBetweenLines(line_obs, line_agent) on closestPointsBetweenLinePaths: both points move along their
// segments with the same parameter
DLSC_HD V3 normal_between_line_paths(const V3& o0, const V3& o1, const V3& a0, const V3& a1) {
    const V3 r0 = a0 - o0, r1 = a1 - o1;                                       // rel_path = line2 - line1
    const Closest rc = closest_point_segment(v3(0.f, 0.f, 0.f), r0, r1);
    const double len = v3_distance(r0, r1);
    double alpha = 0.0;
    if (len > 0) alpha = v3_norm(rc.p2 - r0) / len;
    const V3 c1 = o0 + (o1 - o0) * (float)alpha;
    const V3 c2 = a0 + (a1 - a0) * (float)alpha;
    V3 nv = v3_normalized(c2 - c1);
    if (v3_norm(nv) == 0) {                                                    // heuristic :1089-1098
        const V3 a = a0 - o0, b = a1 - o1;
        if (v3_norm(a) == 0 && v3_norm(b) == 0) nv = v3(1.f, 0.f, 0.f);
        else nv = v3_cross(b - a, v3(0.f, 0.f, 1.f));
    }
    return nv;
}

// one (agent, obstacle, segment) item.  The lines are NOT downwash-transformed (normalVectorDynamicObs :1144-1147);
// only the normal's z is divided afterwards.  near_out = 0: the QP never screens these rows out.
DLSC_HD void lsc_dynamic_segment(const float* init_seg, const float* obs_seg, const double* size_seg, double r_a,
                                 double r_o, double dw_o, float* normal_out, double* d_out, float* near_out) {
    const double downwash = (r_a + dw_o * r_o) / (r_a + r_o);                  // :1156-1157
    const V3 nt = normal_between_line_paths(v3_load(obs_seg), v3_load(obs_seg + (kP - 1) * 3), v3_load(init_seg),
                                            v3_load(init_seg + (kP - 1) * 3));
    v3_store(normal_out, v3(nt.x, nt.y, (float)((double)nt.z / downwash)));    // :618-620
    for (int i = 0; i < kP; i++) d_out[i] = size_seg[i] + r_a;                 // :624
    if (near_out) *near_out = 0.f;
}

// Obstacle::isCollided
DLSC_HD bool obstacle_collides(const V3& opos, const V3& ovel, double oradius, double omax_acc, const V3& point,
                               double agent_radius, double horizon, double uncertainty_horizon) {
    const double step = (0.1 * horizon < 0.1) ? 0.1 * horizon : 0.1;
    for (double t = 0; t <= horizon; t += step) {
        const V3 q = opos + ovel * (float)t;
        const double tm = t < uncertainty_horizon ? t : uncertainty_horizon;
        if (v3_distance(q, point) < agent_radius + oradius + 0.5 * omax_acc * tm * tm) return true;
    }
    return false;
}

// CollisionConstraints::constructCommunicationRange, reached from generateSFC's non-initial branch only
DLSC_HD void comm_box_update(const DevParams& P, bool init, const V3& wp, float* box) {
    if (!P.use_sfc || init || !(P.comm_range > 0)) return;
    const float h = (float)(0.5 * P.comm_range);
    v3_store(box, wp - v3(h, h, h));
    v3_store(box + 3, wp + v3(h, h, h));
}

// checkWaypointTrap for one agent; the lanes of the group share the agent LSCs of the feasibility test and the
// obstacles of the clearing loop.  LSC arrays of this agent as in goal_agent; K counts the dynamic slots.
// Returns 1 when the waypoint is trapped; the LSCs of the dynamic obstacles that can reach the waypoint are cleared.
DLSC_HD int waypoint_trap(const Group& gr, const DevParams& P, const DynObs& O, const V3& goal, const V3& wp,
                          const float* sfc_last, const float* comm_box, int K, float* normal, double* d,
                          const float* anchor_last, double agent_radius) {
    const int M = P.M, nd = P.n_dyn;
    if (K == 0) return 0;                                                      // obstacles.empty() :709
    bool bad = false;                                                          // isPointInFeasibleRegion(goal) and (waypoint)
    for (int oi = nd + gr.lane; oi < K; oi += gr.width) {
        const V3 nv = v3_load(normal + ((size_t)oi * M + (M - 1)) * 3);
        const V3 an = v3_load(anchor_last + oi * 3);
        const double dd = d[((size_t)oi * M + (M - 1)) * kP + (kP - 1)];
        if (!(v3_dot(goal - an, nv) - dd > -kEps) || !(v3_dot(wp - an, nv) - dd > -kEps)) bad = true;   // LSC::isPointInLSC
    }
    if (P.use_sfc && !(point_in_box(box_load(sfc_last), goal) && point_in_box(box_load(sfc_last), wp))) bad = true;
    if (!(point_in_box(box_load(comm_box), goal) && point_in_box(box_load(comm_box), wp))) bad = true;
    if (!gr.any(bad)) return 0;
    for (int oi = gr.lane; oi < nd; oi += gr.width) {
        if (!obstacle_collides(v3_load(O.pos + 3 * oi), v3_load(O.vel + 3 * oi), O.radius[oi], O.max_acc[oi], wp,
                               agent_radius, M * P.dt, P.dyn_horizon)) continue;
        for (int m = 0; m < M; m++) {
            float* nr = normal + ((size_t)oi * M + m) * 3;
            nr[0] = nr[1] = nr[2] = 0.f;                                       // default LSC: skipped by the QP :731
            for (int i = 0; i < kP; i++) d[((size_t)oi * M + m) * kP + i] = 0.0;
        }
    }
    return 1;
}

// ------------------------------------------------------------------------------------------------
// state step: AgentManager::doStep -> Trajectory::getStateAt(t) (trajectory.cpp:111-170, 183-199)
// state[9] = pos, vel, acc.  Powers are formed by repeated multiplication: exact for the
// normalised time 0 or 1, i.e. for t = k*dt, which is the only way the replan loop calls it.
// ------------------------------------------------------------------------------------------------
DLSC_HD double ipow(double x, int e) { double r = 1.0; for (int i = 0; i < e; i++) r *= x; return r; }
DLSC_HD int binom(int n, int k) {                     // polynomial.hpp:9-20
    if (k > n) return 0;
    if (k * 2 > n) k = n - k;
    if (k == 0) return 1;
    int r = n;
    for (int i = 2; i <= k; i++) { r *= (n - i + 1); r /= i; }
    return r;
}
DLSC_HD void state_at(const DevParams& P, const float* traj, double time, float* state) {
    const int M = P.M;
    int mm = -1;
    double tn = 0, seg_end = 0;
    for (int idx = 0; idx < M; idx++) {
        seg_end += P.dt;
        if (time < seg_end) { mm = idx; tn = 1 - (seg_end - time) / P.dt; break; }
    }
    if (mm == -1 && time < seg_end + kEpsF) { mm = M - 1; tn = 1.0; }
    V3 cur[kP];
    for (int i = 0; i < kP; i++) cur[i] = (mm >= 0) ? v3_load(traj + (mm * kP + i) * 3) : v3(0.f, 0.f, 0.f);
    int n = kP - 1;
    for (int order = 0; order < 3; order++) {
        V3 point = v3(0.f, 0.f, 0.f);
        if (mm >= 0) {
            for (int i = 0; i < n + 1; i++) {
                const double b = binom(n, i) * ipow(tn, i) * ipow(1 - tn, n - i);
                point = point + cur[i] * (float)b;
            }
        }
        v3_store(state + 3 * order, point);
        for (int i = 0; i < n; i++) cur[i] = (cur[i + 1] - cur[i]) * (float)(n / P.dt);
        cur[n] = v3(0.f, 0.f, 0.f);
        n -= 1;
    }
}

}  // namespace dlsc
