// dlsc_kernels_exact.cu -- sm_100a kernels of the bit-exact stages (compile with -fmad=false):
//   k_predict      horizon shift / constant-velocity initial guess        thread per control point
//   k_neighbours   comm-range neighbour list, built on device             warp per agent (ballot compaction)
//   k_lsc          LSC separating planes: GJK hull-vs-origin + margins    thread per (agent, neighbour, segment)
//   k_sfc          SFC greedy box expansion over the lattice-vertex mask  warp per agent (vote any)
//   k_goal         intermediate goal line search (closed-form 1-var LP)   thread per agent
//   k_advance      state step at t = dt + record refresh                  thread per agent
// All arithmetic lives in dlsc_stages.cuh / dlsc_math.cuh.
#include <cstdlib>

#include "dlsc_kernels.h"
#include "dlsc_stages.cuh"

namespace dlsc {

// DevParams / DevState travel as __grid_constant__ kernel parameters (constant bank, ~0.8 KB).

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_predict(const __grid_constant__ DevParams P, const __grid_constant__ DevState S, int seq) {
    const int npt = P.M * kP;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P.N * npt) return;
    const int a = (int)(gid / npt), pt = (int)(gid - (long long)a * npt);
    const float* rec = S.rec + (size_t)a * P.rec;
    const int la = a - P.begin;
    const bool local = la >= 0 && la < P.NL;
    predict_point(P, rec, seq, pt, local ? (S.disturbed[la] != 0) : false, S.pred_traj + (size_t)a * npt * 3,
                  local ? S.init_traj + (size_t)la * npt * 3 : nullptr);
    if (local && pt == 0) S.status[la] = 0;
}

void launch_predict(const DevParams& P, const DevState& S, int seq, cudaStream_t st) {
    const long long n = (long long)P.N * P.M * kP;
    k_predict<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, S, seq);
}

// dynamic obstacles (P.n_dyn > 0 only): constant-velocity predictions (traj_planner.cpp:303-305) into rows N.. of
// pred_traj; per local agent the communication box of this step and the heads of the obstacle list
__global__ void __launch_bounds__(128) k_dyn_predict(const __grid_constant__ DevParams P, const __grid_constant__ DevState S) {
    const int npt = P.M * kP, nd = P.n_dyn;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nd * npt) {
        const int o = t / npt, pt = t - o * npt;
        v3_store(S.pred_traj + ((size_t)(P.N + o) * npt + pt) * 3, v3_load(S.dyn_pos + 3 * o) + v3_load(S.dyn_vel + 3 * o) * P.tk[pt]);
    }
    if (t < P.NL) {
        const bool init = S.sfc_init[t] != 0 || S.disturbed[t] != 0;
        comm_box_update(P, init, v3_load(S.waypoint + t * 3), S.comm_box + (size_t)t * 6);
        for (int o = 0; o < nd; o++) S.nbr_idx[(size_t)t * P.K + o] = P.N + o;
    }
}
void launch_dyn_predict(const DevParams& P, const DevState& S, cudaStream_t st) {
    const int n = P.n_dyn * P.M * kP > P.NL ? P.n_dyn * P.M * kP : P.NL;
    k_dyn_predict<<<(n + 127) / 128, 128, 0, st>>>(P, S);
}

// ------------------------------------------------------------------------------------------------
// warp per agent; the 16 warps of a CTA share position tiles staged in shared memory (the positions sit inside
// the 768-byte records: each is fetched once per CTA instead of once per warp).  Ballot compaction keeps the
// reference's ascending neighbour order.
constexpr int kNbrWarps = 16, kNbrTile = kNbrWarps * 32;
__global__ void __launch_bounds__(kNbrTile) k_neighbours(const __grid_constant__ DevParams P, const __grid_constant__ DevState S) {
    __shared__ float tile[kNbrTile * 4];         // position + group (mission index) of the tile's agents
    __shared__ float grange[kNbrWarps * 2];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int la = blockIdx.x * kNbrWarps + w;
    const bool valid = la < P.NL;
    const int a = P.begin + la;
    const int off = P.M * kP * 3;
    const V3 pa = valid ? v3_load(S.rec + (size_t)a * P.rec + off) : v3(0.f, 0.f, 0.f);
    const float ga = valid ? S.rec[(size_t)a * P.rec + off + 11] : 0.f;
    const int Kc = P.K - P.n_dyn;                      // the dynamic obstacles hold the first slots
    int32_t* idx_out = S.nbr_idx + (size_t)la * P.K + P.n_dyn;
    int count = 0;
    for (int base = 0; base < P.N; base += kNbrTile) {
        const int jt = base + threadIdx.x;
        if (jt < P.N) {
            const float* r = S.rec + (size_t)jt * P.rec + off;
            tile[threadIdx.x * 4] = r[0]; tile[threadIdx.x * 4 + 1] = r[1]; tile[threadIdx.x * 4 + 2] = r[2];
            tile[threadIdx.x * 4 + 3] = r[11];
        }
        // mission-index range of the tile: a Monte-Carlo batch keeps the missions contiguous, so all but one or two
        // tiles hold no agent of this warp's mission and are skipped as a whole
        float gmin = (jt < P.N) ? tile[threadIdx.x * 4 + 3] : 3.0e38f, gmax = (jt < P.N) ? tile[threadIdx.x * 4 + 3] : -3.0e38f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
            gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        }
        if (lane == 0) { grange[w * 2] = gmin; grange[w * 2 + 1] = gmax; }
        __syncthreads();
        gmin = grange[0]; gmax = grange[1];
#pragma unroll
        for (int i = 1; i < kNbrWarps; i++) { gmin = fminf(gmin, grange[i * 2]); gmax = fmaxf(gmax, grange[i * 2 + 1]); }
        const int nt = (P.N - base < kNbrTile) ? P.N - base : kNbrTile;
        if (valid && ga >= gmin && ga <= gmax)
            for (int t = 0; t < nt; t += 32) {
                const int j = base + t + lane;
                bool in = false;
                if (t + lane < nt && j != a && tile[(t + lane) * 4 + 3] == ga) in = in_comm_range(P, pa, v3_load(tile + (t + lane) * 4));
                const unsigned mask = __ballot_sync(0xffffffffu, in);
                const int pos = count + __popc(mask & ((1u << lane) - 1u));
                if (in && pos < Kc) idx_out[pos] = j;
                count += __popc(mask);
            }
        __syncthreads();
    }
    if (valid && lane == 0) {
        S.nbr_cnt[la] = P.n_dyn + (count < Kc ? count : Kc);
        if (count > Kc) atomicOr(S.status + la, kStNbrOverflow);
        atomicAdd(S.counters + 0, (unsigned long long)(count < Kc ? count : Kc));
    }
}

// ---- neighbour search through a uniform grid (comm_range > 0) -----------------------------------------------
// The all-pairs kernel above tests N candidates per agent; with a limited communication range only the agents of
// the 3 x 3 surrounding cells of an xy grid with cell size >= range can pass the Chebyshev test.  Same result as
// the all-pairs kernel (same exact test, ascending index order, first K kept on overflow), ~50 candidates per
// agent instead of N.
//   k_nbr_bin    one CTA: counting sort of all N agents by cell (shared-memory histogram + scan)
//   k_nbr_search warp per local agent: candidates of the 3 cell rows, exact test, rank sort
struct NbrGrid { int gx, gy; double x0, y0, inv_h; int* cell_start; int* sorted; float4* pos; };   // pos: {x, y, z, group} in sorted order
constexpr int kNbrMaxCells = 8192, kNbrBinThreads = 1024, kNbrCand = 256, kNbrSearchWarps = 8;

__device__ __forceinline__ int nbr_cell(const NbrGrid& G, float x, float y) {
    int cx = (int)floor(((double)x - G.x0) * G.inv_h), cy = (int)floor(((double)y - G.y0) * G.inv_h);
    cx = cx < 0 ? 0 : (cx >= G.gx ? G.gx - 1 : cx);
    cy = cy < 0 ? 0 : (cy >= G.gy ? G.gy - 1 : cy);
    return cy * G.gx + cx;
}

__global__ void __launch_bounds__(kNbrBinThreads) k_nbr_bin(const __grid_constant__ DevParams P, const float* __restrict__ rec,
                                                             const __grid_constant__ NbrGrid G) {
    __shared__ int cnt[kNbrMaxCells];
    __shared__ int wsum[32];
    const int ncell = G.gx * G.gy, off = P.M * kP * 3;
    for (int c = threadIdx.x; c < ncell; c += kNbrBinThreads) cnt[c] = 0;
    __syncthreads();
    // the first kCache agents of a thread stay in registers between the two passes (all of them when N <= 4096)
    constexpr int kCache = 4;
    float4 pc[kCache];
    int cc[kCache];
#pragma unroll
    for (int u = 0; u < kCache; u++) {
        const int j = threadIdx.x + u * kNbrBinThreads;
        cc[u] = -1;
        if (j < P.N) {
            const float* r = rec + (size_t)j * P.rec + off;
            pc[u] = make_float4(r[0], r[1], r[2], r[11]);
            cc[u] = nbr_cell(G, pc[u].x, pc[u].y);
            atomicAdd(&cnt[cc[u]], 1);
        }
    }
    for (int j = threadIdx.x + kCache * kNbrBinThreads; j < P.N; j += kNbrBinThreads) {
        const float* r = rec + (size_t)j * P.rec + off;
        atomicAdd(&cnt[nbr_cell(G, r[0], r[1])], 1);
    }
    __syncthreads();
    // exclusive scan of cnt[0..ncell) in chunks of 1024 (one element per thread), carry in wsum[31] style
    int carry = 0;
    for (int base = 0; base < ncell; base += kNbrBinThreads) {
        const int c = base + threadIdx.x;
        const int v = c < ncell ? cnt[c] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = wsum[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += t; }
            wsum[threadIdx.x] = wi - w;                       // exclusive prefix of the warp sums
        }
        __syncthreads();
        const int excl = carry + wsum[threadIdx.x >> 5] + incl - v;
        if (c < ncell) { cnt[c] = excl; G.cell_start[c] = excl; }
        __syncthreads();
        if (threadIdx.x == kNbrBinThreads - 1) wsum[0] = excl + v;   // total so far
        __syncthreads();
        carry = wsum[0];
        __syncthreads();
    }
    if (threadIdx.x == 0) G.cell_start[ncell] = carry;
    // order inside a cell is arbitrary: the search sorts
#pragma unroll
    for (int u = 0; u < kCache; u++)
        if (cc[u] >= 0) {
            const int slot = atomicAdd(&cnt[cc[u]], 1);
            G.sorted[slot] = threadIdx.x + u * kNbrBinThreads;
            G.pos[slot] = pc[u];
        }
    for (int j = threadIdx.x + kCache * kNbrBinThreads; j < P.N; j += kNbrBinThreads) {
        const float* r = rec + (size_t)j * P.rec + off;
        const int slot = atomicAdd(&cnt[nbr_cell(G, r[0], r[1])], 1);
        G.sorted[slot] = j;
        G.pos[slot] = make_float4(r[0], r[1], r[2], r[11]);
    }
}

__global__ void __launch_bounds__(kNbrSearchWarps * 32) k_nbr_search(const __grid_constant__ DevParams P, const __grid_constant__ DevState S,
                                                                      const __grid_constant__ NbrGrid G) {
    __shared__ int cand[kNbrSearchWarps][kNbrCand];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int la = blockIdx.x * kNbrSearchWarps + w;
    if (la >= P.NL) return;
    const int a = P.begin + la, off = P.M * kP * 3;
    const float* ra = S.rec + (size_t)a * P.rec + off;
    const V3 pa = v3_load(ra);
    const float ga = ra[11];
    const int Kc = P.K - P.n_dyn;                      // the dynamic obstacles hold the first slots
    int32_t* idx_out = S.nbr_idx + (size_t)la * P.K + P.n_dyn;
    const int ca = nbr_cell(G, pa.x, pa.y), cx = ca % G.gx, cy = ca / G.gx;
    const int x_lo = cx > 0 ? cx - 1 : 0, x_hi = cx < G.gx - 1 ? cx + 1 : G.gx - 1;
    int n = 0;                                   // in-range candidates found (all lanes agree)
    for (int yy = (cy > 0 ? cy - 1 : 0); yy <= (cy < G.gy - 1 ? cy + 1 : G.gy - 1); yy++) {
        const int s0 = G.cell_start[yy * G.gx + x_lo], s1 = G.cell_start[yy * G.gx + x_hi + 1];   // 3 cells, contiguous
        for (int t = s0; t < s1; t += 32) {
            bool in = false;
            int j = -1;
            if (t + lane < s1) {
                j = G.sorted[t + lane];
                const float4 q = G.pos[t + lane];
                if (j != a && q.w == ga) in = in_comm_range(P, pa, v3(q.x, q.y, q.z));
            }
            const unsigned mask = __ballot_sync(0xffffffffu, in);
            const int pos = n + __popc(mask & ((1u << lane) - 1u));
            if (in && pos < kNbrCand) cand[w][pos] = j;
            n += __popc(mask);
        }
    }
    __syncwarp();
    int count;
    if (n <= kNbrCand) {
        // rank sort: ascending agent index, the first K kept (multi_sync_simulator.cpp:481-503 visits j in order)
        for (int e = lane; e < n; e += 32) {
            const int j = cand[w][e];
            int rank = 0;
            for (int f = 0; f < n; f++) rank += (cand[w][f] < j);
            if (rank < Kc) idx_out[rank] = j;
        }
        count = n;
    } else {
        // more candidates than the buffer holds (dense Monte-Carlo batches): plain scan of all agents for this one
        count = 0;
        for (int base = 0; base < P.N; base += 32) {
            const int j = base + lane;
            bool in = false;
            if (j < P.N && j != a) {
                const float* rj = S.rec + (size_t)j * P.rec + off;
                if (rj[11] == ga) in = in_comm_range(P, pa, v3_load(rj));
            }
            const unsigned mask = __ballot_sync(0xffffffffu, in);
            const int pos = count + __popc(mask & ((1u << lane) - 1u));
            if (in && pos < Kc) idx_out[pos] = j;
            count += __popc(mask);
        }
    }
    if (lane == 0) {
        S.nbr_cnt[la] = P.n_dyn + (count < Kc ? count : Kc);
        if (count > Kc) atomicOr(S.status + la, kStNbrOverflow);
        atomicAdd(S.counters + 0, (unsigned long long)(count < Kc ? count : Kc));
    }
}

// parts: 1 = build the grid (once per step, all agents), 2 = search for the local agents of (P, S), 3 = both.
// Returns the number of kernels launched.
static bool nbr_grid_setup(const DevParams& P, const DevState& S, NbrGrid& G) {
    static const bool no_grid = [] { const char* e = getenv("DLSC_NBR_GRID"); return e && e[0] == '0'; }();
    if (!(P.comm_range > 0) || !S.nbr_cell_start || no_grid) return false;
    // cell size a little above the range: the test is made on float differences, which can accept a pair whose
    // exact distance exceeds the range by one float ulp
    const double h = P.comm_range * (1.0 + 1e-5);
    G.x0 = P.world_min[0]; G.y0 = P.world_min[1]; G.inv_h = 1.0 / h;
    G.gx = (int)floor((P.world_max[0] - P.world_min[0]) * G.inv_h) + 1;
    G.gy = (int)floor((P.world_max[1] - P.world_min[1]) * G.inv_h) + 1;
    if (!(G.gx >= 1 && G.gy >= 1 && (long long)G.gx * G.gy <= kNbrMaxCells)) return false;
    G.cell_start = S.nbr_cell_start; G.sorted = S.nbr_sorted; G.pos = S.nbr_sorted_pos;
    return true;
}
int launch_neighbours(const DevParams& P, const DevState& S, cudaStream_t st, int parts) {
    NbrGrid G;
    if (nbr_grid_setup(P, S, G)) {
        int n = 0;
        if (parts & 1) { k_nbr_bin<<<1, kNbrBinThreads, 0, st>>>(P, S.rec, G); n++; }
        if (parts & 2) { k_nbr_search<<<(P.NL + kNbrSearchWarps - 1) / kNbrSearchWarps, kNbrSearchWarps * 32, 0, st>>>(P, S, G); n++; }
        return n;
    }
    if (!(parts & 2)) return 0;
    k_neighbours<<<(P.NL + kNbrWarps - 1) / kNbrWarps, kNbrTile, 0, st>>>(P, S);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// LSC in two phases.
//   k_lsc       CTA per agent.  The agent's neighbour list, its own initial trajectory and the neighbours' (radius,
//               downwash, goal) are staged in shared memory, then thread per (neighbour, segment < M-1) item: the
//               two-vertex GJK (gjk::hull_origin_short) finishes ~90 % of the hulls of a swarm in one to three support
//               evaluations; those are written out at once.  The rest -- hulls that need the triangle / tetrahedron
//               sub-algorithms -- and the last-segment items (segment-segment closest points, a different code path)
//               are queued: collected per CTA in shared memory, one global atomic per CTA.  Small code, 64 registers.
//   k_lsc_rest  persistent grid over the two queues, homogeneous warps: last-segment items, then full GJK.
// Queue entry: la << 14 | neighbour slot << 4 | segment.  Results go to fixed slots, so the (non-deterministic) queue
// order cannot change them.
#ifndef DLSC_LSC_THREADS
#define DLSC_LSC_THREADS 128
#endif
#ifndef DLSC_LSC_MINB
#define DLSC_LSC_MINB 5
#endif
constexpr int kLscThreads = DLSC_LSC_THREADS, kLscRestThreads = 64;
constexpr int kCntSeg = 14, kCntHard = 15;     // S.counters slots used as queue lengths (zeroed with the counters)
__device__ __forceinline__ uint32_t lsc_pack(int la, int c, int m) { return ((uint32_t)la << 14) | ((uint32_t)c << 4) | (uint32_t)m; }

// DYN: dynamic obstacles present (the first P.n_dyn slots belong to k_lsc_dyn); the DYN = false instantiation is the
// swarm-only hot path, compiled without the slot offset
template <bool DYN>
__global__ void __launch_bounds__(kLscThreads, DLSC_LSC_MINB) k_lsc(const __grid_constant__ DevParams P, const __grid_constant__ DevState S) {
    __shared__ float s_init[kMaxPts * 3];
    __shared__ int s_nhard, s_base, s_seg;
    extern __shared__ __align__(8) unsigned char s_dyn[];   // [K] pair constants | [K] neighbour indices | [K (M-1)] hard items
    LscPair* s_pair = reinterpret_cast<LscPair*>(s_dyn);
    int* s_nbr = reinterpret_cast<int*>(s_pair + P.K);
    uint32_t* s_hard = reinterpret_cast<uint32_t*>(s_nbr + P.K);
    const int M = P.M, npt = M * kP;
    const int la = blockIdx.x;
    const int cnt = S.nbr_cnt[la], nd = DYN ? P.n_dyn : 0;  // slots [0, nd): dynamic obstacles, done by k_lsc_dyn
    const int og = npt * 3 + 6;
    if (threadIdx.x == 0) s_nhard = 0;
    for (int e = threadIdx.x; e < npt * 3; e += kLscThreads) s_init[e] = S.init_traj[(size_t)la * npt * 3 + e];
    const double r_a = S.radius[la], dw_a = S.downwash[la];
    for (int c = nd + threadIdx.x; c < cnt; c += kLscThreads) {      // per-neighbour constants once, not once per item
        const int j = S.nbr_idx[(size_t)la * P.K + c];
        const float* rec_j = S.rec + (size_t)j * P.rec;
        s_nbr[c] = j;
        s_pair[c] = lsc_pair_consts(r_a, dw_a, rec_j[og + 3], rec_j[og + 4]);
    }
    __syncthreads();
    int it_sum = 0;
    const int n_gjk = (cnt - nd) * (M - 1);
    const uint32_t mg = fastdiv_magic((uint32_t)(M - 1));
    for (int e = threadIdx.x; e < n_gjk; e += kLscThreads) {
        const int c0 = (int)fastdiv((uint32_t)e, mg), m = e - c0 * (M - 1), c = c0 + nd;
        const size_t pr = (size_t)la * P.K + c;
        const LscPair q = s_pair[c];
        gjk::D3 hc[kP], v;
        lsc_gjk_load(s_init + m * kP * 3, S.pred_traj + ((size_t)s_nbr[c] * npt + m * kP) * 3, q, hc);
        int it = 0;
        if (gjk::hull_origin_short<kP>(hc, v, &it)) {
            lsc_gjk_finish(hc, v, q, m, S.lsc_normal + (pr * M + m) * 3, S.lsc_d + (pr * M + m) * kP, S.lsc_near + pr * M + m);
            it_sum += it;
        } else {
            s_hard[atomicAdd(&s_nhard, 1)] = lsc_pack(la, c, m);
        }
    }
    const int tot = __reduce_add_sync(0xffffffffu, it_sum);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(S.counters + 1, (unsigned long long)tot);
    __syncthreads();
    const int nh = s_nhard;
    if (threadIdx.x == 0) {
        s_base = nh ? (int)atomicAdd(S.counters + kCntHard, (unsigned long long)nh) : 0;
        s_seg = cnt > nd ? (int)atomicAdd(S.counters + kCntSeg, (unsigned long long)(cnt - nd)) : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nh; i += kLscThreads) S.lsc_queue[(size_t)P.NL * P.K + s_base + i] = s_hard[i];
    for (int c = nd + threadIdx.x; c < cnt; c += kLscThreads) S.lsc_queue[s_seg + c - nd] = lsc_pack(la, c, M - 1);
}

__global__ void __launch_bounds__(kLscRestThreads, 8) k_lsc_rest(const __grid_constant__ DevParams P, const __grid_constant__ DevState S) {
    const int M = P.M, npt = M * kP, og = npt * 3 + 6;
    const int n_seg = (int)S.counters[kCntSeg], n_hard = (int)S.counters[kCntHard];
    const int stride = gridDim.x * kLscRestThreads, t0 = blockIdx.x * kLscRestThreads + threadIdx.x;
    for (int i = t0; i < n_seg; i += stride) {
        const uint32_t e = S.lsc_queue[i];
        const int la = (int)(e >> 14), c = (int)((e >> 4) & 1023u);
        const size_t pr = (size_t)la * P.K + c;
        const int j = S.nbr_idx[pr];
        const float* rec_a = S.rec + (size_t)(P.begin + la) * P.rec;
        const float* rec_j = S.rec + (size_t)j * P.rec;
        const LscPair q = lsc_pair_consts(S.radius[la], S.downwash[la], rec_j[og + 3], rec_j[og + 4]);
        lsc_last_segment(P, S.init_traj + (size_t)la * npt * 3, S.pred_traj + (size_t)j * npt * 3, v3_load(rec_a + og), v3_load(rec_j + og), q,
                         S.lsc_normal + (pr * M + (M - 1)) * 3, S.lsc_d + (pr * M + (M - 1)) * kP, S.lsc_anchor_last + pr * 3,
                         S.lsc_near + pr * M + (M - 1));
    }
    int it_sum = 0;
    const uint32_t* hard = S.lsc_queue + (size_t)P.NL * P.K;
    for (int i = t0; i < n_hard; i += stride) {
        const uint32_t e = hard[i];
        const int la = (int)(e >> 14), c = (int)((e >> 4) & 1023u), m = (int)(e & 15u);
        const size_t pr = (size_t)la * P.K + c;
        const int j = S.nbr_idx[pr];
        const float* rec_j = S.rec + (size_t)j * P.rec;
        const LscPair q = lsc_pair_consts(S.radius[la], S.downwash[la], rec_j[og + 3], rec_j[og + 4]);
        gjk::D3 hc[kP];
        lsc_gjk_load(S.init_traj + ((size_t)la * npt + m * kP) * 3, S.pred_traj + ((size_t)j * npt + m * kP) * 3, q, hc);
        int it = 0;
        const gjk::D3 v = gjk::hull_origin<kP>(hc, &it);
        lsc_gjk_finish(hc, v, q, m, S.lsc_normal + (pr * M + m) * 3, S.lsc_d + (pr * M + m) * kP, S.lsc_near + pr * M + m);
        it_sum += it;
    }
    const int tot = __reduce_add_sync(0xffffffffu, it_sum);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(S.counters + 1, (unsigned long long)tot);
}

// dynamic obstacles: thread per (agent, obstacle, segment)
__global__ void __launch_bounds__(128) k_lsc_dyn(const __grid_constant__ DevParams P, const __grid_constant__ DevState S) {
    const int M = P.M, npt = M * kP, nd = P.n_dyn;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.NL * nd * M) return;
    const int la = t / (nd * M), r = t - la * nd * M, o = r / M, m = r - o * M;
    const size_t pr = (size_t)la * P.K + o;
    lsc_dynamic_segment(S.init_traj + ((size_t)la * npt + m * kP) * 3, S.pred_traj + ((size_t)(P.N + o) * npt + m * kP) * 3,
                        S.dyn_size + ((size_t)o * M + m) * kP, S.radius[la], S.dyn_radius[o], S.dyn_downwash[o],
                        S.lsc_normal + (pr * M + m) * 3, S.lsc_d + (pr * M + m) * kP, S.lsc_near + pr * M + m);
}

int launch_lsc(const DevParams& P, const DevState& S, cudaStream_t st) {
    static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
    const size_t lsc_smem = (size_t)P.K * (sizeof(LscPair) + (size_t)P.M * sizeof(int));
    if (P.n_dyn > 0) k_lsc<true><<<P.NL, kLscThreads, lsc_smem, st>>>(P, S);
    else k_lsc<false><<<P.NL, kLscThreads, lsc_smem, st>>>(P, S);
    k_lsc_rest<<<sms * 8, kLscRestThreads, 0, st>>>(P, S);
    if (P.n_dyn > 0) {
        const int n = P.NL * P.n_dyn * P.M;
        k_lsc_dyn<<<(n + 127) / 128, 128, 0, st>>>(P, S);
        return 3;
    }
    return 2;
}

// ------------------------------------------------------------------------------------------------
// per-kernel parity entry (dlsc_gjk_batch): the device function k_lsc calls, one hull per thread, with the leaf tracer
__global__ void __launch_bounds__(128) k_gjk_batch(const double* __restrict__ pts, int n, double* __restrict__ v_out,
                                                   int32_t* __restrict__ iters, int32_t* __restrict__ simplex,
                                                   unsigned long long* __restrict__ leaves) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    gjk::D3 c[kP];
#pragma unroll
    for (int i = 0; i < kP; i++) c[i] = gjk::d3(pts[(size_t)h * 18 + 3 * i], pts[(size_t)h * 18 + 3 * i + 1], pts[(size_t)h * 18 + 3 * i + 2]);
    gjk::MaskTrace tr; tr.m = 0;
    int it = 0, sn = 0;
    const gjk::D3 v = gjk::hull_origin<kP, gjk::MaskTrace>(c, &it, tr, &sn);
    v_out[(size_t)h * 3] = v.x; v_out[(size_t)h * 3 + 1] = v.y; v_out[(size_t)h * 3 + 2] = v.z;
    if (iters) iters[h] = it;
    if (simplex) simplex[h] = sn;
    if (leaves) leaves[h] = tr.m;
}
void launch_gjk_batch(const double* pts, int n, double* v, int32_t* iters, int32_t* simplex, unsigned long long* leaves, cudaStream_t st) {
    k_gjk_batch<<<(n + 127) / 128, 128, 0, st>>>(pts, n, v, iters, simplex, leaves);
}

// ------------------------------------------------------------------------------------------------
// one warp per agent: the greedy control flow is replicated in every lane (uniform), the lattice columns of
// a box test are spread over the lanes (one 16-byte load of the vertex mask = 16 vertices) and combined with
// a warp vote.  4096 agents = 4096 warps: the whole swarm is resident in one wave.
#ifndef DLSC_SFC_MINB
#define DLSC_SFC_MINB 28
#endif
// one warp per CTA: k_sfc's CTAs hold the whole register file of the SMs while they run, and the LSC kernels on the other
// stream get in only as CTAs retire -- agent by agent instead of four at a time (step 0.477 -> 0.473 ms)
#ifndef DLSC_SFC_WARPS
#define DLSC_SFC_WARPS 1
#endif
constexpr int kSfcWarps = DLSC_SFC_WARPS;
__global__ void __launch_bounds__(kSfcWarps * 32, DLSC_SFC_MINB) k_sfc(const __grid_constant__ DevParams P, const __grid_constant__ DevState S) {
    __shared__ SfcTab tabs[kSfcWarps];
    const int w = threadIdx.x >> 5;
    const int la = blockIdx.x * kSfcWarps + w;
    if (la >= P.NL) return;
    Group g; g.lane = threadIdx.x & 31; g.width = 32; g.block = false;
    const int npt = P.M * kP;
    const float* rec = S.rec + (size_t)(P.begin + la) * P.rec;
    const bool init = S.sfc_init[la] != 0 || S.disturbed[la] != 0;       // traj_planner.cpp:439, 693-695
    long long look[6] = {0, 0, 0, 0, 0, 0};
    const int st = sfc_agent(g, P, S.edt, init, v3_load(rec + npt * 3), S.init_traj + (size_t)la * npt * 3,
                             v3_load(rec + npt * 3 + 6), v3_load(S.waypoint + la * 3), S.radius[la],
                             S.max_vel[la], S.sfc + (size_t)la * P.M * 6, &tabs[w], look);
    const unsigned tot = __reduce_add_sync(0xffffffffu, (unsigned)look[0]);
    const unsigned tot_alg = __reduce_add_sync(0xffffffffu, (unsigned)look[4]);
    if (g.lane == 0) {
        S.sfc_init[la] = 0;
        if (st) atomicOr(S.status + la, st);
        atomicAdd(S.counters + 2, (unsigned long long)tot);
        if (look[1]) atomicAdd(S.counters + 5, (unsigned long long)look[1]);
        if (look[2]) atomicAdd(S.counters + 6, (unsigned long long)look[2]);
        if (look[3]) atomicAdd(S.counters + 7, (unsigned long long)look[3]);
        atomicAdd(S.counters + 8, (unsigned long long)tot_alg);
    }
}

void launch_sfc(const DevParams& P, const DevState& S, cudaStream_t st) {
    k_sfc<<<(P.NL + kSfcWarps - 1) / kSfcWarps, kSfcWarps * 32, 0, st>>>(P, S);
}

// ------------------------------------------------------------------------------------------------
// warp per agent: lanes over the rows of the 1-variable LP (max / any / all reductions are order-independent)
__global__ void __launch_bounds__(128) k_goal(const __grid_constant__ DevParams P, const __grid_constant__ DevState S) {
    const int la = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (la >= P.NL) return;
    Group g; g.lane = threadIdx.x & 31; g.width = 32; g.block = false;
    const int npt = P.M * kP;
    const float* rec = S.rec + (size_t)(P.begin + la) * P.rec;
    V3 goal = v3_load(rec + npt * 3 + 6);
    const size_t pr = (size_t)la * P.K + P.n_dyn;          // the goal LP skips the dynamic obstacles (goal_optimizer.cpp:176-178)
    const int st = goal_agent(g, P, S.disturbed[la] != 0, v3_load(rec + npt * 3), v3_load(S.waypoint + la * 3),
                              S.sfc + ((size_t)la * P.M + (P.M - 1)) * 6, S.nbr_cnt[la] - P.n_dyn,
                              S.lsc_normal + pr * P.M * 3, S.lsc_d + pr * P.M * kP, S.lsc_anchor_last + pr * 3, goal);
    if (g.lane == 0) {
        v3_store(S.goal_new + la * 3, goal);     // the record keeps the previous goal until the step is published:
                                                 // the other agents' LSCs of this step must see it (broadcast semantics)
        if (st) atomicOr(S.status + la, st);
    }
}

// checkWaypointTrap (P.n_dyn > 0 only): warp per agent, after LSC and SFC, before the goal stage
__global__ void __launch_bounds__(128) k_trap(const __grid_constant__ DevParams P, const __grid_constant__ DevState S) {
    const int la = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (la >= P.NL) return;
    Group g; g.lane = threadIdx.x & 31; g.width = 32; g.block = false;
    const int npt = P.M * kP;
    const float* rec = S.rec + (size_t)(P.begin + la) * P.rec;
    const size_t pr = (size_t)la * P.K;
    DynObs O; O.pos = S.dyn_pos; O.vel = S.dyn_vel; O.radius = S.dyn_radius; O.downwash = S.dyn_downwash; O.max_acc = S.dyn_max_acc; O.size = S.dyn_size;
    const int tr = waypoint_trap(g, P, O, v3_load(rec + npt * 3 + 6), v3_load(S.waypoint + la * 3),
                                 S.sfc + ((size_t)la * P.M + (P.M - 1)) * 6, S.comm_box + (size_t)la * 6, S.nbr_cnt[la],
                                 S.lsc_normal + pr * P.M * 3, S.lsc_d + pr * P.M * kP, S.lsc_anchor_last + pr * 3, S.radius[la]);
    if (g.lane == 0) S.trap[la] = (uint8_t)tr;
}
void launch_trap(const DevParams& P, const DevState& S, cudaStream_t st) {
    const long long n = (long long)P.NL * 32;
    k_trap<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(P, S);
}

void launch_goal(const DevParams& P, const DevState& S, cudaStream_t st) {
    const long long n = (long long)P.NL * 32;
    k_goal<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(P, S);
}

__global__ void k_goal_copy(const __grid_constant__ DevParams P, const __grid_constant__ DevState S) {
    const int la = blockIdx.x * blockDim.x + threadIdx.x;
    if (la >= P.NL) return;
    const float* rec = S.rec + (size_t)(P.begin + la) * P.rec + P.M * kP * 3 + 6;
    for (int k = 0; k < 3; k++) S.goal_new[la * 3 + k] = rec[k];
}
void launch_goal_copy(const DevParams& P, const DevState& S, cudaStream_t st) {
    k_goal_copy<<<(P.NL + 127) / 128, 128, 0, st>>>(P, S);
}

// ------------------------------------------------------------------------------------------------
// move: state := traj(dt) (AgentManager::doStep); always: record.traj := traj (prev_traj, traj_planner.cpp:57)
__global__ void __launch_bounds__(256) k_advance(const __grid_constant__ DevParams P, const __grid_constant__ DevState S, int move) {
    const int npt = P.M * kP, per = npt * 3;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P.NL * per) return;
    const int la = (int)(gid / per), e = (int)(gid - (long long)la * per);
    float* rec = S.rec + (size_t)(P.begin + la) * P.rec;
    const float* tr = S.traj + (size_t)la * per;
    rec[e] = tr[e];
    if (e == 0) {
        if (move) {
            float st[9];
            state_at(P, tr, P.dt, st);
            for (int k = 0; k < 3; k++) { rec[per + k] = st[k]; rec[per + 3 + k] = st[3 + k]; S.acc[la * 3 + k] = st[6 + k]; }
        }
        for (int k = 0; k < 3; k++) rec[per + 6 + k] = S.goal_new[la * 3 + k];
    }
}

void launch_advance(const DevParams& P, const DevState& S, bool move, cudaStream_t st) {
    const long long n = (long long)P.NL * P.M * kP * 3;
    k_advance<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, S, move ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------
__global__ void k_edt_pack(const float* dist, const int32_t* obst, int4* cells, size_t ncell) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncell) return;
    int4 r;
    r.x = __float_as_int(dist[i]); r.y = obst[3 * i]; r.z = obst[3 * i + 1]; r.w = obst[3 * i + 2];
    cells[i] = r;
}
void launch_edt_pack(const float* dist, const int32_t* obst, int4* cells, size_t ncell, cudaStream_t st) {
    k_edt_pack<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(dist, obst, cells, ncell);
}

// lattice-vertex mask of the SFC vertex test (dlsc_stages.cuh edt_vertex_mask): thread per (vx, vy, vz)
__global__ void __launch_bounds__(256) k_edt_mask(const __grid_constant__ EdtDev E, double margin, uint8_t* mask, int* unsafe) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)(E.dims[0] + 1) * (E.dims[1] + 1) * E.zs;
    if (i >= n) return;
    const int vz = (int)(i % E.zs);
    const size_t t = i / E.zs;
    const int vy = (int)(t % (E.dims[1] + 1)), vx = (int)(t / (E.dims[1] + 1));
    uint8_t b = 0;
    if (vz <= E.dims[2]) {
        bool u = false;
        b = edt_vertex_mask(E, vx, vy, vz, margin, &u);
        if (u) *unsafe = 1;
    }
    mask[i] = b;
}
void launch_edt_mask(const EdtDev& E, double margin, uint8_t* mask, int* unsafe, cudaStream_t st) {
    const size_t n = (size_t)(E.dims[0] + 1) * (E.dims[1] + 1) * E.zs;
    k_edt_mask<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(E, margin, mask, unsafe);
}

// summed-area table over the flagged lattice vertices (dlsc_stages.cuh sat_indicator): indicator, then one
// running-sum pass per axis (thread per line; lines of the y / x passes are coalesced across threads)
__global__ void __launch_bounds__(256) k_sat_init(const __grid_constant__ EdtDev E, int32_t* sat) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n1 = E.dims[1] + 2, n2 = E.dims[2] + 2;
    const size_t n = (size_t)(E.dims[0] + 2) * n1 * n2;
    if (i >= n) return;
    const int k = (int)(i % n2);
    const size_t t = i / n2;
    const int j = (int)(t % n1), ii = (int)(t / n1);
    sat[i] = (ii > 0 && j > 0 && k > 0) ? sat_indicator(E, ii - 1, j - 1, k - 1) : 0;
}
__global__ void __launch_bounds__(256) k_sat_scan(const __grid_constant__ EdtDev E, int32_t* sat, int axis) {
    const int n0 = E.dims[0] + 2, n1 = E.dims[1] + 2, n2 = E.dims[2] + 2;
    const size_t line = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t base, stride; int len;
    if (axis == 2) { if (line >= (size_t)n0 * n1) return; base = line * n2; stride = 1; len = n2; }
    else if (axis == 1) { if (line >= (size_t)n0 * n2) return; base = (line / n2) * n1 * n2 + line % n2; stride = n2; len = n1; }
    else { if (line >= (size_t)n1 * n2) return; base = line; stride = (size_t)n1 * n2; len = n0; }
    int acc = 0;
    for (int t = 0; t < len; t++) { acc += sat[base + t * stride]; sat[base + t * stride] = acc; }
}
void launch_sat_build(const EdtDev& E, int32_t* sat, cudaStream_t st) {
    const size_t n0 = E.dims[0] + 2, n1 = E.dims[1] + 2, n2 = E.dims[2] + 2, n = n0 * n1 * n2;
    k_sat_init<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(E, sat);
    k_sat_scan<<<(unsigned)((n0 * n1 + 255) / 256), 256, 0, st>>>(E, sat, 2);
    k_sat_scan<<<(unsigned)((n0 * n2 + 255) / 256), 256, 0, st>>>(E, sat, 1);
    k_sat_scan<<<(unsigned)((n1 * n2 + 255) / 256), 256, 0, st>>>(E, sat, 0);
}

// LSC anchors in the reference layout [NL][K][M][P][3]: predicted control points, or the segment-case
// witness point for the last segment (traj_planner.cpp:638, 657)
__global__ void k_expand_anchor(const __grid_constant__ DevParams P, const __grid_constant__ DevState S, float* out) {
    const int npt = P.M * kP;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)P.NL * P.K * npt;
    if (gid >= total) return;
    const int pt = (int)(gid % npt);
    const long long pr = gid / npt;
    const int c = (int)(pr % P.K), la = (int)(pr / P.K);
    float v[3] = {0.f, 0.f, 0.f};
    if (c < S.nbr_cnt[la]) {
        const int j = S.nbr_idx[(size_t)la * P.K + c];
        const float* src = (pt / kP < P.M - 1 || c < P.n_dyn) ? S.pred_traj + ((size_t)j * npt + pt) * 3 : S.lsc_anchor_last + (size_t)pr * 3;
        v[0] = src[0]; v[1] = src[1]; v[2] = src[2];
    }
    out[gid * 3] = v[0]; out[gid * 3 + 1] = v[1]; out[gid * 3 + 2] = v[2];
}
void launch_expand_anchor(const DevParams& P, const DevState& S, float* anchor_out, cudaStream_t st) {
    const long long n = (long long)P.NL * P.K * P.M * kP;
    k_expand_anchor<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, S, anchor_out);
}

// planner state reset (agent_manager.cpp:4-32): record = {traj: start everywhere, pos = start, vel = 0,
// goal = start, radius, downwash}; waypoint = start; acc = 0; initialize_sfc = true
__global__ void k_reset(const __grid_constant__ DevParams P, const __grid_constant__ DevState S, const float* start) {
    const int la = blockIdx.x * blockDim.x + threadIdx.x;
    if (la >= P.NL) return;
    const int npt = P.M * kP;
    float* rec = S.rec + (size_t)(P.begin + la) * P.rec;
    const float s0 = start[la * 3], s1 = start[la * 3 + 1], s2 = start[la * 3 + 2];
    for (int e = 0; e < npt; e++) { rec[e * 3] = s0; rec[e * 3 + 1] = s1; rec[e * 3 + 2] = s2; }
    const int o = npt * 3;
    rec[o] = s0; rec[o + 1] = s1; rec[o + 2] = s2;
    rec[o + 3] = 0.f; rec[o + 4] = 0.f; rec[o + 5] = 0.f;
    rec[o + 6] = s0; rec[o + 7] = s1; rec[o + 8] = s2;
    rec[o + 9] = (float)S.radius[la]; rec[o + 10] = (float)S.downwash[la];
    for (int e = o + 11; e < P.rec; e++) rec[e] = 0.f;
    for (int k = 0; k < 3; k++) { S.acc[la * 3 + k] = 0.f; S.waypoint[la * 3 + k] = start[la * 3 + k]; S.goal_new[la * 3 + k] = start[la * 3 + k]; }
    S.disturbed[la] = 0; S.sfc_init[la] = 1; S.status[la] = 0;
    for (int e = 0; e < 6; e++) S.comm_box[(size_t)la * 6 + e] = 0.f;
    for (int e = 0; e < npt * 3; e++) S.traj[(size_t)la * npt * 3 + e] = rec[e];
}
void launch_reset(const DevParams& P, const DevState& S, const float* start_dev, cudaStream_t st) {
    k_reset<<<(P.NL + 127) / 128, 128, 0, st>>>(P, S, start_dev);
}

// ------------------------------------------------------------------------------------------------
// dlsc_set_agents: positions / velocities arrive as dense [NL][3] arrays (one contiguous H2D copy each into a
// staging buffer) and are scattered into the 768-byte records here -- a strided 2-D copy would cost the DMA engine one
// 12-byte transaction per agent.
__global__ void k_set_state(const __grid_constant__ DevParams P, float* __restrict__ rec, const float* __restrict__ pos,
                            const float* __restrict__ vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.NL * 3) return;
    const int la = i / 3, k = i - la * 3;
    float* r = rec + (size_t)(P.begin + la) * P.rec + P.M * kP * 3;
    if (pos) r[k] = pos[i];
    if (vel) r[3 + k] = vel[i];
}
__global__ void k_get_state(const __grid_constant__ DevParams P, const float* __restrict__ rec, float* __restrict__ pos,
                            float* __restrict__ vel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.NL * 3) return;
    const int la = i / 3, k = i - la * 3;
    const float* r = rec + (size_t)(P.begin + la) * P.rec + P.M * kP * 3;
    pos[i] = r[k];
    vel[i] = r[3 + k];
}
void launch_get_state(const DevParams& P, const float* rec, float* pos, float* vel, cudaStream_t st) {
    k_get_state<<<(P.NL * 3 + 255) / 256, 256, 0, st>>>(P, rec, pos, vel);
}
void launch_set_state(const DevParams& P, float* rec, const float* pos, const float* vel, cudaStream_t st) {
    k_set_state<<<(P.NL * 3 + 255) / 256, 256, 0, st>>>(P, rec, pos, vel);
}

// ------------------------------------------------------------------------------------------------
// Record exchange over NVLink peer memory (one process per GPU, dlsc_p2p_*): the all-gather of the agent records as
// direct stores.  k_p2p_push copies this rank's slice of the records into the *next* record buffer of every rank
// (its own included); the last CTA to finish publishes the step number in every rank's flag slot (system-scope fence
// before, so the data is visible first).  k_p2p_wait spins until every rank's flag has reached the step.  Records are
// double buffered: a rank that runs ahead writes the buffer its peers are not reading.
struct P2PPeers { float* dst[kP2PMaxWorld]; unsigned long long* flag[kP2PMaxWorld]; };

__global__ void __launch_bounds__(256) k_p2p_push(const float4* __restrict__ src, size_t n4, size_t slice_off4, int world, int rank,
                                                  unsigned long long step, const __grid_constant__ P2PPeers peers, unsigned* done) {
    for (int r = 0; r < world; r++) {
        float4* dst = reinterpret_cast<float4*>(peers.dst[r]) + slice_off4;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {                       // every CTA's stores are fenced: publish
            *done = 0;
            __threadfence_system();
            for (int r = 0; r < world; r++) {
                volatile unsigned long long* f = peers.flag[r] + rank;
                *f = step;
            }
            __threadfence_system();
        }
    }
}

__global__ void k_p2p_wait(const unsigned long long* flags, int world, unsigned long long step, long long timeout_cycles, int* err) {
    const int r = threadIdx.x;
    if (r >= world) return;
    const volatile unsigned long long* f = flags + r;
    const long long t0 = clock64();
    while (*f < step) {
        if (clock64() - t0 > timeout_cycles) { *reinterpret_cast<volatile int*>(err) = 1; break; }    // err: mapped host memory
        __nanosleep(200);
    }
    __threadfence_system();
}

void launch_p2p_push(const float* src, size_t n_floats, size_t slice_off_floats, int world, int rank, unsigned long long step,
                     float* const* dst, unsigned long long* const* flag, unsigned* done, cudaStream_t st) {
    P2PPeers pp;
    for (int r = 0; r < world; r++) { pp.dst[r] = dst[r]; pp.flag[r] = flag[r]; }
    const size_t n4 = n_floats / 4;
    int ctas = (int)((n4 + 255) / 256);
    if (ctas > 64) ctas = 64;
    if (ctas < 1) ctas = 1;
    k_p2p_push<<<ctas, 256, 0, st>>>(reinterpret_cast<const float4*>(src), n4, slice_off_floats / 4, world, rank, step, pp, done);
}
void launch_p2p_wait(const unsigned long long* flags, int world, unsigned long long step, int* err, cudaStream_t st) {
    k_p2p_wait<<<1, 32, 0, st>>>(flags, world, step, 20000000000LL, err);   // ~10 s
}

}  // namespace dlsc
