// dlsc_kernels_qp.cu -- sm_100a kernel of the batched min-jerk QP (interior point, FP64 pipe).
// Persistent CTAs: grid = SMs x resident CTAs, each CTA pulls agents from a device-side counter
// (agents differ widely in neighbour count, hence in row count) and reuses one global scratch slab.
#include <cstdlib>

#include "dlsc_kernels.h"
#include "dlsc_qp_gi.cuh"

namespace dlsc {

constexpr int kQpThreads = 128;     // interior-point fallback kernel
#ifndef DLSC_QP_THREADS_DYN
#define DLSC_QP_THREADS_DYN 512
#endif
#ifndef DLSC_GI_THREADS
#define DLSC_GI_THREADS 128
#endif
constexpr int kGiThreads = DLSC_GI_THREADS;
#ifndef DLSC_GI_MINB
#define DLSC_GI_MINB 6
#endif
constexpr int kGiBlocksPerSm = DLSC_GI_MINB;

__device__ __forceinline__ void qp_load_agent(const DevParams& P, const DevState& S, const QpTab& T, int la, QpIn& in, QpOut& out) {
    const int npt = P.M * kP;
    const float* rec = S.rec + (size_t)(P.begin + la) * P.rec;
    in.pos = v3_load(rec + npt * 3); in.vel = v3_load(rec + npt * 3 + 3);
    in.acc = v3_load(S.acc + la * 3); in.goal = v3_load(S.goal_new + la * 3);
    in.wp = v3_load(S.waypoint + la * 3);
    in.radius = S.radius[la]; in.max_vel = S.max_vel[la]; in.max_acc = S.max_acc[la];
    in.nominal_vel = S.nominal_vel[la];
    in.sfc = S.sfc + (size_t)la * P.M * 6;
    in.init_traj = S.init_traj + (size_t)la * npt * 3;
    in.K = S.nbr_cnt[la];
    const size_t pr = (size_t)la * P.K;
    in.nbr_idx = S.nbr_idx + pr;
    in.normal = S.lsc_normal + pr * P.M * 3;
    in.d = S.lsc_d + pr * P.M * kP;
    in.anchor_last = S.lsc_anchor_last + pr * 3;
    in.pred_traj = S.pred_traj;
    in.near = S.lsc_near + pr * P.M;
    out.traj = S.traj + (size_t)la * npt * 3;
    out.traj_host = S.traj_host ? S.traj_host + (size_t)la * npt * 3 : nullptr;
    out.x = S.qp_x + (size_t)la * T.nx;
    out.cost = S.cost + la; out.viol = S.viol + la; out.iters = S.qp_iters + la; out.status = S.status + la;
    out.rows = nullptr;
    out.slack = P.n_dyn > 0 ? S.qp_slack + (size_t)la * kMaxDyn * P.M : nullptr;
}

// Fast path: one warp per agent, no block-level synchronisation.  Agents whose unconstrained optimum is feasible
// (most of a swarm in transit) are finished here; the others are queued, with the violated row found, for k_qp_gi.
constexpr int kFastWarps = 4;
#ifndef DLSC_GI_HEAVY_ROWS
#define DLSC_GI_HEAVY_ROWS 3
#endif
constexpr int kGiHeavyRows = DLSC_GI_HEAVY_ROWS;
#ifndef DLSC_FAST_MINB
#define DLSC_FAST_MINB 7
#endif
// DYN (here and in k_qp_gi): dynamic obstacles present.  The DYN = false instantiations are the swarm-only hot path,
// compiled without the slack-variable code.
template <bool DYN>
__global__ void __launch_bounds__(kFastWarps * 32, DLSC_FAST_MINB) k_qp_fast(const __grid_constant__ DevParams P,
                                                                             const __grid_constant__ DevState S,
                                                                             const __grid_constant__ QpTab T, int per_warp_doubles) {
    extern __shared__ __align__(16) double smem[];
    const int w = threadIdx.x >> 5;
    const int la = blockIdx.x * kFastWarps + w;
    if (la >= P.NL) return;
    QpSmem sm;
    fast_smem_carve(T, smem + (size_t)w * per_warp_doubles, sm);
    Cta c; c.tid = threadIdx.x & 31; c.nthr = 32; c.red = nullptr; c.warp = true;
    QpIn in; QpOut out;
    qp_load_agent(P, S, T, la, in, out);
    long long rows = 0;
    if (c.tid == 0) out.rows = &rows;
    double* seed = S.qp_seed + (size_t)la * 4;
    const bool done = qp_agent_fast<DYN>(c, P, T, in, out, sm, seed);
    if (c.tid == 0) {
        if (!done) {
            // longest-job-first: agents with several violated rows (they iterate longest) are queued from the front
            // of the list and started first, the others from the back; k_qp_gi's runtime is set by its tail
            if (seed[3] >= (double)kGiHeavyRows) S.qp_list_gi[atomicAdd(S.qp_next + 3, 1)] = la;
            else S.qp_list_gi[P.NL - 1 - atomicAdd(S.qp_next + 0, 1)] = la;
        } else atomicAdd(S.counters + 4, (unsigned long long)rows);
    }
}

// Dual active set (dlsc_qp_gi.cuh): one small CTA per queued agent, straight off the constraint arrays.
// Agents it cannot finish are appended to S.qp_list for k_qp.  from_list = 0: every agent, no seed.
template <bool DYN>
__global__ void __launch_bounds__(kGiThreads, DLSC_GI_MINB) k_qp_gi(const __grid_constant__ DevParams P, const __grid_constant__ DevState S,
                                                                     const __grid_constant__ QpTab T, int from_list) {
    extern __shared__ __align__(16) double smem[];
    int la = blockIdx.x;
    if (from_list) {
        const int n_heavy = S.qp_next[3], n_light = S.qp_next[0], b = blockIdx.x;
        if (b >= n_heavy + n_light) return;
        la = (b < n_heavy) ? S.qp_list_gi[b] : S.qp_list_gi[P.NL - 1 - (b - n_heavy)];
    }
    QpSmem sm;
    gi_smem_carve(T, smem, sm);
    Cta c; c.tid = threadIdx.x; c.nthr = blockDim.x; c.red = sm.red;
#ifdef DLSC_QP_CYCLES
    const long long t_begin = clock64();
    long long ticks[12] = {t_begin, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (threadIdx.x == 0) c.ticks = ticks;
#endif
    QpIn in; QpOut out;
    qp_load_agent(P, S, T, la, in, out);
    long long rows = 0;
    if (threadIdx.x == 0) out.rows = &rows;
    const bool done = qp_agent_gi<kGiQ, DYN>(c, P, T, in, out, sm, from_list ? S.qp_seed + (size_t)la * 4 : nullptr);
#ifdef DLSC_QP_CYCLES
    if (threadIdx.x == 0) {                                                 // diagnostic build only (scripts/qp_hist.py)
        S.viol[la] = (double)(clock64() - t_begin);
        // phase sums over the agents of this kernel: counters 9..15 <- prologue (1+2), map_x (3), far2 (4), scan (5),
        // candidate + Hinv staging (8), w = H^-1 a (9), serial step of thread 0 (10) ; y update (11) goes to counter 0 + 64-bit pack
#ifdef DLSC_QP_CYCLES_HEAVY
        if (S.qp_iters[la] >= DLSC_QP_CYCLES_HEAVY)
#endif
        {
        atomicAdd(S.counters + 9, (unsigned long long)(ticks[1] + ticks[2]));
        atomicAdd(S.counters + 10, (unsigned long long)ticks[3]);
        atomicAdd(S.counters + 11, (unsigned long long)ticks[5]);
        atomicAdd(S.counters + 12, (unsigned long long)ticks[8]);
        atomicAdd(S.counters + 13, (unsigned long long)ticks[9]);
        atomicAdd(S.counters + 14, (unsigned long long)ticks[10]);
        atomicAdd(S.counters + 15, (unsigned long long)ticks[11]);
        }
    }
#endif
    if (threadIdx.x == 0) {
        if (!done) S.qp_list[atomicAdd(S.qp_next + 1, 1)] = la;
        else {
            const int it = S.qp_iters[la];
            if (it) atomicAdd(S.counters + 3, (unsigned long long)it);
            atomicAdd(S.counters + 4, (unsigned long long)rows);
        }
    }
}

// Fallback kernel (interior point, dlsc_qp.cuh): persistent CTAs pull agents from S.qp_list (or, with
// all_agents, every agent: qp_solver = 1).
// DYN (dynamic obstacles): the kernel then serves a handful of slack-heavy agents with ~3000-row working sets, and the step
// waits for the slowest of them: four times the threads per agent (the row passes and the trailing updates scale with them: 8.0 / 6.1 / 5.0 ms per step at 128 / 256 / 512)
constexpr int kQpThreadsDyn = DLSC_QP_THREADS_DYN;
template <bool DYN>
__global__ void __launch_bounds__(DYN ? kQpThreadsDyn : kQpThreads, DYN ? (kQpThreadsDyn >= 512 ? 1 : 512 / kQpThreadsDyn) : 4) k_qp(const __grid_constant__ DevParams P,
                                                   const __grid_constant__ DevState S,
                                                   const __grid_constant__ QpTab T, size_t scratch_doubles, int all_agents) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_agent;
    const int n_list = all_agents ? P.NL : S.qp_next[1];
    if (n_list == 0) return;
    QpSmem sm;
    qp_smem_carve(T, P.K, smem, sm);
    Cta c; c.tid = threadIdx.x; c.nthr = blockDim.x; c.red = sm.red;
    double* scratch = S.qp_scratch + (size_t)blockIdx.x * scratch_doubles;
    unsigned long long it_sum = 0, row_sum = 0;
    for (;;) {
        if (threadIdx.x == 0) s_agent = atomicAdd(S.qp_next + 2, 1);
        __syncthreads();
        const int i = s_agent;
        __syncthreads();
        if (i >= n_list) break;
        const int la = all_agents ? i : S.qp_list[i];
        QpIn in; QpOut out;
        qp_load_agent(P, S, T, la, in, out);
        long long rows = 0;
        if (threadIdx.x == 0) out.rows = &rows;
        qp_agent<DYN>(c, P, T, in, out, sm, scratch, !all_agents);
        if (threadIdx.x == 0) { it_sum += (unsigned long long)S.qp_iters[la]; row_sum += (unsigned long long)rows; }
    }
    if (threadIdx.x == 0) {
        if (it_sum) atomicAdd(S.counters + 3, it_sum);
        if (row_sum) atomicAdd(S.counters + 4, row_sum);
    }
}

QpLaunch qp_launch_config(const DevParams& P, const QpTab& T, int device) {
    QpLaunch L;
    L.threads = kQpThreads;
    L.smem = qp_smem_bytes(T, P.K);
    cudaFuncSetAttribute(k_qp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    cudaFuncSetAttribute(k_qp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    L.gi_smem = gi_smem_doubles(T, P.K) * sizeof(double);
    cudaFuncSetAttribute(k_qp_gi<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.gi_smem);
    cudaFuncSetAttribute(k_qp_gi<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.gi_smem);
    L.fast_smem = ((fast_smem_doubles(T, P.K) + 1) / 2 * 2) * sizeof(double) * kFastWarps;
    cudaFuncSetAttribute(k_qp_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.fast_smem);
    cudaFuncSetAttribute(k_qp_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.fast_smem);
    int per_sm = 1, sms = 148;
    L.sms = 148;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_qp<false>, kQpThreads, L.smem);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    L.sms = sms;
    if (per_sm < 1) per_sm = 1;
    L.ctas = sms * per_sm;
    L.scratch_doubles = qp_scratch_doubles(T, P.K);
    return L;
}

// returns the number of kernels launched
int launch_qp(const DevParams& P, const DevState& S, const QpTab& T, const QpLaunch& L, cudaStream_t st) {
    cudaMemsetAsync(S.qp_next, 0, 4 * sizeof(int), st);
    const int ctas = L.ctas < P.NL ? L.ctas : P.NL;
    const int all = (P.qp_solver == 1) ? 1 : 0;
    int n = 1;
    if (!all) {
        // The warp-per-agent fast path pays when the agents fill the GPU several times over (4096 agents: 0.165 against
        // 0.185 ms for the QP stage).  A small block -- one rank's share of a sharded swarm -- fits the CTA-per-agent kernel
        // in up to three waves, and there a whole CTA per first scan is quicker than one warp (512 agents: 0.141 against
        // 0.173 ms; 1024: 0.081 against 0.110; 2048 per rank on 2 GPUs: 0.108 against 0.134; 3072: a tie).  qp_solver = 2 / 3 (or DLSC_QP_FAST=1 / 0) forces either way.
        static const int fast_env = [] { const char* e = getenv("DLSC_QP_FAST"); return !e ? -1 : (e[0] == '0' ? 0 : 1); }();
        const bool no_fast = P.qp_solver == 2 ? false : P.qp_solver == 3 ? true
                             : fast_env >= 0 ? fast_env == 0 : P.NL <= 3 * kGiBlocksPerSm * L.sms;
        const bool dyn = P.n_dyn > 0;
        const int fast_grid = (P.NL + kFastWarps - 1) / kFastWarps, per_warp = (int)(L.fast_smem / sizeof(double) / kFastWarps);
        if (no_fast) {
            if (dyn) k_qp_gi<true><<<P.NL, kGiThreads, L.gi_smem, st>>>(P, S, T, 0);
            else k_qp_gi<false><<<P.NL, kGiThreads, L.gi_smem, st>>>(P, S, T, 0);
            n = 2;
        } else if (dyn) {
            k_qp_fast<true><<<fast_grid, kFastWarps * 32, L.fast_smem, st>>>(P, S, T, per_warp);
            k_qp_gi<true><<<P.NL, kGiThreads, L.gi_smem, st>>>(P, S, T, 1);
            n = 3;
        } else {
            k_qp_fast<false><<<fast_grid, kFastWarps * 32, L.fast_smem, st>>>(P, S, T, per_warp);
            k_qp_gi<false><<<P.NL, kGiThreads, L.gi_smem, st>>>(P, S, T, 1);
            n = 3;
        }
    }
    if (P.n_dyn > 0) k_qp<true><<<ctas, kQpThreadsDyn, L.smem, st>>>(P, S, T, L.scratch_doubles, all);
    else k_qp<false><<<ctas, L.threads, L.smem, st>>>(P, S, T, L.scratch_doubles, all);
    return n;
}

// ---- FP64 FMA peak probe: 8 independent dependent-FMA chains per thread, 1024 threads, 2 CTAs per SM ----
__global__ void __launch_bounds__(1024) k_fp64_peak(double* sink, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, b = 1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456) sink[0] = r;
}

double measure_fp64_peak(int device, cudaStream_t st) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    double* sink = nullptr;
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, ctas = sms * 2;
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, st);
        k_fp64_peak<<<ctas, 1024, 0, st>>>(sink, iters);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * iters * 1024.0 * ctas / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    return best;
}

}  // namespace dlsc
