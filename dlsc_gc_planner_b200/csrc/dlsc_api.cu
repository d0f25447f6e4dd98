// dlsc_api.cu -- context management and the C ABI of libdlsc_b200.so (include/dlsc_b200.h).
// No CPU fallback: every entry point requires a usable CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dlsc_b200.h"
#include "dlsc_kernels.h"
#include "dlsc_stages.cuh"
#include "dlsc_qp_tables.h"
#include "dlsc_types.h"

using namespace dlsc;

static thread_local std::string g_err;

struct dlsc_ctx {
    dlsc_params hp;
    DevParams P;
    DevState S;
    RecLayout rl;
    QpTabHost th;
    QpTab T;
    QpLaunch qpl;
    void* tab_blob = nullptr;
    int device = 0;
    int seq = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // SFC runs beside neighbour search + LSC (both only need the prediction stage): fork / join on a second stream
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap = true;               // DLSC_OVERLAP=0 disables; per-stage timing (dlsc_enable_timing) serialises too
    bool rec_owned = false;
    void* traj_host_registered = nullptr;   // host buffer pinned by dlsc_bind_traj_host (unregistered on rebind / destroy)
    bool have_edt = false;
    int4* edt_cells = nullptr;
    float* edt_centre = nullptr;
    int32_t* edt_sat = nullptr;        // summed-area table over the flagged mask vertices
    uint8_t* edt_mask = nullptr;       // lattice-vertex mask of the SFC vertex test (built lazily for margin_host)
    bool mask_dirty = false;
    // record exchange over peer memory (dlsc_p2p_*): one shared block per rank = [flags 256 B][records x 2]
    struct {
        bool exported = false, on = false;
        int world = 1, rank = 0, cur = 0;
        unsigned long long step = 0;
        void* block = nullptr;                     // own block (cudaMalloc, IPC-exported)
        void* peer_block[kP2PMaxWorld] = {};       // mapped blocks of the peers (own entry = block)
        unsigned* done = nullptr;
        int* err = nullptr;                        // device view of the error flag (mapped pinned host memory)
        volatile int* err_host = nullptr;          // host view: polled without synchronising
        bool failed = false;                       // sticky
    } p2p;
    // CUDA graphs of one whole dlsc_step (all stages, both streams): replayed instead of ~12 separate launches.  Kernel
    // parameters are baked in, so an entry is keyed by the DevParams / DevState bytes it was captured with (the record
    // pointer alternates between the two peer-memory buffers: two entries).
    struct StepGraph { cudaGraphExec_t exec = nullptr; DevParams P; DevState S; int kernels = 0; };
    StepGraph graphs[2];
    int graph_next = 0;
    bool use_graph = true;             // DLSC_GRAPH=0 disables
    double edt_build_ms = 0.0;         // device time of the last dlsc_build_edt* (the three EDT passes)
    double margin_host = 0.0;          // radius of the first local agent (all BASELINE missions: 0.15 for every agent)
    int64_t launches = 0;
    bool timing = false;
    std::vector<cudaEvent_t> evpool;   // (DLSC_N_STAGES + 1) events per timed step, resolved lazily
    int ev_used = 0;                   // steps recorded since the last dlsc_get_timings
    double t_ms[DLSC_N_STAGES];
    int t_steps = 0;
    std::vector<void*> allocs;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return -1;                                                                             \
        }                                                                                          \
    } while (0)

static int fail(const char* msg) { g_err = msg; return -1; }
extern "C" { static int d2h_raw(dlsc_ctx* c, void* host, const void* dev, size_t bytes); }

template <class T>
static int dev_alloc(dlsc_ctx* c, T** p, size_t n) {
    void* q = nullptr;
    CK(cudaMalloc(&q, (n ? n : 1) * sizeof(T)));
    CK(cudaMemset(q, 0, (n ? n : 1) * sizeof(T)));
    c->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

extern "C" {

const char* dlsc_last_error(void) { return g_err.c_str(); }
int dlsc_abi_version(void) { return DLSC_ABI_VERSION; }
// only the CUDA build exports this: the Python loader refuses any other library handed to it as the product
int dlsc_cuda_build(void) { return 12090; }
int dlsc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static int upload_tables(dlsc_ctx* c) {
    const QpTabHost& h = c->th;
    // serialise every table into one blob (256-byte aligned sections)
    struct Sec { const void* src; size_t bytes; size_t off; };
    std::vector<Sec> secs;
    size_t off = 0;
    auto add = [&](const void* src, size_t bytes) {
        Sec s{src, bytes, off};
        secs.push_back(s);
        off += (bytes + 255) / 256 * 256;
        return secs.size() - 1;
    };
#define SEC(v) add(h.v.data(), h.v.size() * sizeof(h.v[0]))
    const size_t i_xm_nv = SEC(xm_nv), i_xm_cidx = SEC(xm_cidx), i_xm_idx = SEC(xm_idx), i_xm_coef = SEC(xm_coef);
    const size_t i_pr_fam = SEC(pr_fam), i_pr_axis = SEC(pr_axis), i_pr_nnz = SEC(pr_nnz), i_pr_pt = SEC(pr_pt);
    const size_t i_pr_idx = SEC(pr_idx), i_pr_val = SEC(pr_val), i_pr_cc = SEC(pr_cc);
    const size_t i_yi_ptr = SEC(yi_ptr), i_yi_row = SEC(yi_row), i_yi_coef = SEC(yi_coef);
    const size_t i_yp_ptr = SEC(yp_ptr), i_yp_pt = SEC(yp_pt), i_yp_coef = SEC(yp_coef);
    const size_t i_wi_ptr = SEC(wi_ptr), i_wi_row = SEC(wi_row), i_wi_coef = SEC(wi_coef);
    const size_t i_wp_ptr = SEC(wp_ptr), i_wp_pt = SEC(wp_pt), i_wp_coef = SEC(wp_coef);
    const size_t i_H1 = SEC(H1), i_Q2 = SEC(Q2), i_tri = SEC(tri_p), i_nz = SEC(nz_e);
    const size_t i_desc = SEC(pr_desc), i_hdr = SEC(nz_hdr), i_nzh = SEC(nz_h), i_hinv = SEC(Hinv), i_y0 = SEC(Y0);
#undef SEC
    std::vector<char> blob(off ? off : 256, 0);
    for (auto& s : secs) memcpy(blob.data() + s.off, s.src, s.bytes);
    CK(cudaMalloc(&c->tab_blob, blob.size()));
    CK(cudaMemcpy(c->tab_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    char* base = (char*)c->tab_blob;
    QpTab& T = c->T;
    T.D = h.D; T.M = h.M; T.nyd = h.nyd; T.ny = h.ny; T.npt = h.npt; T.nx = h.nx; T.np = h.np; T.ntri = h.ntri;
    T.ntri_local = h.ntri_local; T.use_comm = h.use_comm;
#define PTR(type, i) (const type*)(base + secs[i].off)
    T.xm_nv = PTR(int8_t, i_xm_nv); T.xm_cidx = PTR(int8_t, i_xm_cidx); T.xm_idx = PTR(int16_t, i_xm_idx);
    T.xm_coef = PTR(double, i_xm_coef);
    T.pr_fam = PTR(uint8_t, i_pr_fam); T.pr_axis = PTR(uint8_t, i_pr_axis); T.pr_nnz = PTR(uint8_t, i_pr_nnz);
    T.pr_pt = PTR(int16_t, i_pr_pt); T.pr_idx = PTR(int16_t, i_pr_idx); T.pr_val = PTR(double, i_pr_val);
    T.pr_cc = PTR(double, i_pr_cc);
    T.yi_ptr = PTR(int, i_yi_ptr); T.yi_row = PTR(int16_t, i_yi_row); T.yi_coef = PTR(double, i_yi_coef);
    T.yp_ptr = PTR(int, i_yp_ptr); T.yp_pt = PTR(int16_t, i_yp_pt); T.yp_coef = PTR(double, i_yp_coef);
    T.wi_ptr = PTR(int, i_wi_ptr); T.wi_row = PTR(int16_t, i_wi_row); T.wi_coef = PTR(double, i_wi_coef);
    T.wp_ptr = PTR(int, i_wp_ptr); T.wp_pt = PTR(int16_t, i_wp_pt); T.wp_coef = PTR(double, i_wp_coef);
    T.H1 = PTR(double, i_H1); T.Q2 = PTR(double, i_Q2); T.tri_p = PTR(uint8_t, i_tri);
    T.nz_e = PTR(uint16_t, i_nz); T.nnzw = h.nnzw;
    T.pr_desc = PTR(uint32_t, i_desc); T.nz_hdr = PTR(uint4, i_hdr); T.nz_h = PTR(double, i_nzh); T.Hinv = PTR(double, i_hinv); T.Y0 = PTR(double, i_y0);
    {
        const int M = h.M, MP = M * kP;
        T.row_npl = h.np / h.D; T.row_bv = MP - 3; T.row_ba = T.row_bv + (M * 5 - 2); T.row_bc = T.row_ba + (M * 4 - 1);
        T.scv = h.scv; T.sca = h.sca;
    }
#undef PTR
    return 0;
}

int dlsc_create(const dlsc_params* hp, int n_agents, int agent_begin, int n_local, int device, dlsc_ctx** out) {
    if (!hp || !out) return fail("dlsc_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail("dlsc_create: no CUDA device available (libdlsc_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("dlsc_create: bad device index");
    if (hp->n != 5 || hp->phi != 3) return fail("dlsc_create: only traj/n = 5, traj/phi = 3 are supported");
    if (hp->dim != 2 && hp->dim != 3) return fail("dlsc_create: world/dimension must be 2 or 3");
    if (hp->M < 2 || hp->M > kMaxM) return fail("dlsc_create: traj/M out of range [2,16]");
    if (hp->dim * (3 * hp->M - 2) > 128) return fail("dlsc_create: dim*(3M-2) must be <= 128");
    if (hp->max_nbr < 1 || hp->max_nbr > 1024) return fail("dlsc_create: max_nbr must be in [1, 1024]");
    if (hp->qp_solver < 0 || hp->qp_solver > 3) return fail("dlsc_create: qp_solver must be 0..3");
    if (n_agents < 1 || agent_begin < 0 || n_local < 1 || agent_begin + n_local > n_agents)
        return fail("dlsc_create: bad agent block");
    if (!(hp->dt > 0) || !(hp->world_res > 0)) return fail("dlsc_create: dt and world_res must be positive");
    CK(cudaSetDevice(device));
    dlsc_ctx* c = new dlsc_ctx();
    c->hp = *hp;
    c->device = device;
    DevParams& P = c->P;
    memset(&P, 0, sizeof(P));
    P.M = hp->M; P.D = hp->dim; P.use_sfc = hp->use_sfc; P.K = hp->max_nbr;
    P.N = n_agents; P.begin = agent_begin; P.NL = n_local;
    c->rl = rec_layout(hp->M);
    P.rec = c->rl.size;
    P.qp_max_iter = hp->qp_max_iter > 0 ? hp->qp_max_iter : 80;
    P.qp_screen = hp->qp_screen_slack == 0.0 ? 0.5 : hp->qp_screen_slack;
    P.qp_solver = hp->qp_solver;
    P.qp_active_max = (hp->qp_active_max > 0 && hp->qp_active_max < 32) ? hp->qp_active_max : 32;
    P.dt = hp->dt; P.world_res = hp->world_res; P.grid_res = hp->grid_res; P.z_2d = hp->z_2d;
    P.comm_range = hp->comm_range; P.w_control = hp->w_control; P.w_terminal = hp->w_terminal;
    P.reset_threshold = hp->reset_threshold;
    for (int k = 0; k < 3; k++) {
        P.world_min[k] = (double)(float)hp->world_min[k];
        P.world_max[k] = (double)(float)hp->world_max[k];
    }
    {   // trajectory.cpp:84-90: time += dt/n after every control point
        double time = 0;
        for (int e = 0; e < hp->M * kP; e++) { P.tk[e] = (float)time; time += hp->dt / hp->n; }
    }
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    {
        const char* ov = getenv("DLSC_OVERLAP");
        c->overlap = !(ov && ov[0] == '0');
        const char* gv = getenv("DLSC_GRAPH");
        c->use_graph = !(gv && gv[0] == '0');
    }
    c->own_stream = true;
    memset(c->t_ms, 0, sizeof(c->t_ms));

    DevState& S = c->S;
    memset(&S, 0, sizeof(S));
    const size_t N = n_agents, NL = n_local, K = hp->max_nbr, M = hp->M, npt = M * kP;
    int rc = 0;
    rc |= dev_alloc(c, &S.rec, N * P.rec);
    c->rec_owned = true;
    rc |= dev_alloc(c, &S.acc, NL * 3);
    rc |= dev_alloc(c, &S.stage_pos, (size_t)NL * 3);
    rc |= dev_alloc(c, &S.stage_vel, (size_t)NL * 3);
    rc |= dev_alloc(c, &S.waypoint, NL * 3);
    rc |= dev_alloc(c, &S.goal_new, NL * 3);
    rc |= dev_alloc(c, &S.disturbed, NL);
    rc |= dev_alloc(c, &S.sfc_init, NL);
    rc |= dev_alloc(c, &S.radius, NL); rc |= dev_alloc(c, &S.downwash, NL); rc |= dev_alloc(c, &S.max_vel, NL);
    rc |= dev_alloc(c, &S.max_acc, NL); rc |= dev_alloc(c, &S.nominal_vel, NL);
    rc |= dev_alloc(c, &S.pred_traj, (N + kMaxDyn) * npt * 3);
    rc |= dev_alloc(c, &S.init_traj, NL * npt * 3);
    rc |= dev_alloc(c, &S.nbr_idx, NL * K);
    rc |= dev_alloc(c, &S.nbr_cnt, NL);
    rc |= dev_alloc(c, &S.lsc_normal, NL * K * M * 3);
    rc |= dev_alloc(c, &S.lsc_d, NL * K * M * kP);
    rc |= dev_alloc(c, &S.lsc_anchor_last, NL * K * 3);
    rc |= dev_alloc(c, &S.lsc_near, NL * K * M);
    rc |= dev_alloc(c, &S.lsc_queue, NL * K * M);
    rc |= dev_alloc(c, &S.sfc, NL * M * 6);
    rc |= dev_alloc(c, &S.traj, NL * npt * 3);
    rc |= dev_alloc(c, &S.qp_x, NL * (size_t)hp->dim * npt);
    rc |= dev_alloc(c, &S.cost, NL); rc |= dev_alloc(c, &S.viol, NL);
    rc |= dev_alloc(c, &S.qp_iters, NL); rc |= dev_alloc(c, &S.status, NL);
    rc |= dev_alloc(c, &S.counters, DLSC_N_COUNTERS);
    rc |= dev_alloc(c, &S.qp_next, 4);
    rc |= dev_alloc(c, &S.qp_list, NL);
    rc |= dev_alloc(c, &S.nbr_cell_start, (size_t)8192 + 1);
    rc |= dev_alloc(c, &S.nbr_sorted, (size_t)N);
    rc |= dev_alloc(c, &S.nbr_sorted_pos, (size_t)N);
    rc |= dev_alloc(c, &S.qp_list_gi, (size_t)NL);
    rc |= dev_alloc(c, &S.qp_seed, (size_t)NL * 4);
    rc |= dev_alloc(c, &S.dyn_pos, (size_t)kMaxDyn * 3); rc |= dev_alloc(c, &S.dyn_vel, (size_t)kMaxDyn * 3);
    rc |= dev_alloc(c, &S.dyn_radius, (size_t)kMaxDyn); rc |= dev_alloc(c, &S.dyn_downwash, (size_t)kMaxDyn);
    rc |= dev_alloc(c, &S.dyn_max_acc, (size_t)kMaxDyn);
    rc |= dev_alloc(c, &S.dyn_size, (size_t)kMaxDyn * M * kP);
    rc |= dev_alloc(c, &S.comm_box, NL * 6);
    rc |= dev_alloc(c, &S.trap, NL);
    if (rc) { dlsc_destroy(c); return -1; }

    build_qp_tables(hp->M, hp->dim, hp->dt, hp->w_control, hp->w_terminal, hp->comm_range > 0, c->th);
    if (upload_tables(c)) { dlsc_destroy(c); return -1; }
    c->qpl = qp_launch_config(P, c->T, device);
    if (c->qpl.smem > 227 * 1024) { dlsc_destroy(c); return fail("dlsc_create: QP shared memory exceeds 227 KB"); }
    if (dev_alloc(c, &S.qp_scratch, (size_t)c->qpl.ctas * c->qpl.scratch_doubles)) { dlsc_destroy(c); return -1; }
    {   // a 1-cell "free space" grid so the SFC kernel is well defined before dlsc_set_edt
        S.edt.dims[0] = S.edt.dims[1] = S.edt.dims[2] = 0;
        S.edt.res = hp->world_res; S.edt.inv_res = 1.0 / hp->world_res; S.edt.cells = nullptr;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); dlsc_destroy(c); return -1; }
    *out = c;
    return 0;
}

void dlsc_destroy(dlsc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->traj_host_registered) cudaHostUnregister(c->traj_host_registered);
    for (void* p : c->allocs) cudaFree(p);
    if (c->tab_blob) cudaFree(c->tab_blob);
    if (c->edt_cells) cudaFree(c->edt_cells);
    if (c->edt_centre) cudaFree(c->edt_centre);
    if (c->edt_mask) cudaFree(c->edt_mask);
    if (c->edt_sat) cudaFree(c->edt_sat);
    if (c->p2p.exported) {
        for (int r = 0; r < c->p2p.world; r++)
            if (c->p2p.on && r != c->p2p.rank && c->p2p.peer_block[r]) cudaIpcCloseMemHandle(c->p2p.peer_block[r]);
        cudaFree(c->p2p.block); cudaFree(c->p2p.done);
        if (c->p2p.err_host) cudaFreeHost((void*)c->p2p.err_host);
    }
    for (auto& g : c->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (c->side_stream) { cudaStreamSynchronize(c->side_stream); cudaStreamDestroy(c->side_stream); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (auto& e : c->evpool) if (e) cudaEventDestroy(e);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int dlsc_set_stream(dlsc_ctx* c, void* s) {
    if (!c) return fail("null ctx");
    if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = (cudaStream_t)s;
    c->own_stream = false;
    return 0;
}
void* dlsc_get_stream(dlsc_ctx* c) { return c ? (void*)c->stream : nullptr; }

// common tail of dlsc_set_edt / dlsc_build_edt*: publish the grid view and the per-axis cell-centre tables
static int finish_edt(dlsc_ctx* c, const int32_t dims[3], const int32_t min_key[3], double res) {
    EdtDev& E = c->S.edt;
    for (int k = 0; k < 3; k++) { E.dims[k] = dims[k]; E.min_key[k] = min_key[k]; }
    E.res = res; E.inv_res = 1.0 / res; E.cells = c->edt_cells;
    {   // cell-centre coordinate of every cell index along each axis (what DynamicEDTOctomap returns as the
        // closest obstacle: keyToCoord = (key + 0.5) * res), same double arithmetic as the per-vertex formula
        std::vector<float> centre((size_t)dims[0] + dims[1] + dims[2]);
        size_t o = 0;
        for (int k = 0; k < 3; k++)
            for (int i = 0; i < dims[k]; i++) centre[o++] = (float)(((double)(i + min_key[k]) + 0.5) * res);
        if (c->edt_centre) { cudaFree(c->edt_centre); c->edt_centre = nullptr; }
        CK(cudaMalloc(&c->edt_centre, centre.size() * sizeof(float)));
        CK(cudaMemcpy(c->edt_centre, centre.data(), centre.size() * sizeof(float), cudaMemcpyHostToDevice));
        E.centre[0] = c->edt_centre; E.centre[1] = c->edt_centre + dims[0]; E.centre[2] = c->edt_centre + dims[0] + dims[1];
    }
    c->have_edt = true;
    c->mask_dirty = true;
    CK(cudaGetLastError());
    return 0;
}

int dlsc_set_edt(dlsc_ctx* c, const float* dist, const int32_t* obst, const int32_t dims[3],
                 const int32_t min_key[3], double res) {
    if (!c || !dist || !obst) return fail("dlsc_set_edt: null argument");
    CK(cudaSetDevice(c->device));
    const size_t nc = (size_t)dims[0] * dims[1] * dims[2];
    if (nc == 0) return fail("dlsc_set_edt: empty grid");
    float* d_dist = nullptr; int32_t* d_obst = nullptr;
    CK(cudaMalloc(&d_dist, nc * sizeof(float)));
    CK(cudaMalloc(&d_obst, nc * 3 * sizeof(int32_t)));
    CK(cudaMemcpyAsync(d_dist, dist, nc * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_obst, obst, nc * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    if (c->edt_cells) { CK(cudaStreamSynchronize(c->stream)); cudaFree(c->edt_cells); c->edt_cells = nullptr; }
    CK(cudaMalloc(&c->edt_cells, nc * sizeof(int4)));
    launch_edt_pack(d_dist, d_obst, c->edt_cells, nc, c->stream);
    c->launches++;
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_dist); cudaFree(d_obst);
    return finish_edt(c, dims, min_key, res);
}

// grid extent of the mission world: DynamicEDTOctomap(maxdist, octree, world_min, world_max) covers the octomap keys
// floor(coord / res) of both corners (float coordinates), src/map_manager.cpp:76-78
int dlsc_edt_dims(const dlsc_ctx* c, int32_t dims[3], int32_t min_key[3]) {
    if (!c || !dims || !min_key) return fail("dlsc_edt_dims: null argument");
    const double inv = 1.0 / c->hp.world_res;
    for (int k = 0; k < 3; k++) {
        const int lo = (int)std::floor(inv * (double)(float)c->hp.world_min[k]);
        const int hi = (int)std::floor(inv * (double)(float)c->hp.world_max[k]);
        min_key[k] = lo; dims[k] = hi - lo + 1;
    }
    return 0;
}

// occupancy (device, 1 byte per cell) -> 16-byte records; frees nothing it did not allocate
static int edt_from_occupancy_dev(dlsc_ctx* c, const uint8_t* d_occ, const int32_t dims[3], const int32_t mk[3], double maxdist) {
    const size_t nc = (size_t)dims[0] * dims[1] * dims[2];
    const double res = c->hp.world_res;
    const int maxd = (int)(maxdist / res + 1);          // DynamicEDTOctomap: maxdist in cells
    uint32_t *ta = nullptr, *tb = nullptr;
    uint8_t* tcol = nullptr;
    CK(cudaMalloc(&ta, nc * sizeof(uint32_t)));
    CK(cudaMalloc(&tb, nc * sizeof(uint32_t)));
    CK(cudaMalloc(&tcol, 3 * (size_t)dims[0] * dims[1]));
    if (c->edt_cells) { CK(cudaStreamSynchronize(c->stream)); cudaFree(c->edt_cells); c->edt_cells = nullptr; }
    CK(cudaMalloc(&c->edt_cells, nc * sizeof(int4)));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, c->stream));
    const int nl = launch_edt_build(d_occ, ta, tb, tcol, c->edt_cells, dims, res, maxd, c->stream);
    CK(cudaEventRecord(e1, c->stream));
    if (nl < 0) { cudaFree(ta); cudaFree(tb); cudaFree(tcol); return fail("dlsc_build_edt: maxdist / resolution + 1 must be in [1, 16] cells"); }
    c->launches += nl;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    c->edt_build_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(ta); cudaFree(tb); cudaFree(tcol);
    return finish_edt(c, dims, mk, res);
}

int dlsc_build_edt(dlsc_ctx* c, const float* boxes, int n_boxes, double maxdist) {
    if (!c || (n_boxes > 0 && !boxes) || n_boxes < 0) return fail("dlsc_build_edt: bad argument");
    CK(cudaSetDevice(c->device));
    int32_t dims[3], mk[3];
    dlsc_edt_dims(c, dims, mk);
    const size_t nc = (size_t)dims[0] * dims[1] * dims[2];
    if (nc == 0 || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail("dlsc_build_edt: empty world");
    uint8_t* d_occ = nullptr; float* d_boxes = nullptr;
    CK(cudaMalloc(&d_occ, nc));
    CK(cudaMemsetAsync(d_occ, 0, nc, c->stream));
    if (n_boxes > 0) {
        CK(cudaMalloc(&d_boxes, (size_t)n_boxes * 6 * sizeof(float)));
        CK(cudaMemcpyAsync(d_boxes, boxes, (size_t)n_boxes * 6 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        launch_edt_raster(d_boxes, n_boxes, c->hp.world_res, dims, mk, d_occ, c->stream);
        c->launches++;
    }
    const int rc = edt_from_occupancy_dev(c, d_occ, dims, mk, maxdist);
    cudaFree(d_occ);
    if (d_boxes) cudaFree(d_boxes);
    return rc;
}

int dlsc_build_edt_occupancy(dlsc_ctx* c, const uint8_t* occ, double maxdist) {
    if (!c || !occ) return fail("dlsc_build_edt_occupancy: null argument");
    CK(cudaSetDevice(c->device));
    int32_t dims[3], mk[3];
    dlsc_edt_dims(c, dims, mk);
    const size_t nc = (size_t)dims[0] * dims[1] * dims[2];
    if (nc == 0 || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail("dlsc_build_edt_occupancy: empty world");
    uint8_t* d_occ = nullptr;
    CK(cudaMalloc(&d_occ, nc));
    CK(cudaMemcpyAsync(d_occ, occ, nc, cudaMemcpyHostToDevice, c->stream));
    const int rc = edt_from_occupancy_dev(c, d_occ, dims, mk, maxdist);
    cudaFree(d_occ);
    return rc;
}

int dlsc_get_edt(dlsc_ctx* c, float* dist, int32_t* obst) {
    if (!c || !dist || !obst) return fail("dlsc_get_edt: null argument");
    if (!c->have_edt || !c->edt_cells) return fail("dlsc_get_edt: no grid");
    CK(cudaSetDevice(c->device));
    const EdtDev& E = c->S.edt;
    const size_t nc = (size_t)E.dims[0] * E.dims[1] * E.dims[2];
    float* d_dist = nullptr; int32_t* d_obst = nullptr;
    CK(cudaMalloc(&d_dist, nc * sizeof(float)));
    CK(cudaMalloc(&d_obst, nc * 3 * sizeof(int32_t)));
    launch_edt_unpack(c->edt_cells, d_dist, d_obst, nc, c->stream);
    c->launches++;
    CK(cudaMemcpyAsync(dist, d_dist, nc * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(obst, d_obst, nc * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_dist); cudaFree(d_obst);
    return 0;
}

double dlsc_edt_build_ms(const dlsc_ctx* c) { return c ? c->edt_build_ms : -1.0; }

// (Re)build the lattice-vertex mask for the current grid and margin.  The mask is dropped (record path only)
// when a tabulated decision is not robust (edt_vertex_mask) or the grid is too tall for the z tables.
static int build_vertex_mask(dlsc_ctx* c) {
    EdtDev& E = c->S.edt;
    c->mask_dirty = false;
    if (c->edt_mask) { CK(cudaStreamSynchronize(c->stream)); cudaFree(c->edt_mask); c->edt_mask = nullptr; }
    if (c->edt_sat) { cudaFree(c->edt_sat); c->edt_sat = nullptr; }
    E.sat = nullptr;
    E.vmask = nullptr; E.zs = edt_mask_zs(E.dims[2]); E.mask_margin = c->margin_host;
    const char* env = getenv("DLSC_SFC_MASK");
    if ((env && env[0] == '0') || E.zs > kSfcZsMax || !E.cells) return 0;
    const size_t n = (size_t)(E.dims[0] + 1) * (E.dims[1] + 1) * E.zs;
    int* d_unsafe = nullptr;
    CK(cudaMalloc(&c->edt_mask, n));
    CK(cudaMalloc(&d_unsafe, sizeof(int)));
    CK(cudaMemsetAsync(d_unsafe, 0, sizeof(int), c->stream));
    launch_edt_mask(E, c->margin_host, c->edt_mask, d_unsafe, c->stream);
    c->launches++;
    int unsafe = 1;
    CK(cudaMemcpyAsync(&unsafe, d_unsafe, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_unsafe);
    if (unsafe) { cudaFree(c->edt_mask); c->edt_mask = nullptr; }
    E.vmask = c->edt_mask;
    const char* env_sat = getenv("DLSC_SFC_SAT");
    if (E.vmask && !(env_sat && env_sat[0] == '0')) {
        const size_t ns = (size_t)(E.dims[0] + 2) * (E.dims[1] + 2) * (E.dims[2] + 2);
        if (ns < (size_t)1 << 31) {
            CK(cudaMalloc(&c->edt_sat, ns * sizeof(int32_t)));
            launch_sat_build(E, c->edt_sat, c->stream);
            c->launches += 4;
            CK(cudaStreamSynchronize(c->stream));
            E.sat = c->edt_sat;
        }
    }
    CK(cudaGetLastError());
    return 0;
}

int dlsc_set_agent_props(dlsc_ctx* c, const dlsc_agent_props* p) {
    if (!c || !p) return fail("dlsc_set_agent_props: null argument");
    CK(cudaSetDevice(c->device));
    const size_t b = (size_t)c->P.NL * sizeof(double);
    if (p->radius) {
        CK(cudaMemcpyAsync(c->S.radius, p->radius, b, cudaMemcpyHostToDevice, c->stream));
        if (p->radius[0] != c->margin_host) { c->margin_host = p->radius[0]; c->mask_dirty = true; }
    }
    if (p->downwash) CK(cudaMemcpyAsync(c->S.downwash, p->downwash, b, cudaMemcpyHostToDevice, c->stream));
    if (p->max_vel) CK(cudaMemcpyAsync(c->S.max_vel, p->max_vel, b, cudaMemcpyHostToDevice, c->stream));
    if (p->max_acc) CK(cudaMemcpyAsync(c->S.max_acc, p->max_acc, b, cudaMemcpyHostToDevice, c->stream));
    if (p->nominal_vel) CK(cudaMemcpyAsync(c->S.nominal_vel, p->nominal_vel, b, cudaMemcpyHostToDevice, c->stream));
    if (p->radius || p->downwash) {
        // the other agents read radius / downwash out of the exchanged records (through float, agent_manager.cpp:256-258):
        // keep the local records in step with the properties whatever the call order relative to dlsc_reset
        std::vector<float> f((size_t)c->P.NL);
        float* dst = c->S.rec + (size_t)c->P.begin * c->P.rec;
        if (p->radius) {
            for (int i = 0; i < c->P.NL; i++) f[i] = (float)p->radius[i];
            CK(cudaMemcpy2DAsync(dst + c->rl.radius, (size_t)c->P.rec * sizeof(float), f.data(), 4, 4, c->P.NL, cudaMemcpyHostToDevice, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
        if (p->downwash) {
            for (int i = 0; i < c->P.NL; i++) f[i] = (float)p->downwash[i];
            CK(cudaMemcpy2DAsync(dst + c->rl.downwash, (size_t)c->P.rec * sizeof(float), f.data(), 4, 4, c->P.NL, cudaMemcpyHostToDevice, c->stream));
        }
    }
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int dlsc_set_obstacles(dlsc_ctx* c, const dlsc_obstacles* o, const dlsc_obstacle_params* op) {
    if (!c) return fail("null ctx");
    CK(cudaSetDevice(c->device));
    const int n = o ? o->n : 0;
    if (n < 0 || n > kMaxDyn) return fail("dlsc_set_obstacles: at most 16 dynamic obstacles");
    if (n == 0) { c->P.n_dyn = 0; return 0; }
    if (!o->pos || !o->vel || !o->radius || !o->downwash || !o->max_acc || !op) return fail("dlsc_set_obstacles: null member");
    if (!(op->slack_collision_weight > 0)) return fail("dlsc_set_obstacles: slack_collision_weight must be positive");
    if (n >= c->P.K) return fail("dlsc_set_obstacles: max_nbr must exceed the number of dynamic obstacles");
    if (!c->S.qp_slack && dev_alloc(c, &c->S.qp_slack, (size_t)c->P.NL * kMaxDyn * c->P.M)) return -1;
    std::vector<double> size((size_t)n * c->P.M * kP);
    for (int i = 0; i < n; i++) {
        if (!(o->radius[i] > 0)) return fail("dlsc_set_obstacles: radius must be positive");
        dyn_obstacle_sizes(c->P, op->size_prediction != 0, op->uncertainty_horizon, o->radius[i], o->max_acc[i], size.data() + (size_t)i * c->P.M * kP);
    }
    CK(cudaMemcpyAsync(c->S.dyn_pos, o->pos, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->S.dyn_vel, o->vel, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->S.dyn_radius, o->radius, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->S.dyn_downwash, o->downwash, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->S.dyn_max_acc, o->max_acc, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->S.dyn_size, size.data(), size.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));       // the host arrays may go away after the call
    c->P.n_dyn = n;
    c->P.slack_w = op->slack_collision_weight;
    c->P.dyn_horizon = op->uncertainty_horizon;
    return 0;
}
int dlsc_get_slack(dlsc_ctx* c, double* slack) {
    if (!c || !slack) return fail("dlsc_get_slack: null argument");
    const int nd = c->P.n_dyn, M = c->P.M;
    if (nd == 0) return 0;
    std::vector<double> tmp((size_t)c->P.NL * kMaxDyn * M);
    if (d2h_raw(c, tmp.data(), c->S.qp_slack, tmp.size() * 8)) return -1;
    for (int a = 0; a < c->P.NL; a++)
        memcpy(slack + (size_t)a * nd * M, tmp.data() + (size_t)a * kMaxDyn * M, (size_t)nd * M * 8);
    return 0;
}
int dlsc_get_trap(dlsc_ctx* c, uint8_t* trap) {
    if (!c || !trap) return fail("dlsc_get_trap: null argument");
    return d2h_raw(c, trap, c->S.trap, (size_t)c->P.NL);
}

int dlsc_reset(dlsc_ctx* c, const float* start) {
    if (!c || !start) return fail("dlsc_reset: null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->S.init_traj, start, (size_t)c->P.NL * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    launch_reset(c->P, c->S, c->S.init_traj, c->stream);
    c->launches++;
    c->seq = 0;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    return 0;
}

int dlsc_set_groups(dlsc_ctx* c, const int32_t* group) {
    if (!c || !group) return fail("dlsc_set_groups: null argument");
    CK(cudaSetDevice(c->device));
    const DevParams& P = c->P;
    std::vector<float> g((size_t)P.NL);
    for (int i = 0; i < P.NL; i++) {
        if (group[i] < 0 || group[i] >= (1 << 24)) return fail("dlsc_set_groups: group index out of range [0, 2^24)");
        g[i] = (float)group[i];
    }
    float* dst = c->S.rec + (size_t)P.begin * P.rec + c->rl.group;
    CK(cudaMemcpy2DAsync(dst, (size_t)P.rec * sizeof(float), g.data(), 4, 4, P.NL, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int dlsc_set_agents(dlsc_ctx* c, const dlsc_agents* a) {
    if (!c || !a) return fail("dlsc_set_agents: null argument");
    CK(cudaSetDevice(c->device));
    const DevParams& P = c->P;
    if (a->pos) CK(cudaMemcpyAsync(c->S.stage_pos, a->pos, (size_t)P.NL * 12, cudaMemcpyHostToDevice, c->stream));
    if (a->vel) CK(cudaMemcpyAsync(c->S.stage_vel, a->vel, (size_t)P.NL * 12, cudaMemcpyHostToDevice, c->stream));
    if (a->pos || a->vel) {
        launch_set_state(P, c->S.rec, a->pos ? c->S.stage_pos : nullptr, a->vel ? c->S.stage_vel : nullptr, c->stream);
        c->launches++;
    }
    if (a->acc) CK(cudaMemcpyAsync(c->S.acc, a->acc, (size_t)P.NL * 12, cudaMemcpyHostToDevice, c->stream));
    if (a->waypoint) CK(cudaMemcpyAsync(c->S.waypoint, a->waypoint, (size_t)P.NL * 12, cudaMemcpyHostToDevice, c->stream));
    if (a->disturbed) CK(cudaMemcpyAsync(c->S.disturbed, a->disturbed, (size_t)P.NL, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int dlsc_bind_traj_host(dlsc_ctx* c, float* host) {
    if (!c) return fail("null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (c->traj_host_registered) { cudaHostUnregister(c->traj_host_registered); c->traj_host_registered = nullptr; }
    c->S.traj_host = nullptr;
    if (!host) return 0;
    void* dev = nullptr;
    if (cudaHostGetDevicePointer(&dev, host, 0) != cudaSuccess) {       // not pinned yet: pin and map it here
        cudaGetLastError();
        const size_t bytes = (size_t)c->P.NL * c->P.M * kP * 12;
        if (cudaHostRegister(host, bytes, cudaHostRegisterMapped) != cudaSuccess) {
            cudaGetLastError();
            return fail("dlsc_bind_traj_host: the buffer is neither pinned nor registrable");
        }
        c->traj_host_registered = host;
        CK(cudaHostGetDevicePointer(&dev, host, 0));
    }
    c->S.traj_host = static_cast<float*>(dev);
    return 0;
}

float* dlsc_records_device(dlsc_ctx* c) { return c ? c->S.rec : nullptr; }
int dlsc_record_floats(const dlsc_ctx* c) { return c ? c->P.rec : 0; }
int dlsc_bind_records(dlsc_ctx* c, float* p) {
    if (!c || !p) return fail("dlsc_bind_records: null argument");
    if (c->p2p.on) return fail("dlsc_bind_records: records live in the connected peer-memory block (dlsc_p2p_connect)");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(p, c->S.rec, (size_t)c->P.N * c->P.rec * sizeof(float), cudaMemcpyDeviceToDevice));
    c->S.rec = p;
    c->rec_owned = false;
    return 0;
}
// ---- record exchange over NVLink peer memory ------------------------------------------------------------------
static size_t p2p_buf_floats(const dlsc_ctx* c) { return ((size_t)c->P.N * c->P.rec + 63) / 64 * 64; }
static float* p2p_buffer(const dlsc_ctx* c, void* block, int which) {
    return reinterpret_cast<float*>(static_cast<char*>(block) + 256) + (size_t)which * p2p_buf_floats(c);
}
int dlsc_p2p_export(dlsc_ctx* c, void* handle64) {
    if (!c || !handle64) return fail("dlsc_p2p_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaSetDevice(c->device));
    if (!c->p2p.exported) {
        if (!c->rec_owned) return fail("dlsc_p2p_export: records are bound to caller memory (dlsc_bind_records)");
        if (c->P.rec % 4 != 0) return fail("dlsc_p2p_export: record size must be a multiple of 4 floats");
        const size_t bytes = 256 + 2 * p2p_buf_floats(c) * sizeof(float);
        CK(cudaMalloc(&c->p2p.block, bytes));
        CK(cudaMemset(c->p2p.block, 0, bytes));
        CK(cudaMalloc(&c->p2p.done, sizeof(unsigned)));
        CK(cudaMemset(c->p2p.done, 0, sizeof(unsigned)));
        {   // error flag in mapped pinned memory: the wait kernel raises it, the host reads it without a synchronisation
            int* h = nullptr;
            CK(cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped));
            *h = 0;
            c->p2p.err_host = h;
            CK(cudaHostGetDevicePointer((void**)&c->p2p.err, h, 0));
        }
        CK(cudaStreamSynchronize(c->stream));
        const size_t nb = (size_t)c->P.N * c->P.rec * sizeof(float);
        CK(cudaMemcpy(p2p_buffer(c, c->p2p.block, 0), c->S.rec, nb, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(p2p_buffer(c, c->p2p.block, 1), c->S.rec, nb, cudaMemcpyDeviceToDevice));
        c->S.rec = p2p_buffer(c, c->p2p.block, 0);      // the old array stays allocated until dlsc_destroy
        c->p2p.cur = 0;
        c->p2p.exported = true;
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->p2p.block));
    memcpy(handle64, &h, 64);
    return 0;
}
int dlsc_p2p_connect(dlsc_ctx* c, int world, int rank, const void* handles) {
    if (!c || !handles) return fail("dlsc_p2p_connect: null argument");
    if (!c->p2p.exported) return fail("dlsc_p2p_connect: call dlsc_p2p_export first");
    if (world < 1 || world > kP2PMaxWorld || rank < 0 || rank >= world) return fail("dlsc_p2p_connect: bad world / rank");
    CK(cudaSetDevice(c->device));
    for (int r = 0; r < world; r++) {
        if (r == rank) { c->p2p.peer_block[r] = c->p2p.block; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t)r * 64, 64);
        const cudaError_t e = cudaIpcOpenMemHandle(&c->p2p.peer_block[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {                      // unmap what was opened so far: nothing is left half connected
            for (int q = 0; q < r; q++)
                if (q != rank && c->p2p.peer_block[q]) { cudaIpcCloseMemHandle(c->p2p.peer_block[q]); c->p2p.peer_block[q] = nullptr; }
            c->p2p.peer_block[r] = nullptr;
            g_err = std::string("dlsc_p2p_connect: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
            return -1;
        }
    }
    c->p2p.world = world; c->p2p.rank = rank; c->p2p.on = true;
    return 0;
}
// a peer that missed an exchange makes every later step of this context fail (the records it would plan against are stale)
static int p2p_check(dlsc_ctx* c) {
    if (!c->p2p.on || !c->p2p.err_host) return 0;
    if (*c->p2p.err_host) { c->p2p.failed = true; }
    if (c->p2p.failed) return fail("record exchange failed: a peer did not publish its records within the timeout (dlsc_p2p_status)");
    return 0;
}
int dlsc_exchange_records(dlsc_ctx* c) {
    if (!c) return fail("null ctx");
    if (!c->p2p.on) return fail("dlsc_exchange_records: not connected (dlsc_p2p_connect)");
    CK(cudaSetDevice(c->device));
    if (p2p_check(c)) return -1;                   // a timed-out exchange is never followed by a buffer switch
    auto& X = c->p2p;
    const int next = X.cur ^ 1;
    float* dst[kP2PMaxWorld]; unsigned long long* flag[kP2PMaxWorld];
    for (int r = 0; r < X.world; r++) {
        dst[r] = p2p_buffer(c, X.peer_block[r], next);
        flag[r] = reinterpret_cast<unsigned long long*>(X.peer_block[r]);
    }
    X.step++;
    const size_t off = (size_t)c->P.begin * c->P.rec, n = (size_t)c->P.NL * c->P.rec;
    launch_p2p_push(c->S.rec + off, n, off, X.world, X.rank, X.step, dst, flag, X.done, c->stream);
    launch_p2p_wait(reinterpret_cast<unsigned long long*>(X.block), X.world, X.step, X.err, c->stream);
    c->launches += 2;
    X.cur = next;
    c->S.rec = p2p_buffer(c, X.block, next);
    CK(cudaGetLastError());
    return 0;
}
// unmap the peers' blocks (call on every rank, then a barrier, before any rank destroys its context)
int dlsc_p2p_disconnect(dlsc_ctx* c) {
    if (!c) return fail("null ctx");
    if (!c->p2p.on) return 0;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < c->p2p.world; r++)
        if (r != c->p2p.rank && c->p2p.peer_block[r]) { cudaIpcCloseMemHandle(c->p2p.peer_block[r]); c->p2p.peer_block[r] = nullptr; }
    c->p2p.on = false;
    return 0;
}
int dlsc_p2p_status(dlsc_ctx* c) {
    if (!c) return fail("null ctx");
    if (!c->p2p.exported) return 0;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (c->p2p.err_host && *c->p2p.err_host) c->p2p.failed = true;
    if (c->p2p.failed) return fail("dlsc_p2p_status: a peer did not publish its records within the timeout");
    return 0;
}

int dlsc_set_records(dlsc_ctx* c, int first, int count, const float* host) {
    if (!c || !host || first < 0 || count < 0 || first + count > c->P.N) return fail("dlsc_set_records: bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->S.rec + (size_t)first * c->P.rec, host, (size_t)count * c->P.rec * sizeof(float),
                       cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
int dlsc_get_records(dlsc_ctx* c, int first, int count, float* host) {
    if (!c || !host || first < 0 || count < 0 || first + count > c->P.N) return fail("dlsc_get_records: bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(host, c->S.rec + (size_t)first * c->P.rec, (size_t)count * c->P.rec * sizeof(float),
                       cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

static int run_stages_impl(dlsc_ctx* c, int mask, const DevParams& P, const DevState& S, int seq_override = -1);

int dlsc_run_stages(dlsc_ctx* c, int mask) {
    if (!c) return fail("null ctx");
    return run_stages_impl(c, mask, c->P, c->S);
}

// view of the local agents [first, first + count): every per-agent array offset, P.begin / P.NL adjusted
static void make_view(const dlsc_ctx* c, int first, int count, DevParams& P, DevState& S) {
    P = c->P;
    S = c->S;
    const size_t f = (size_t)first, K = P.K, M = P.M, npt = (size_t)P.M * kP;
    P.begin += first; P.NL = count;
    S.acc += f * 3; S.waypoint += f * 3; S.goal_new += f * 3; S.disturbed += f; S.sfc_init += f;
    S.radius += f; S.downwash += f; S.max_vel += f; S.max_acc += f; S.nominal_vel += f;
    S.init_traj += f * npt * 3; S.nbr_idx += f * K; S.nbr_cnt += f;
    S.lsc_normal += f * K * M * 3; S.lsc_d += f * K * M * kP; S.lsc_anchor_last += f * K * 3; S.lsc_near += f * K * M;
    S.sfc += f * M * 6; S.traj += f * npt * 3; S.qp_x += f * (size_t)P.D * npt;
    if (S.traj_host) S.traj_host += f * npt * 3;
    S.cost += f; S.viol += f; S.qp_iters += f; S.status += f;
    S.qp_list += f; S.qp_list_gi += f; S.qp_seed += f * 4;
    S.comm_box += f * 6; S.trap += f;
    if (S.qp_slack) S.qp_slack += f * kMaxDyn * M;
}

int dlsc_run_stages_subset(dlsc_ctx* c, int mask, int first, int count) {
    if (!c) return fail("null ctx");
    if (first < 0 || count < 1 || first + count > c->P.NL) return fail("dlsc_run_stages_subset: bad range");
    DevParams P;
    DevState S;
    make_view(c, first, count, P, S);
    return run_stages_impl(c, mask, P, S);
}

static int run_stages_impl(dlsc_ctx* c, int mask, const DevParams& Pr, const DevState& Sr, int seq_override) {
    CK(cudaSetDevice(c->device));
    if (p2p_check(c)) return -1;
    if ((mask & DLSC_STAGE_SFC) && c->P.use_sfc && !c->have_edt) return fail("dlsc_run_stages: use_sfc set but no EDT grid (dlsc_set_edt)");
    if ((mask & DLSC_STAGE_SFC) && c->P.use_sfc && c->mask_dirty && build_vertex_mask(c)) return -1;
    DevState Sfix = Sr;
    Sfix.edt = c->S.edt;            // the grid may have been (re)set after a subset view was built
    const DevState& Sx = Sfix;
    cudaStream_t st = c->stream;
    // planner_seq after the increment in TrajPlanner::plan (traj_planner.cpp:40); the kernels only ask "seq < 2?"
    const int seq = seq_override >= 0 ? seq_override : c->seq + 1;
    const bool tm = c->timing;
    cudaEvent_t* ev = nullptr;
    if (tm) {
        const size_t need = (size_t)(c->ev_used + 1) * (DLSC_N_STAGES + 1);
        while (c->evpool.size() < need) { cudaEvent_t e; CK(cudaEventCreate(&e)); c->evpool.push_back(e); }
        ev = c->evpool.data() + (size_t)c->ev_used * (DLSC_N_STAGES + 1);
    }
    if (mask & (DLSC_STAGE_NBR | DLSC_STAGE_LSC | DLSC_STAGE_SFC | DLSC_STAGE_QP))
        CK(cudaMemsetAsync(c->S.counters, 0, DLSC_N_COUNTERS * sizeof(unsigned long long), st));
    if (tm) CK(cudaEventRecord(ev[0], st));
    if (mask & DLSC_STAGE_PREDICT) {
        launch_predict(Pr, Sx, seq, st); c->launches++;
        if (Pr.n_dyn > 0) { launch_dyn_predict(Pr, Sx, st); c->launches++; }
    }
    if (tm) CK(cudaEventRecord(ev[1], st));
    const bool sfc_on = (mask & DLSC_STAGE_SFC) && c->P.use_sfc;
    const bool fork = sfc_on && (mask & DLSC_STAGE_LSC) && !tm && c->overlap && c->side_stream;
    if (fork) {   // k_sfc is latency bound, k_lsc FP64 bound: they share the SMs well
        CK(cudaEventRecord(c->ev_fork, st));
        CK(cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
        launch_sfc(Pr, Sx, c->side_stream); c->launches++;
        CK(cudaEventRecord(c->ev_join, c->side_stream));
    }
    if (mask & DLSC_STAGE_NBR) c->launches += launch_neighbours(Pr, Sx, st);
    if (tm) CK(cudaEventRecord(ev[2], st));
    if (mask & DLSC_STAGE_LSC) c->launches += launch_lsc(Pr, Sx, st);
    if (tm) CK(cudaEventRecord(ev[3], st));
    if (fork) CK(cudaStreamWaitEvent(st, c->ev_join, 0));
    else if (sfc_on) { launch_sfc(Pr, Sx, st); c->launches++; }
    if (tm) CK(cudaEventRecord(ev[4], st));
    if (Pr.n_dyn > 0 && (mask & DLSC_STAGE_GOAL)) { launch_trap(Pr, Sx, st); c->launches++; }   // checkWaypointTrap sits between SFC and goal
    if (mask & DLSC_STAGE_GOAL) { launch_goal(Pr, Sx, st); c->launches++; }
    else if (mask & DLSC_STAGE_QP) { launch_goal_copy(Pr, Sx, st); c->launches++; }   // QP alone: goal from the record
    if (tm) CK(cudaEventRecord(ev[5], st));
    if (mask & DLSC_STAGE_QP) c->launches += launch_qp(Pr, Sx, c->T, c->qpl, st);
    if (tm) { CK(cudaEventRecord(ev[6], st)); c->ev_used++; }
    CK(cudaGetLastError());
    return 0;
}

// One whole step as a CUDA graph.  Usable from the second replan on (the first one takes the constant-velocity branch of
// the prediction, a different kernel argument), outside the per-stage timing mode, once the distance grid is in place.
static int step_graph(dlsc_ctx* c, bool& done) {
    done = false;
    if (!c->use_graph || c->timing || c->seq + 1 < 2) return 0;
    if (c->P.use_sfc && (!c->have_edt || c->mask_dirty)) return 0;
    CK(cudaSetDevice(c->device));
    if (p2p_check(c)) return -1;
    dlsc_ctx::StepGraph* g = nullptr;
    for (auto& e : c->graphs)
        if (e.exec && memcmp(&e.P, &c->P, sizeof(DevParams)) == 0 && memcmp(&e.S, &c->S, sizeof(DevState)) == 0) g = &e;
    if (!g) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(c->stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return 0;   // caller is capturing
        g = &c->graphs[c->graph_next];
        c->graph_next ^= 1;
        if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
        const int64_t l0 = c->launches;
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return 0; }
        const int rc = run_stages_impl(c, DLSC_STAGE_ALL, c->P, c->S, 2);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        g->kernels = (int)(c->launches - l0);
        c->launches = l0;
        if (rc != 0 || e != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); c->use_graph = false; return 0; }
        const cudaError_t ei = cudaGraphInstantiate(&g->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) { g->exec = nullptr; cudaGetLastError(); c->use_graph = false; return 0; }
        g->P = c->P; g->S = c->S;
    }
    CK(cudaGraphLaunch(g->exec, c->stream));
    c->launches += g->kernels;
    done = true;
    return 0;
}

int dlsc_step(dlsc_ctx* c) {
    if (!c) return fail("null ctx");
    bool done = false;
    if (step_graph(c, done)) return -1;
    if (done) { c->seq++; return 0; }
    const int rc = dlsc_run_stages(c, DLSC_STAGE_ALL);
    if (rc == 0) c->seq++;
    return rc;
}

int dlsc_advance(dlsc_ctx* c) {
    if (!c) return fail("null ctx");
    CK(cudaSetDevice(c->device));
    launch_advance(c->P, c->S, true, c->stream);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}
int dlsc_publish_records(dlsc_ctx* c) {
    if (!c) return fail("null ctx");
    CK(cudaSetDevice(c->device));
    launch_advance(c->P, c->S, false, c->stream);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}
int dlsc_sync(dlsc_ctx* c) {
    if (!c) return fail("null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
int dlsc_get_seq(const dlsc_ctx* c) { return c ? c->seq : -1; }
int dlsc_set_seq(dlsc_ctx* c, int seq) { if (!c) return fail("null ctx"); c->seq = seq; return 0; }

static int d2h(dlsc_ctx* c, void* host, const void* dev, size_t bytes);
static int d2h_raw(dlsc_ctx* c, void* host, const void* dev, size_t bytes) { return d2h(c, host, dev, bytes); }
static int d2h(dlsc_ctx* c, void* host, const void* dev, size_t bytes) {
    if (!c || !host) return fail("null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int dlsc_get_traj(dlsc_ctx* c, float* t) { return c ? d2h(c, t, c->S.traj, (size_t)c->P.NL * c->P.M * kP * 12) : fail("null ctx"); }
int dlsc_get_qp_x(dlsc_ctx* c, double* x) { return c ? d2h(c, x, c->S.qp_x, (size_t)c->P.NL * c->T.nx * 8) : fail("null ctx"); }
int dlsc_get_cost(dlsc_ctx* c, double* v) { return c ? d2h(c, v, c->S.cost, (size_t)c->P.NL * 8) : fail("null ctx"); }
int dlsc_get_violation(dlsc_ctx* c, double* v) { return c ? d2h(c, v, c->S.viol, (size_t)c->P.NL * 8) : fail("null ctx"); }
int dlsc_get_qp_iters(dlsc_ctx* c, int32_t* v) { return c ? d2h(c, v, c->S.qp_iters, (size_t)c->P.NL * 4) : fail("null ctx"); }
int dlsc_get_status(dlsc_ctx* c, int32_t* v) { return c ? d2h(c, v, c->S.status, (size_t)c->P.NL * 4) : fail("null ctx"); }
int dlsc_get_init_traj(dlsc_ctx* c, float* t) { return c ? d2h(c, t, c->S.init_traj, (size_t)c->P.NL * c->P.M * kP * 12) : fail("null ctx"); }
int dlsc_get_pred_traj(dlsc_ctx* c, float* t) { return c ? d2h(c, t, c->S.pred_traj, (size_t)c->P.N * c->P.M * kP * 12) : fail("null ctx"); }
int dlsc_get_obstacle_pred(dlsc_ctx* c, float* t) {
    return c ? d2h(c, t, c->S.pred_traj + (size_t)c->P.N * c->P.M * kP * 3, (size_t)c->P.n_dyn * c->P.M * kP * 12) : fail("null ctx");
}
int dlsc_get_sfc(dlsc_ctx* c, float* s) { return c ? d2h(c, s, c->S.sfc, (size_t)c->P.NL * c->P.M * 24) : fail("null ctx"); }

int dlsc_get_goal(dlsc_ctx* c, float* goal) { return c ? d2h(c, goal, c->S.goal_new, (size_t)c->P.NL * 12) : fail("null ctx"); }
int dlsc_get_state(dlsc_ctx* c, float* pos, float* vel, float* acc) {
    if (!c) return fail("null ctx");
    CK(cudaSetDevice(c->device));
    const DevParams& P = c->P;
    if (pos || vel) {      // gather out of the 768-byte records on the device, then dense copies
        launch_get_state(P, c->S.rec, c->S.stage_pos, c->S.stage_vel, c->stream);
        c->launches++;
    }
    if (pos) CK(cudaMemcpyAsync(pos, c->S.stage_pos, (size_t)P.NL * 12, cudaMemcpyDeviceToHost, c->stream));
    if (vel) CK(cudaMemcpyAsync(vel, c->S.stage_vel, (size_t)P.NL * 12, cudaMemcpyDeviceToHost, c->stream));
    if (acc) CK(cudaMemcpyAsync(acc, c->S.acc, (size_t)P.NL * 12, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
int dlsc_get_neighbours(dlsc_ctx* c, int32_t* idx, int32_t* cnt) {
    if (!c) return fail("null ctx");
    if (idx && d2h(c, idx, c->S.nbr_idx, (size_t)c->P.NL * c->P.K * 4)) return -1;
    if (cnt && d2h(c, cnt, c->S.nbr_cnt, (size_t)c->P.NL * 4)) return -1;
    return 0;
}
int dlsc_get_lsc(dlsc_ctx* c, float* normal, float* anchor, double* d) {
    if (!c) return fail("null ctx");
    const DevParams& P = c->P;
    const size_t pairs = (size_t)P.NL * P.K;
    if (normal && d2h(c, normal, c->S.lsc_normal, pairs * P.M * 12)) return -1;
    if (d && d2h(c, d, c->S.lsc_d, pairs * P.M * kP * 8)) return -1;
    if (anchor) {
        float* tmp = nullptr;
        const size_t bytes = pairs * P.M * kP * 12;
        CK(cudaMalloc(&tmp, bytes));
        launch_expand_anchor(P, c->S, tmp, c->stream);
        c->launches++;
        const int rc = d2h(c, anchor, tmp, bytes);
        cudaFree(tmp);
        if (rc) return -1;
    }
    return 0;
}
static int h2d(dlsc_ctx* c, void* dev, const void* host, size_t bytes) {
    if (!c || !host) return fail("null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
// injected stage inputs invalidate the QP row screen that k_lsc derived from the previous ones: every item "near"
static int screen_off(dlsc_ctx* c) {
    CK(cudaMemsetAsync(c->S.lsc_near, 0, (size_t)c->P.NL * c->P.K * c->P.M * sizeof(float), c->stream));   // slack 0: never screened
    return 0;
}
int dlsc_set_init_traj(dlsc_ctx* c, const float* t) {
    if (!c) return fail("null ctx");
    if (screen_off(c)) return -1;
    return h2d(c, c->S.init_traj, t, (size_t)c->P.NL * c->P.M * kP * 12);
}
int dlsc_set_pred_traj(dlsc_ctx* c, const float* t) {
    if (!c) return fail("null ctx");
    if (screen_off(c)) return -1;
    return h2d(c, c->S.pred_traj, t, (size_t)c->P.N * c->P.M * kP * 12);
}
int dlsc_set_neighbours(dlsc_ctx* c, const int32_t* idx, const int32_t* cnt) {
    if (!c) return fail("null ctx");
    if (screen_off(c)) return -1;
    if (h2d(c, c->S.nbr_idx, idx, (size_t)c->P.NL * c->P.K * 4)) return -1;
    return h2d(c, c->S.nbr_cnt, cnt, (size_t)c->P.NL * 4);
}
int dlsc_set_lsc(dlsc_ctx* c, const float* normal, const float* anchor_last, const double* d) {
    if (!c) return fail("null ctx");
    const size_t pairs = (size_t)c->P.NL * c->P.K;
    if (h2d(c, c->S.lsc_normal, normal, pairs * c->P.M * 12)) return -1;
    if (h2d(c, c->S.lsc_anchor_last, anchor_last, pairs * 12)) return -1;
    if (screen_off(c)) return -1;
    return h2d(c, c->S.lsc_d, d, pairs * c->P.M * kP * 8);
}
int dlsc_set_sfc(dlsc_ctx* c, const float* sfc, const uint8_t* init_flag) {
    if (!c) return fail("null ctx");
    CK(cudaSetDevice(c->device));
    if (sfc) CK(cudaMemcpyAsync(c->S.sfc, sfc, (size_t)c->P.NL * c->P.M * 24, cudaMemcpyHostToDevice, c->stream));
    if (init_flag) CK(cudaMemcpyAsync(c->S.sfc_init, init_flag, (size_t)c->P.NL, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
int dlsc_enable_timing(dlsc_ctx* c, int on) {
    if (!c) return fail("null ctx");
    c->timing = on != 0;
    memset(c->t_ms, 0, sizeof(c->t_ms));
    c->t_steps = 0;
    c->ev_used = 0;
    return 0;
}
int dlsc_get_timings(dlsc_ctx* c, double ms[DLSC_N_STAGES], int* n_steps) {
    if (!c || !ms) return fail("null argument");
    CK(cudaSetDevice(c->device));
    if (c->ev_used > 0) {
        CK(cudaStreamSynchronize(c->stream));
        for (int s2 = 0; s2 < c->ev_used; s2++) {
            cudaEvent_t* ev = c->evpool.data() + (size_t)s2 * (DLSC_N_STAGES + 1);
            for (int i = 0; i < DLSC_N_STAGES; i++) {
                float t = 0.f;
                CK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
                c->t_ms[i] += t;
            }
            c->t_steps++;
        }
        c->ev_used = 0;
    }
    for (int i = 0; i < DLSC_N_STAGES; i++) ms[i] = c->t_steps ? c->t_ms[i] / c->t_steps : 0.0;
    if (n_steps) *n_steps = c->t_steps;
    memset(c->t_ms, 0, sizeof(c->t_ms));
    c->t_steps = 0;
    return 0;
}
int64_t dlsc_launch_count(const dlsc_ctx* c) { return c ? c->launches : 0; }
int dlsc_get_counters(dlsc_ctx* c, int64_t counters[DLSC_N_COUNTERS]) {
    return c ? d2h(c, counters, c->S.counters, DLSC_N_COUNTERS * sizeof(int64_t)) : fail("null ctx");
}
int dlsc_set_waypoints_device(dlsc_ctx* c, const float* p) {
    if (!c || !p) return fail("dlsc_set_waypoints_device: null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->S.waypoint, p, (size_t)c->P.NL * 12, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}
int dlsc_measure_fp64_peak(dlsc_ctx* c, double* tflops) {
    if (!c || !tflops) return fail("dlsc_measure_fp64_peak: null argument");
    CK(cudaSetDevice(c->device));
    *tflops = measure_fp64_peak(c->device, c->stream);
    c->launches += 4;
    CK(cudaGetLastError());
    return 0;
}
int dlsc_gjk_batch(dlsc_ctx* c, const double* pts, int n, double* v, int32_t* iters, int32_t* simplex, uint64_t* leaves) {
    if (!c || !pts || !v || n < 0) return fail("dlsc_gjk_batch: bad argument");
    if (n == 0) return 0;
    CK(cudaSetDevice(c->device));
    double *d_pts = nullptr, *d_v = nullptr;
    int32_t *d_it = nullptr, *d_sn = nullptr;
    unsigned long long* d_lv = nullptr;
    int rc = 0;
    auto body = [&]() -> int {
        CK(cudaMalloc(&d_pts, (size_t)n * 18 * 8)); CK(cudaMalloc(&d_v, (size_t)n * 3 * 8));
        CK(cudaMalloc(&d_it, (size_t)n * 4)); CK(cudaMalloc(&d_sn, (size_t)n * 4)); CK(cudaMalloc(&d_lv, (size_t)n * 8));
        CK(cudaMemcpyAsync(d_pts, pts, (size_t)n * 18 * 8, cudaMemcpyHostToDevice, c->stream));
        launch_gjk_batch(d_pts, n, d_v, d_it, d_sn, d_lv, c->stream);
        c->launches++;
        CK(cudaMemcpyAsync(v, d_v, (size_t)n * 3 * 8, cudaMemcpyDeviceToHost, c->stream));
        if (iters) CK(cudaMemcpyAsync(iters, d_it, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (simplex) CK(cudaMemcpyAsync(simplex, d_sn, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (leaves) CK(cudaMemcpyAsync(leaves, d_lv, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaGetLastError());
        return 0;
    };
    rc = body();
    cudaFree(d_pts); cudaFree(d_v); cudaFree(d_it); cudaFree(d_sn); cudaFree(d_lv);
    return rc;
}
float* dlsc_waypoint_device(dlsc_ctx* c) { return c ? c->S.waypoint : nullptr; }
float* dlsc_traj_device(dlsc_ctx* c) { return c ? c->S.traj : nullptr; }

}  // extern "C"
