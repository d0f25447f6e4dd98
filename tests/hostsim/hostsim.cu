// hostsim.cu -- TEST INFRASTRUCTURE ONLY (never loaded by the product package).
//
// Runs the very same __host__ __device__ cores the sm_100a kernels execute (dlsc_stages.cuh,
// dlsc_qp.cuh) serially on the CPU, one "lane" / one "thread" per cooperative group, behind the same
// C-ABI names as libdlsc_b200.so.  Purpose: check the kernel arithmetic against the oracle in the
// CPU-only test tier (-m "not gpu"), so that a GPU box is only needed for what a GPU adds
// (SIMT mapping, synchronisation, shuffles).  It is not a fallback: the product loader
// (dlsc_gc_planner_b200/capi.py) only ever opens libdlsc_b200.so.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dlsc_b200.h"
#include "../../dlsc_gc_planner_b200/csrc/dlsc_edt.cuh"
#include "../../dlsc_gc_planner_b200/csrc/dlsc_qp_gi.cuh"
#include "../../dlsc_gc_planner_b200/csrc/dlsc_stages.cuh"

using namespace dlsc;

static std::string g_err;
static int fail(const char* m) { g_err = m; return -1; }

struct dlsc_ctx {
    DevParams P;
    RecLayout rl;
    QpTabHost th;
    QpTab T;
    int seq = 0;
    std::vector<float> rec, acc, waypoint, goal_new, pred_traj, init_traj, lsc_normal, lsc_anchor_last, sfc, traj;
    std::vector<uint8_t> disturbed, sfc_init;
    std::vector<float> lsc_near;
    std::vector<double> radius, downwash, max_vel, max_acc, nominal_vel, lsc_d, qp_x, cost, viol, scratch, smem, smem_gi;
    std::vector<int32_t> nbr_idx, nbr_cnt, qp_iters, status;
    std::vector<int4> cells;
    std::vector<float> centre;
    std::vector<uint8_t> vmask;
    std::vector<int32_t> sat;
    bool mask_dirty = false;
    EdtDev edt;
    bool have_edt = false;
    int64_t counters[DLSC_N_COUNTERS] = {0};
    float* traj_host = nullptr;
    // dynamic obstacles
    std::vector<float> dyn_pos, dyn_vel, comm_box;
    std::vector<double> dyn_radius, dyn_downwash, dyn_max_acc, dyn_size, qp_slack;
    std::vector<uint8_t> trap;
};

extern "C" {

const char* dlsc_last_error(void) { return g_err.c_str(); }
int dlsc_abi_version(void) { return DLSC_ABI_VERSION; }
int dlsc_device_count(void) { return 0; }

int dlsc_create(const dlsc_params* hp, int n_agents, int agent_begin, int n_local, int device, dlsc_ctx** out) {
    (void)device;
    if (hp->n != 5 || hp->phi != 3) return fail("only n = 5, phi = 3");
    if (hp->dim * (3 * hp->M - 2) > 128 || hp->M > kMaxM || hp->M < 2) return fail("bad M");
    dlsc_ctx* c = new dlsc_ctx();
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
    double time = 0;
    for (int e = 0; e < hp->M * kP; e++) { P.tk[e] = (float)time; time += hp->dt / hp->n; }
    const size_t N = n_agents, NL = n_local, K = hp->max_nbr, M = hp->M, npt = M * kP;
    c->rec.assign(N * P.rec, 0.f); c->acc.assign(NL * 3, 0.f); c->waypoint.assign(NL * 3, 0.f); c->goal_new.assign(NL * 3, 0.f);
    c->disturbed.assign(NL, 0); c->sfc_init.assign(NL, 1);
    c->radius.assign(NL, 0); c->downwash.assign(NL, 0); c->max_vel.assign(NL, 0); c->max_acc.assign(NL, 0);
    c->nominal_vel.assign(NL, 0);
    c->pred_traj.assign((N + kMaxDyn) * npt * 3, 0.f);
    c->comm_box.assign(NL * 6, 0.f); c->trap.assign(NL, 0); c->qp_slack.assign(NL * kMaxDyn * M, 0.0); c->init_traj.assign(NL * npt * 3, 0.f);
    c->nbr_idx.assign(NL * K, 0); c->nbr_cnt.assign(NL, 0);
    c->lsc_normal.assign(NL * K * M * 3, 0.f); c->lsc_d.assign(NL * K * M * kP, 0.0);
    c->lsc_anchor_last.assign(NL * K * 3, 0.f);
    c->lsc_near.assign(NL * K * M, 0.0f);
    c->sfc.assign(NL * M * 6, 0.f); c->traj.assign(NL * npt * 3, 0.f);
    c->qp_x.assign(NL * hp->dim * npt, 0.0); c->cost.assign(NL, 0); c->viol.assign(NL, 0);
    c->qp_iters.assign(NL, 0); c->status.assign(NL, 0);
    build_qp_tables(hp->M, hp->dim, hp->dt, hp->w_control, hp->w_terminal, hp->comm_range > 0, c->th);
    const QpTabHost& h = c->th;
    QpTab& T = c->T;
    T.D = h.D; T.M = h.M; T.nyd = h.nyd; T.ny = h.ny; T.npt = h.npt; T.nx = h.nx; T.np = h.np; T.ntri = h.ntri;
    T.ntri_local = h.ntri_local; T.use_comm = h.use_comm;
    T.xm_nv = h.xm_nv.data(); T.xm_cidx = h.xm_cidx.data(); T.xm_idx = h.xm_idx.data(); T.xm_coef = h.xm_coef.data();
    T.pr_fam = h.pr_fam.data(); T.pr_axis = h.pr_axis.data(); T.pr_nnz = h.pr_nnz.data(); T.pr_pt = h.pr_pt.data();
    T.pr_idx = h.pr_idx.data(); T.pr_val = h.pr_val.data(); T.pr_cc = h.pr_cc.data();
    T.yi_ptr = h.yi_ptr.data(); T.yi_row = h.yi_row.data(); T.yi_coef = h.yi_coef.data();
    T.yp_ptr = h.yp_ptr.data(); T.yp_pt = h.yp_pt.data(); T.yp_coef = h.yp_coef.data();
    T.wi_ptr = h.wi_ptr.data(); T.wi_row = h.wi_row.data(); T.wi_coef = h.wi_coef.data();
    T.wp_ptr = h.wp_ptr.data(); T.wp_pt = h.wp_pt.data(); T.wp_coef = h.wp_coef.data();
    T.H1 = h.H1.data(); T.Q2 = h.Q2.data(); T.tri_p = h.tri_p.data(); T.nz_e = h.nz_e.data(); T.nnzw = h.nnzw;
    T.pr_desc = h.pr_desc.data(); T.nz_hdr = reinterpret_cast<const uint4*>(h.nz_hdr.data()); T.nz_h = h.nz_h.data(); T.Hinv = h.Hinv.data(); T.Y0 = h.Y0.data();
    {
        const int M = h.M, MP = M * kP;
        T.row_npl = h.np / h.D; T.row_bv = MP - 3; T.row_ba = T.row_bv + (M * 5 - 2); T.row_bc = T.row_ba + (M * 4 - 1);
        T.scv = h.scv; T.sca = h.sca;
    }
    c->scratch.assign(qp_scratch_doubles(T, P.K), 0.0);
    c->smem.assign(qp_smem_bytes(T, P.K) / 8 + 8, 0.0);
    c->smem_gi.assign(gi_smem_doubles(T, P.K) + 8, 0.0);
    memset(&c->edt, 0, sizeof(c->edt));
    c->edt.res = hp->world_res; c->edt.inv_res = 1.0 / hp->world_res;
    *out = c;
    return 0;
}
void dlsc_destroy(dlsc_ctx* c) { delete c; }

int dlsc_set_edt(dlsc_ctx* c, const float* dist, const int32_t* obst, const int32_t dims[3],
                 const int32_t min_key[3], double res) {
    const size_t nc = (size_t)dims[0] * dims[1] * dims[2];
    c->cells.resize(nc);
    for (size_t i = 0; i < nc; i++) {
        union { float f; int i; } u; u.f = dist[i];
        c->cells[i].x = u.i; c->cells[i].y = obst[3 * i]; c->cells[i].z = obst[3 * i + 1]; c->cells[i].w = obst[3 * i + 2];
    }
    for (int k = 0; k < 3; k++) { c->edt.dims[k] = dims[k]; c->edt.min_key[k] = min_key[k]; }
    c->edt.res = res; c->edt.inv_res = 1.0 / res; c->edt.cells = c->cells.data();
    c->centre.clear();
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < dims[k]; i++) c->centre.push_back((float)(((double)(i + min_key[k]) + 0.5) * res));
    c->edt.centre[0] = c->centre.data(); c->edt.centre[1] = c->centre.data() + dims[0];
    c->edt.centre[2] = c->centre.data() + dims[0] + dims[1];
    c->have_edt = true;
    c->mask_dirty = true;
    return 0;
}

// the device grid construction (dlsc_kernels_edt.cu) cell by cell with the same cores
int dlsc_edt_dims(const dlsc_ctx* c, int32_t dims[3], int32_t min_key[3]) {
    const double inv = 1.0 / c->P.world_res;
    for (int k = 0; k < 3; k++) {
        const int lo = (int)std::floor(inv * c->P.world_min[k]);
        const int hi = (int)std::floor(inv * c->P.world_max[k]);
        min_key[k] = lo; dims[k] = hi - lo + 1;
    }
    return 0;
}
static int edt_from_occ(dlsc_ctx* c, const std::vector<uint8_t>& occ, const int32_t dims[3], const int32_t mk[3], double maxdist) {
    const double res = c->P.world_res;
    const int maxd = (int)(maxdist / res + 1), R = maxd - 1;
    EdtDistTab tab; float cap;
    if (!edt_make_tab(res, maxd, &tab, &cap)) return fail("dlsc_build_edt: maxdist / resolution + 1 must be in [1, 16] cells");
    const size_t nc = occ.size();
    std::vector<uint32_t> ta(nc), tb(nc);
    for (size_t i = 0; i < nc; i++) ta[i] = edt_pass_z_cell(occ.data(), i, dims[2], R);
    for (size_t i = 0; i < nc; i++) tb[i] = edt_pass_y_cell(ta.data(), i, dims[1], dims[2], R);
    std::vector<float> dist(nc); std::vector<int32_t> obst(3 * nc);
    for (size_t i = 0; i < nc; i++) {
        const EdtRecord r = edt_pass_x_cell(tb.data(), i, dims[0], dims[1], dims[2], R, maxd * maxd, cap, tab);
        memcpy(&dist[i], &r.x, 4); obst[3 * i] = r.y; obst[3 * i + 1] = r.z; obst[3 * i + 2] = r.w;
    }
    return dlsc_set_edt(c, dist.data(), obst.data(), dims, mk, res);
}
int dlsc_build_edt(dlsc_ctx* c, const float* boxes, int nb, double maxdist) {
    int32_t dims[3], mk[3];
    dlsc_edt_dims(c, dims, mk);
    std::vector<uint8_t> occ((size_t)dims[0] * dims[1] * dims[2], 0);
    const double res = c->P.world_res, inv = 1.0 / res;
    for (int b = 0; b < nb; b++) {
        int s[3], e[3];
        for (int k = 0; k < 3; k++) edt_box_range(boxes[6 * b + k], boxes[6 * b + 3 + k], res, &s[k], &e[k]);
        for (int i = s[0]; i < e[0]; i++)
            for (int j = s[1]; j < e[1]; j++)
                for (int k = s[2]; k < e[2]; k++) {
                    const int mx = edt_voxel_cell(i, res, inv, mk[0]), my = edt_voxel_cell(j, res, inv, mk[1]),
                              mz = edt_voxel_cell(k, res, inv, mk[2]);
                    if (mx >= 0 && mx < dims[0] && my >= 0 && my < dims[1] && mz >= 0 && mz < dims[2])
                        occ[((size_t)mx * dims[1] + my) * dims[2] + mz] = 1;
                }
    }
    return edt_from_occ(c, occ, dims, mk, maxdist);
}
int dlsc_build_edt_occupancy(dlsc_ctx* c, const uint8_t* occ_in, double maxdist) {
    int32_t dims[3], mk[3];
    dlsc_edt_dims(c, dims, mk);
    std::vector<uint8_t> occ(occ_in, occ_in + (size_t)dims[0] * dims[1] * dims[2]);
    return edt_from_occ(c, occ, dims, mk, maxdist);
}
int dlsc_get_edt(dlsc_ctx* c, float* dist, int32_t* obst) {
    if (!c->have_edt) return fail("dlsc_get_edt: no grid");
    for (size_t i = 0; i < c->cells.size(); i++) {
        memcpy(&dist[i], &c->cells[i].x, 4);
        obst[3 * i] = c->cells[i].y; obst[3 * i + 1] = c->cells[i].z; obst[3 * i + 2] = c->cells[i].w;
    }
    return 0;
}
double dlsc_edt_build_ms(const dlsc_ctx*) { return 0.0; }

// same lazy build as dlsc_api.cu build_vertex_mask
static void build_vertex_mask(dlsc_ctx* c) {
    EdtDev& E = c->edt;
    c->mask_dirty = false;
    E.vmask = nullptr; E.sat = nullptr; E.zs = edt_mask_zs(E.dims[2]); E.mask_margin = c->radius.empty() ? 0.0 : c->radius[0];
    const char* env = getenv("DLSC_SFC_MASK");
    if ((env && env[0] == '0') || E.zs > kSfcZsMax || !E.cells) return;
    c->vmask.assign((size_t)(E.dims[0] + 1) * (E.dims[1] + 1) * E.zs, 0);
    bool unsafe = false;
    for (int vx = 0; vx <= E.dims[0]; vx++)
        for (int vy = 0; vy <= E.dims[1]; vy++)
            for (int vz = 0; vz <= E.dims[2]; vz++)
                c->vmask[((size_t)vx * (E.dims[1] + 1) + vy) * E.zs + vz] = edt_vertex_mask(E, vx, vy, vz, E.mask_margin, &unsafe);
    if (unsafe) return;
    E.vmask = c->vmask.data();
    const char* env_sat = getenv("DLSC_SFC_SAT");
    if (env_sat && env_sat[0] == '0') return;
    const int n0 = E.dims[0] + 2, n1 = E.dims[1] + 2, n2 = E.dims[2] + 2;
    c->sat.assign((size_t)n0 * n1 * n2, 0);
    for (int i = 1; i < n0; i++)
        for (int j = 1; j < n1; j++)
            for (int k = 1; k < n2; k++) c->sat[sat_index(E, i, j, k)] = sat_indicator(E, i - 1, j - 1, k - 1);
    for (int axis = 2; axis >= 0; axis--)
        for (int i = 0; i < n0; i++)
            for (int j = 0; j < n1; j++)
                for (int k = 0; k < n2; k++) {
                    const int pi = i - (axis == 0), pj = j - (axis == 1), pk = k - (axis == 2);
                    if (pi >= 0 && pj >= 0 && pk >= 0) c->sat[sat_index(E, i, j, k)] += c->sat[sat_index(E, pi, pj, pk)];
                }
    E.sat = c->sat.data();
}

int dlsc_set_agent_props(dlsc_ctx* c, const dlsc_agent_props* p) {
    const size_t NL = c->P.NL;
    if (p->radius) { c->radius.assign(p->radius, p->radius + NL); c->mask_dirty = true; }
    if (p->downwash) c->downwash.assign(p->downwash, p->downwash + NL);
    if (p->max_vel) c->max_vel.assign(p->max_vel, p->max_vel + NL);
    if (p->max_acc) c->max_acc.assign(p->max_acc, p->max_acc + NL);
    if (p->nominal_vel) c->nominal_vel.assign(p->nominal_vel, p->nominal_vel + NL);
    return 0;
}

int dlsc_bind_traj_host(dlsc_ctx* c, float* host) { c->traj_host = host; return 0; }
int dlsc_set_obstacles(dlsc_ctx* c, const dlsc_obstacles* o, const dlsc_obstacle_params* op) {
    const int n = o ? o->n : 0;
    if (n < 0 || n > kMaxDyn) return fail("too many obstacles");
    if (n == 0) { c->P.n_dyn = 0; return 0; }
    if (!(op->slack_collision_weight > 0) || n >= c->P.K) return fail("bad obstacle setup");
    c->dyn_pos.assign(o->pos, o->pos + 3 * n); c->dyn_vel.assign(o->vel, o->vel + 3 * n);
    c->dyn_radius.assign(o->radius, o->radius + n); c->dyn_downwash.assign(o->downwash, o->downwash + n);
    c->dyn_max_acc.assign(o->max_acc, o->max_acc + n);
    c->dyn_size.assign((size_t)n * c->P.M * kP, 0.0);
    for (int i = 0; i < n; i++)
        dyn_obstacle_sizes(c->P, op->size_prediction != 0, op->uncertainty_horizon, o->radius[i], o->max_acc[i], c->dyn_size.data() + (size_t)i * c->P.M * kP);
    c->P.n_dyn = n; c->P.slack_w = op->slack_collision_weight; c->P.dyn_horizon = op->uncertainty_horizon;
    return 0;
}
int dlsc_get_slack(dlsc_ctx* c, double* slack) {
    const int nd = c->P.n_dyn, M = c->P.M;
    for (int a = 0; a < c->P.NL; a++) memcpy(slack + (size_t)a * nd * M, c->qp_slack.data() + (size_t)a * kMaxDyn * M, (size_t)nd * M * 8);
    return 0;
}
int dlsc_get_trap(dlsc_ctx* c, uint8_t* trap) { memcpy(trap, c->trap.data(), c->trap.size()); return 0; }
int dlsc_get_obstacle_pred(dlsc_ctx* c, float* t) {
    memcpy(t, c->pred_traj.data() + (size_t)c->P.N * c->P.M * kP * 3, (size_t)c->P.n_dyn * c->P.M * kP * 12);
    return 0;
}

int dlsc_reset(dlsc_ctx* c, const float* start) {
    const DevParams& P = c->P;
    const int npt = P.M * kP;
    for (int la = 0; la < P.NL; la++) {
        float* rec = c->rec.data() + (size_t)(P.begin + la) * P.rec;
        for (int e = 0; e < npt; e++) for (int k = 0; k < 3; k++) rec[e * 3 + k] = start[la * 3 + k];
        const int o = npt * 3;
        for (int k = 0; k < 3; k++) { rec[o + k] = start[la * 3 + k]; rec[o + 3 + k] = 0.f; rec[o + 6 + k] = start[la * 3 + k]; }
        rec[o + 9] = (float)c->radius[la]; rec[o + 10] = (float)c->downwash[la];
        for (int e = o + 11; e < P.rec; e++) rec[e] = 0.f;
        for (int k = 0; k < 3; k++) { c->acc[la * 3 + k] = 0.f; c->waypoint[la * 3 + k] = start[la * 3 + k]; c->goal_new[la * 3 + k] = start[la * 3 + k]; }
        c->disturbed[la] = 0; c->sfc_init[la] = 1; c->status[la] = 0;
        for (int e = 0; e < 6; e++) c->comm_box[(size_t)la * 6 + e] = 0.f;
        for (int e = 0; e < npt * 3; e++) c->traj[(size_t)la * npt * 3 + e] = rec[e];
    }
    c->seq = 0;
    return 0;
}

int dlsc_set_groups(dlsc_ctx* c, const int32_t* group) {
    const DevParams& P = c->P;
    for (int i = 0; i < P.NL; i++) c->rec[(size_t)(P.begin + i) * P.rec + c->rl.group] = (float)group[i];
    return 0;
}

int dlsc_set_agents(dlsc_ctx* c, const dlsc_agents* a) {
    const DevParams& P = c->P;
    for (int la = 0; la < P.NL; la++) {
        float* rec = c->rec.data() + (size_t)(P.begin + la) * P.rec + c->rl.pos;
        for (int k = 0; k < 3; k++) {
            if (a->pos) rec[k] = a->pos[la * 3 + k];
            if (a->vel) rec[3 + k] = a->vel[la * 3 + k];
            if (a->acc) c->acc[la * 3 + k] = a->acc[la * 3 + k];
            if (a->waypoint) c->waypoint[la * 3 + k] = a->waypoint[la * 3 + k];
        }
        if (a->disturbed) c->disturbed[la] = a->disturbed[la];
    }
    return 0;
}

float* dlsc_records_device(dlsc_ctx* c) { return c->rec.data(); }
int dlsc_record_floats(const dlsc_ctx* c) { return c->P.rec; }
int dlsc_set_records(dlsc_ctx* c, int first, int count, const float* host) {
    memcpy(c->rec.data() + (size_t)first * c->P.rec, host, (size_t)count * c->P.rec * 4);
    return 0;
}
int dlsc_get_records(dlsc_ctx* c, int first, int count, float* host) {
    memcpy(host, c->rec.data() + (size_t)first * c->P.rec, (size_t)count * c->P.rec * 4);
    return 0;
}

int dlsc_run_stages(dlsc_ctx* c, int mask) {
    const DevParams& P = c->P;
    const int M = P.M, npt = M * kP, K = P.K;
    const int seq = c->seq + 1;
    Group g; g.lane = 0; g.width = 1; g.block = false;
    if ((mask & DLSC_STAGE_SFC) && P.use_sfc && !c->have_edt) return fail("no EDT");
    if (mask & (DLSC_STAGE_NBR | DLSC_STAGE_LSC | DLSC_STAGE_SFC | DLSC_STAGE_QP)) memset(c->counters, 0, sizeof(c->counters));
    if (mask & DLSC_STAGE_PREDICT)
        for (int a = 0; a < P.N; a++) {
            const int la = a - P.begin;
            const bool local = la >= 0 && la < P.NL;
            for (int pt = 0; pt < npt; pt++)
                predict_point(P, c->rec.data() + (size_t)a * P.rec, seq, pt, local ? c->disturbed[la] != 0 : false,
                              c->pred_traj.data() + (size_t)a * npt * 3,
                              local ? c->init_traj.data() + (size_t)la * npt * 3 : nullptr);
            if (local) c->status[la] = 0;
        }
    const int nd = P.n_dyn;
    if ((mask & DLSC_STAGE_PREDICT) && nd > 0) {                       // k_dyn_predict
        for (int o = 0; o < nd; o++)
            for (int pt = 0; pt < npt; pt++)
                v3_store(c->pred_traj.data() + ((size_t)(P.N + o) * npt + pt) * 3,
                         v3_load(c->dyn_pos.data() + 3 * o) + v3_load(c->dyn_vel.data() + 3 * o) * P.tk[pt]);
        for (int la = 0; la < P.NL; la++) {
            comm_box_update(P, c->sfc_init[la] != 0 || c->disturbed[la] != 0, v3_load(c->waypoint.data() + la * 3), c->comm_box.data() + (size_t)la * 6);
            for (int o = 0; o < nd; o++) c->nbr_idx[(size_t)la * K + o] = P.N + o;
        }
    }
    if (mask & DLSC_STAGE_NBR)
        for (int la = 0; la < P.NL; la++) {
            const int Kc = K - nd;
            const int cnt = neighbours_agent(g, P, c->rec.data(), P.begin + la, c->nbr_idx.data() + (size_t)la * K + nd);
            c->nbr_cnt[la] = nd + (cnt < Kc ? cnt : Kc);
            if (cnt > Kc) c->status[la] |= kStNbrOverflow;
            c->counters[0] += c->nbr_cnt[la] - nd;
        }
    if (mask & DLSC_STAGE_LSC)
        for (int la = 0; la < P.NL; la++)
            for (int cc = 0; cc < c->nbr_cnt[la]; cc++) {
                const int j = c->nbr_idx[(size_t)la * K + cc];
                if (cc < nd) {                                         // k_lsc_dyn
                    const size_t prd = (size_t)la * K + cc;
                    for (int m = 0; m < M; m++)
                        lsc_dynamic_segment(c->init_traj.data() + ((size_t)la * npt + m * kP) * 3, c->pred_traj.data() + ((size_t)(P.N + cc) * npt + m * kP) * 3,
                                            c->dyn_size.data() + ((size_t)cc * M + m) * kP, c->radius[la], c->dyn_radius[cc], c->dyn_downwash[cc],
                                            c->lsc_normal.data() + (prd * M + m) * 3, c->lsc_d.data() + (prd * M + m) * kP, c->lsc_near.data() + prd * M + m);
                    continue;
                }
                const float* rec_a = c->rec.data() + (size_t)(P.begin + la) * P.rec;
                const float* rec_j = c->rec.data() + (size_t)j * P.rec;
                const int og = npt * 3 + 6;
                const size_t pr = (size_t)la * K + cc;
                for (int m = 0; m < M; m++) {
                    int it = 0;
                    lsc_segment(P, c->init_traj.data() + (size_t)la * npt * 3, c->pred_traj.data() + (size_t)j * npt * 3,
                                v3_load(rec_a + og), v3_load(rec_j + og), c->radius[la], c->downwash[la], rec_j[og + 3],
                                rec_j[og + 4], m, c->lsc_normal.data() + (pr * M + m) * 3,
                                c->lsc_d.data() + (pr * M + m) * kP, c->lsc_anchor_last.data() + pr * 3, &it,
                                c->lsc_near.data() + pr * M + m);
                    c->counters[1] += it;
                }
            }
    if ((mask & DLSC_STAGE_SFC) && P.use_sfc && c->mask_dirty) build_vertex_mask(c);
    if ((mask & DLSC_STAGE_SFC) && P.use_sfc)
        for (int la = 0; la < P.NL; la++) {
            const float* rec = c->rec.data() + (size_t)(P.begin + la) * P.rec;
            const bool init = c->sfc_init[la] != 0 || c->disturbed[la] != 0;
            long long lookups[6] = {0, 0, 0, 0, 0, 0};
            SfcTab memo;
            const int st = sfc_agent(g, P, c->edt, init, v3_load(rec + npt * 3), c->init_traj.data() + (size_t)la * npt * 3,
                                     v3_load(rec + npt * 3 + 6), v3_load(c->waypoint.data() + la * 3), c->radius[la],
                                     c->max_vel[la], c->sfc.data() + (size_t)la * M * 6, &memo, lookups);
            c->sfc_init[la] = 0;
            c->status[la] |= st;
            c->counters[2] += lookups[0]; c->counters[5] += lookups[1]; c->counters[6] += lookups[2]; c->counters[7] += lookups[3]; c->counters[8] += lookups[4];
        }
    if ((mask & DLSC_STAGE_GOAL) && nd > 0)                            // k_trap
        for (int la = 0; la < P.NL; la++) {
            const float* rec = c->rec.data() + (size_t)(P.begin + la) * P.rec;
            const size_t pr = (size_t)la * K;
            DynObs O; O.pos = c->dyn_pos.data(); O.vel = c->dyn_vel.data(); O.radius = c->dyn_radius.data(); O.downwash = c->dyn_downwash.data();
            O.max_acc = c->dyn_max_acc.data(); O.size = c->dyn_size.data();
            c->trap[la] = (uint8_t)waypoint_trap(g, P, O, v3_load(rec + npt * 3 + 6), v3_load(c->waypoint.data() + la * 3),
                                                 c->sfc.data() + ((size_t)la * M + (M - 1)) * 6, c->comm_box.data() + (size_t)la * 6, c->nbr_cnt[la],
                                                 c->lsc_normal.data() + pr * M * 3, c->lsc_d.data() + pr * M * kP, c->lsc_anchor_last.data() + pr * 3,
                                                 c->radius[la]);
        }
    if (mask & DLSC_STAGE_GOAL)
        for (int la = 0; la < P.NL; la++) {
            const float* rec = c->rec.data() + (size_t)(P.begin + la) * P.rec;
            V3 goal = v3_load(rec + npt * 3 + 6);
            const size_t pr = (size_t)la * K + nd;
            const int st = goal_agent(g, P, c->disturbed[la] != 0, v3_load(rec + npt * 3), v3_load(c->waypoint.data() + la * 3),
                                      c->sfc.data() + ((size_t)la * M + (M - 1)) * 6, c->nbr_cnt[la] - nd,
                                      c->lsc_normal.data() + pr * M * 3, c->lsc_d.data() + pr * M * kP,
                                      c->lsc_anchor_last.data() + pr * 3, goal);
            v3_store(c->goal_new.data() + la * 3, goal);
            c->status[la] |= st;
        }
    else if (mask & DLSC_STAGE_QP)
        for (int la = 0; la < P.NL; la++)
            for (int k = 0; k < 3; k++) c->goal_new[la * 3 + k] = c->rec[(size_t)(P.begin + la) * P.rec + npt * 3 + 6 + k];
    if (mask & DLSC_STAGE_QP) {
        QpSmem sm;
        qp_smem_carve(c->T, K, c->smem.data(), sm);
        Cta cta; cta.tid = 0; cta.nthr = 1; cta.red = sm.red;
        for (int la = 0; la < P.NL; la++) {
            const float* rec = c->rec.data() + (size_t)(P.begin + la) * P.rec;
            QpIn in;
            in.pos = v3_load(rec + npt * 3); in.vel = v3_load(rec + npt * 3 + 3);
            in.acc = v3_load(c->acc.data() + la * 3); in.goal = v3_load(c->goal_new.data() + la * 3);
            in.wp = v3_load(c->waypoint.data() + la * 3);
            in.radius = c->radius[la]; in.max_vel = c->max_vel[la]; in.max_acc = c->max_acc[la];
            in.nominal_vel = c->nominal_vel[la];
            in.sfc = c->sfc.data() + (size_t)la * M * 6;
            in.init_traj = c->init_traj.data() + (size_t)la * npt * 3;
            in.K = c->nbr_cnt[la];
            const size_t pr = (size_t)la * K;
            in.nbr_idx = c->nbr_idx.data() + pr;
            in.normal = c->lsc_normal.data() + pr * M * 3;
            in.d = c->lsc_d.data() + pr * M * kP;
            in.anchor_last = c->lsc_anchor_last.data() + pr * 3;
            in.near = c->lsc_near.data() + pr * M;
            in.pred_traj = c->pred_traj.data();
            long long rows = 0;
            QpOut out;
            out.traj = c->traj.data() + (size_t)la * npt * 3;
            out.traj_host = c->traj_host ? c->traj_host + (size_t)la * npt * 3 : nullptr;
            out.x = c->qp_x.data() + (size_t)la * c->T.nx;
            out.cost = &c->cost[la]; out.viol = &c->viol[la]; out.iters = &c->qp_iters[la]; out.status = &c->status[la];
            out.rows = &rows;
            out.slack = nd > 0 ? c->qp_slack.data() + (size_t)la * kMaxDyn * M : nullptr;
            bool done = false;
            if (P.qp_solver != 1) {
                QpSmem sg;
                gi_smem_carve(c->T, c->smem_gi.data(), sg);
                Cta cg; cg.tid = 0; cg.nthr = 1; cg.red = sg.red;
                // same two kernels as launch_qp: the warp-per-agent fast path, then the seeded dual active set
                std::vector<double> fast_mem(fast_smem_doubles(c->T, P.K) + 2);
                QpSmem sf;
                fast_smem_carve(c->T, fast_mem.data(), sf);
                Cta cf; cf.tid = 0; cf.nthr = 1; cf.red = nullptr; cf.warp = true;
                double seed[4] = {0, 0, 0, 0};
                done = nd > 0 ? qp_agent_fast<true>(cf, P, c->T, in, out, sf, seed) : qp_agent_fast<false>(cf, P, c->T, in, out, sf, seed);
                if (!done) done = nd > 0 ? qp_agent_gi<kGiQ, true>(cg, P, c->T, in, out, sg, seed) : qp_agent_gi<kGiQ, false>(cg, P, c->T, in, out, sg, seed);
            }
            if (!done) {
                if (nd > 0) qp_agent<true>(cta, P, c->T, in, out, sm, c->scratch.data(), P.qp_solver != 1);
                else qp_agent<false>(cta, P, c->T, in, out, sm, c->scratch.data(), P.qp_solver != 1);
            }
            c->counters[3] += c->qp_iters[la];
            c->counters[4] += rows;
        }
    }
    return 0;
}

int dlsc_step(dlsc_ctx* c) {
    const int rc = dlsc_run_stages(c, DLSC_STAGE_ALL);
    if (rc == 0) c->seq++;
    return rc;
}

static void advance_impl(dlsc_ctx* c, bool move) {
    const DevParams& P = c->P;
    const int npt = P.M * kP;
    for (int la = 0; la < P.NL; la++) {
        float* rec = c->rec.data() + (size_t)(P.begin + la) * P.rec;
        const float* tr = c->traj.data() + (size_t)la * npt * 3;
        if (move) {
            float st[9];
            state_at(P, tr, P.dt, st);
            for (int k = 0; k < 3; k++) { rec[npt * 3 + k] = st[k]; rec[npt * 3 + 3 + k] = st[3 + k]; c->acc[la * 3 + k] = st[6 + k]; }
        }
        for (int e = 0; e < npt * 3; e++) rec[e] = tr[e];
        for (int k = 0; k < 3; k++) rec[npt * 3 + 6 + k] = c->goal_new[la * 3 + k];
    }
}
int dlsc_advance(dlsc_ctx* c) { advance_impl(c, true); return 0; }
int dlsc_publish_records(dlsc_ctx* c) { advance_impl(c, false); return 0; }
int dlsc_sync(dlsc_ctx*) { return 0; }
int dlsc_get_seq(const dlsc_ctx* c) { return c->seq; }
int dlsc_set_seq(dlsc_ctx* c, int s) { c->seq = s; return 0; }

#define GET(name, type, vec) int name(dlsc_ctx* c, type* o) { memcpy(o, c->vec.data(), c->vec.size() * sizeof(type)); return 0; }
GET(dlsc_get_traj, float, traj)
GET(dlsc_get_qp_x, double, qp_x)
GET(dlsc_get_cost, double, cost)
GET(dlsc_get_violation, double, viol)
GET(dlsc_get_qp_iters, int32_t, qp_iters)
GET(dlsc_get_status, int32_t, status)
GET(dlsc_get_init_traj, float, init_traj)
int dlsc_get_pred_traj(dlsc_ctx* c, float* t) { memcpy(t, c->pred_traj.data(), (size_t)c->P.N * c->P.M * kP * 12); return 0; }
GET(dlsc_get_sfc, float, sfc)
#undef GET

int dlsc_get_goal(dlsc_ctx* c, float* goal) { memcpy(goal, c->goal_new.data(), c->goal_new.size() * 4); return 0; }
int dlsc_get_state(dlsc_ctx* c, float* pos, float* vel, float* acc) {
    const DevParams& P = c->P;
    for (int la = 0; la < P.NL; la++)
        for (int k = 0; k < 3; k++) {
            if (pos) pos[la * 3 + k] = c->rec[(size_t)(P.begin + la) * P.rec + c->rl.pos + k];
            if (vel) vel[la * 3 + k] = c->rec[(size_t)(P.begin + la) * P.rec + c->rl.vel + k];
            if (acc) acc[la * 3 + k] = c->acc[la * 3 + k];
        }
    return 0;
}
int dlsc_get_neighbours(dlsc_ctx* c, int32_t* idx, int32_t* cnt) {
    if (idx) memcpy(idx, c->nbr_idx.data(), c->nbr_idx.size() * 4);
    if (cnt) memcpy(cnt, c->nbr_cnt.data(), c->nbr_cnt.size() * 4);
    return 0;
}
int dlsc_get_lsc(dlsc_ctx* c, float* normal, float* anchor, double* d) {
    const DevParams& P = c->P;
    const int npt = P.M * kP;
    if (normal) memcpy(normal, c->lsc_normal.data(), c->lsc_normal.size() * 4);
    if (d) memcpy(d, c->lsc_d.data(), c->lsc_d.size() * 8);
    if (anchor)
        for (int la = 0; la < P.NL; la++)
            for (int cc = 0; cc < P.K; cc++)
                for (int pt = 0; pt < npt; pt++) {
                    const size_t pr = (size_t)la * P.K + cc;
                    float v[3] = {0.f, 0.f, 0.f};
                    if (cc < c->nbr_cnt[la]) {
                        const int j = c->nbr_idx[pr];
                        const float* src = (pt / kP < P.M - 1 || cc < P.n_dyn) ? c->pred_traj.data() + ((size_t)j * npt + pt) * 3
                                                               : c->lsc_anchor_last.data() + pr * 3;
                        v[0] = src[0]; v[1] = src[1]; v[2] = src[2];
                    }
                    for (int k = 0; k < 3; k++) anchor[(pr * npt + pt) * 3 + k] = v[k];
                }
    return 0;
}
int dlsc_set_sfc(dlsc_ctx* c, const float* sfc, const uint8_t* init_flag) {
    if (sfc) memcpy(c->sfc.data(), sfc, c->sfc.size() * 4);
    if (init_flag) memcpy(c->sfc_init.data(), init_flag, c->sfc_init.size());
    return 0;
}
int dlsc_get_counters(dlsc_ctx* c, int64_t counters[DLSC_N_COUNTERS]) { memcpy(counters, c->counters, sizeof(c->counters)); return 0; }
int64_t dlsc_launch_count(const dlsc_ctx*) { return 0; }
int dlsc_enable_timing(dlsc_ctx*, int) { return 0; }
int dlsc_get_timings(dlsc_ctx*, double ms[DLSC_N_STAGES], int* n) { for (int i = 0; i < DLSC_N_STAGES; i++) ms[i] = 0; if (n) *n = 0; return 0; }

int dlsc_run_stages_subset(dlsc_ctx*, int, int, int) { return fail("hostsim: not supported"); }
int dlsc_set_init_traj(dlsc_ctx* c, const float* t) { std::fill(c->lsc_near.begin(), c->lsc_near.end(), 0.0f); memcpy(c->init_traj.data(), t, c->init_traj.size() * 4); return 0; }
int dlsc_set_pred_traj(dlsc_ctx* c, const float* t) { std::fill(c->lsc_near.begin(), c->lsc_near.end(), 0.0f); memcpy(c->pred_traj.data(), t, (size_t)c->P.N * c->P.M * kP * 12); return 0; }
int dlsc_set_neighbours(dlsc_ctx* c, const int32_t* idx, const int32_t* cnt) {
    std::fill(c->lsc_near.begin(), c->lsc_near.end(), 0.0f);
    memcpy(c->nbr_idx.data(), idx, c->nbr_idx.size() * 4); memcpy(c->nbr_cnt.data(), cnt, c->nbr_cnt.size() * 4); return 0;
}
int dlsc_set_lsc(dlsc_ctx* c, const float* normal, const float* anchor_last, const double* d) {
    memcpy(c->lsc_normal.data(), normal, c->lsc_normal.size() * 4);
    std::fill(c->lsc_near.begin(), c->lsc_near.end(), 0.0f);
    memcpy(c->lsc_anchor_last.data(), anchor_last, c->lsc_anchor_last.size() * 4);
    memcpy(c->lsc_d.data(), d, c->lsc_d.size() * 8); return 0;
}
int dlsc_set_waypoints_device(dlsc_ctx* c, const float* p) { memcpy(c->waypoint.data(), p, c->waypoint.size() * 4); return 0; }
int dlsc_measure_fp64_peak(dlsc_ctx*, double* t) { *t = 0.0; return 0; }
int dlsc_gjk_batch(dlsc_ctx*, const double* pts, int n, double* v, int32_t* iters, int32_t* simplex, uint64_t* leaves) {
    for (int h = 0; h < n; h++) {
        gjk::D3 c[kP];
        for (int i = 0; i < kP; i++) c[i] = gjk::d3(pts[(size_t)h * 18 + 3 * i], pts[(size_t)h * 18 + 3 * i + 1], pts[(size_t)h * 18 + 3 * i + 2]);
        gjk::MaskTrace tr; tr.m = 0;
        int it = 0, sn = 0;
        const gjk::D3 w = gjk::hull_origin<kP, gjk::MaskTrace>(c, &it, tr, &sn);
        v[(size_t)h * 3] = w.x; v[(size_t)h * 3 + 1] = w.y; v[(size_t)h * 3 + 2] = w.z;
        if (iters) iters[h] = it;
        if (simplex) simplex[h] = sn;
        if (leaves) leaves[h] = tr.m;
    }
    return 0;
}
int dlsc_set_stream(dlsc_ctx*, void*) { return 0; }
void* dlsc_get_stream(dlsc_ctx*) { return nullptr; }
int dlsc_bind_records(dlsc_ctx*, float*) { return fail("hostsim: not supported"); }
float* dlsc_waypoint_device(dlsc_ctx* c) { return c->waypoint.data(); }
float* dlsc_traj_device(dlsc_ctx* c) { return c->traj.data(); }

}  // extern "C"
