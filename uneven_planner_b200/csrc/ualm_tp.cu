// ualm_tp.cu -- host side of the throughput engine (precision 32 / 65 of include/ualm.h): slot pool, admission of batches,
// the round driver (ka -> [ks] -> kb per round, continuous batching across the batches in flight) and result gathering.
// Kernels: ualm_tp_kernels.cuh, ualm_tp_samples.cuh.  Compiled with FMA contraction ON (its own translation unit).
#include "ualm_tp_samples.cuh"
#include "ualm_tp_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace ualm_tp {

#define TP_LANES TP_MAXLANES

#define TCK(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) { if (err) *err = std::string(#call) + ": " + cudaGetErrorString(e_); return UALM_ENOCUDA; } \
    } while (0)

template <class T>
struct Buf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct TpLane {
    bool have_batch = false, in_flight = false, collected = false;
    int B = 0, Nmax = 1, Mmax = 1, mode = 0;
    std::vector<int32_t> N, M;
    std::vector<int> slots;
    std::vector<AdmitDesc> ad;
    std::vector<GatherDesc> gd;
    long long tot_x = 0, tot_s = 0, tot_cxy = 0, tot_cyaw = 0;
    Buf<AdmitDesc> d_ad;
    Buf<GatherDesc> d_gd;
    Buf<double> d_x0, d_cxy, d_cyaw, d_xout, d_f, d_grad, d_hx, d_gx, d_sfx, d_scx, d_lam, d_mu, d_scin, d_sfin;
    Buf<long long> d_offs;
    Buf<ualm_result_t> d_res;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int launches = 0;
    double rho_eval = 1.0;
    bool has_lam = false, has_mu = false, has_scx = false, has_sfx = false;
};

struct TpEngine {
    int device = 0, precision = 32;
    bool f32() const { return precision == 32; }
    size_t esz() const { return precision == 32 ? 4 : 8; }
    cudaStream_t stream = nullptr;
    cudaEvent_t evA = nullptr, evB = nullptr;
    bool have_params = false, have_map = false;
    TpParams p;
    TpMap map;
    CUtensorMap tmap;
    bool have_tmap = false;
    int use_tma = 0;              // map corners of kb_kernel: 0 = direct gather through L1 (default: measured faster), 1 = TMA-staged tiles (UALM_TP_TMA=1)
    Buf<float4> cells;
    // pool
    TpPool E;
    int capacity = 0;
    std::vector<int> free_slots;
    int live = 0;                 // slots admitted and not yet freed
    // groups: every batch in flight is split into TP_SUBGROUPS groups (group = lane * TP_SUBGROUPS + k); each group has its own
    // stream, active list and round sequence, so ka of one group overlaps kb of another and a slow trajectory only holds up
    // the rounds of its own group
    cudaStream_t gstream[TP_NGROUPS] = {};
    cudaEvent_t gev[TP_NGROUPS] = {};
    int glive[TP_NGROUPS] = {};           // slots admitted and not yet freed, per group
    int gscale[TP_NGROUPS] = {};          // rounds that still have to run ks_kernel
    bool gcompact[TP_NGROUPS] = {};
    long long grounds[TP_NGROUPS] = {};
    int sg = 1;                           // subgroups per batch in use (UALM_TP_SUBGROUPS, <= TP_SUBGROUPS); 1 measured best: the batches in flight already are the groups
    int Nmax_live = 1, Mmax_live = 1;
    Buf<TpState> st;
    Buf<int> active, n_active, remaining;
    Buf<double> vec, cd, gw, kb_cost, lm_ys, lm_alpha, wf, wt;
    Buf<unsigned char> cr, gdc, gdt, dual, hs, hy, wway;
    Buf<long long> wf_off, wway_off, kprof;
    int *h_remaining = nullptr;   // pinned
    TpLane lanes[TP_LANES];
    long long rounds_total = 0;
    int chunk = 8;
    int ka_warps = 1;             // warps (= trajectories) per ka_kernel CTA: 1, 2 or 4 (UALM_TP_KA_WARPS); 1 measured best by ~1 %: no warp waits for a CTA mate
    bool prof = false;            // developer profile (UALM_TP_PROFILE=1): in-kernel phase cycle counters, printed at destroy
};

// The columns of A(1)^-1 a MINCO right-hand side can excite, for every piece count (ualm_tp_kernels.cuh: w_tables_kernel).  The LU
// factors they are solved with are scratch
static int build_tables(TpEngine *e, std::string *err)
{
    const int Pmax = TP_MMAX;
    std::vector<int> lu_off(Pmax + 1, 0);
    std::vector<long long> wf_off(Pmax + 1, 0), w_off(Pmax + 1, 0);
    size_t lu_tot = 0, wf_tot = 0, w_tot = 0;
    for (int P = 1; P <= Pmax; P++) {
        lu_off[P] = (int)(lu_tot + (size_t)TP_FPAD * TP_FW);
        lu_tot += (size_t)(6 * P + 2 * TP_FPAD) * TP_FW;
        wf_off[P] = (long long)wf_tot;
        wf_tot += (size_t)(P + 5) * 6 * P;
        w_off[P] = (long long)w_tot;
        w_tot += (size_t)std::max(P - 1, 0) * 6 * P;
    }
    Buf<double> lu;
    Buf<int> d_lu_off;
    TCK(lu.ensure(lu_tot));
    TCK(d_lu_off.ensure(Pmax + 1));
    TCK(e->wf_off.ensure(Pmax + 1));
    TCK(e->wway_off.ensure(Pmax + 1));
    TCK(e->wf.ensure(wf_tot));
    TCK(e->wt.ensure(w_tot));
    TCK(e->wway.ensure(w_tot * e->esz()));
    TCK(cudaMemcpyAsync(d_lu_off.p, lu_off.data(), sizeof(int) * (Pmax + 1), cudaMemcpyHostToDevice, e->stream));
    TCK(cudaMemcpyAsync(e->wf_off.p, wf_off.data(), sizeof(long long) * (Pmax + 1), cudaMemcpyHostToDevice, e->stream));
    TCK(cudaMemcpyAsync(e->wway_off.p, w_off.data(), sizeof(long long) * (Pmax + 1), cudaMemcpyHostToDevice, e->stream));
    lu_tables_kernel<<<(Pmax + 31) / 32, 32, 0, e->stream>>>(lu.p, d_lu_off.p, Pmax);
    TCK(cudaGetLastError());
    const int tpb = 32, gx = (Pmax + 5 + tpb - 1) / tpb;
    dim3 grid(gx, Pmax);
    if (e->f32()) w_tables_kernel<float><<<grid, tpb, 0, e->stream>>>(lu.p, d_lu_off.p, e->wf.p, e->wf_off.p, e->wt.p, (float *)e->wway.p, e->wway_off.p, Pmax);
    else w_tables_kernel<double><<<grid, tpb, 0, e->stream>>>(lu.p, d_lu_off.p, e->wf.p, e->wf_off.p, e->wt.p, (double *)e->wway.p, e->wway_off.p, Pmax);
    TCK(cudaGetLastError());
    TCK(cudaStreamSynchronize(e->stream));
    lu.release();
    d_lu_off.release();
    return UALM_OK;
}

int tp_create(TpEngine **out, int device, int precision, std::string *err)
{
    if (precision != 32 && precision != 65) { if (err) *err = "throughput engine: precision must be 32 or 65"; return UALM_EINVAL; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0) { if (err) *err = std::string("no CUDA device: ") + cudaGetErrorString(ce); return UALM_ENOCUDA; }
    if (device < 0 || device >= ndev) { if (err) *err = "device ordinal out of range"; return UALM_EINVAL; }
    TCK(cudaSetDevice(device));
    TpEngine *e = new TpEngine();
    e->device = device; e->precision = precision;
    if (const char *s = getenv("UALM_TP_TMA")) e->use_tma = atoi(s) ? 1 : 0;
    if (const char *s = getenv("UALM_TP_NOTMA")) { if (atoi(s)) e->use_tma = 0; }
    if (const char *s = getenv("UALM_TP_CHUNK")) e->chunk = std::max(1, atoi(s));
    if (const char *s = getenv("UALM_TP_KA_WARPS")) { const int w = atoi(s); if (w == 1 || w == 2 || w == 4) e->ka_warps = w; }
    if (const char *s = getenv("UALM_TP_PROFILE")) e->prof = atoi(s) != 0;
    if (const char *s = getenv("UALM_TP_SUBGROUPS")) e->sg = std::min(TP_SUBGROUPS, std::max(1, atoi(s)));
    memset(&e->E, 0, sizeof(e->E));
    auto failed = [&](int rc) { tp_destroy(e); return rc; };
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&e->evA) != cudaSuccess ||
        cudaEventCreate(&e->evB) != cudaSuccess || cudaMallocHost(&e->h_remaining, sizeof(int) * TP_MAX_TICKETS) != cudaSuccess) {
        if (err) *err = "throughput engine: stream / event / pinned allocation failed";
        return failed(UALM_ENOCUDA);
    }
    for (auto &l : e->lanes)
        if (cudaEventCreate(&l.ev0) != cudaSuccess || cudaEventCreate(&l.ev1) != cudaSuccess) { if (err) *err = "event creation failed"; return failed(UALM_ENOCUDA); }
    for (int g = 0; g < TP_NGROUPS; g++)
        if (cudaStreamCreateWithFlags(&e->gstream[g], cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&e->gev[g], cudaEventDisableTiming) != cudaSuccess) {
            if (err) *err = "group stream creation failed";
            return failed(UALM_ENOCUDA);
        }
    if (e->remaining.ensure(TP_MAX_TICKETS) != cudaSuccess || e->n_active.ensure(TP_NGROUPS) != cudaSuccess) { if (err) *err = "allocation failed"; return failed(UALM_ENOCUDA); }
    cudaMemsetAsync(e->remaining.p, 0, sizeof(int) * TP_MAX_TICKETS, e->stream);
    cudaMemsetAsync(e->n_active.p, 0, sizeof(int) * TP_NGROUPS, e->stream);
    int rc = build_tables(e, err);
    if (rc) return failed(rc);
    *out = e;
    return UALM_OK;
}

void tp_destroy(TpEngine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    if (e->prof && e->kprof.p) {
        long long h[32];
        cudaMemcpy(h, e->kprof.p, sizeof(h), cudaMemcpyDeviceToHost);
        const char *nm[KP_N] = {"finish:pre", "finish:gradient-tables", "finish:post", "advance:line-search", "advance:post-ls", "two-loop", "alm-update", "forward:pre",
                                "forward:coefficient-tables", "forward:post", "(unused)", "warps"};
        long long tot = 0;
        for (int q = 0; q < KP_WARPS; q++) tot += h[q];
        fprintf(stderr, "[ualm-tp] ka phases (SM cycles per warp-round, %lld warp-rounds):", h[KP_WARPS]);
        for (int q = 0; q < KP_WARPS; q++) fprintf(stderr, " %s %.0f (%.0f%%);", nm[q], (double)h[q] / std::max(1ll, h[KP_WARPS]), 100.0 * h[q] / std::max(1ll, tot));
        fprintf(stderr, " total %.0f\n", (double)tot / std::max(1ll, h[KP_WARPS]));
        const char *nb[7] = {"setup", "spline+stencil", "tile issue+wait", "terrain+penalties", "barrier", "reductions", "final"};
        long long tb = 0;
        for (int q = 0; q < 7; q++) tb += h[16 + q];
        fprintf(stderr, "[ualm-tp] kb phases (SM cycles of thread 0 per CTA, %lld CTAs):", h[16 + 7]);
        for (int q = 0; q < 7; q++) fprintf(stderr, " %s %.0f (%.0f%%);", nb[q], (double)h[16 + q] / std::max(1ll, h[16 + 7]), 100.0 * h[16 + q] / std::max(1ll, tb));
        fprintf(stderr, " total %.0f\n", (double)tb / std::max(1ll, h[16 + 7]));
        e->kprof.release();
    }
    for (int g = 0; g < TP_NGROUPS; g++) { if (e->gstream[g]) cudaStreamDestroy(e->gstream[g]); if (e->gev[g]) cudaEventDestroy(e->gev[g]); }
    for (auto &l : e->lanes) {
        l.d_ad.release(); l.d_gd.release(); l.d_x0.release(); l.d_cxy.release(); l.d_cyaw.release(); l.d_xout.release(); l.d_f.release(); l.d_grad.release();
        l.d_hx.release(); l.d_gx.release(); l.d_sfx.release(); l.d_scx.release(); l.d_lam.release(); l.d_mu.release(); l.d_scin.release(); l.d_sfin.release();
        l.d_offs.release(); l.d_res.release();
        if (l.ev0) cudaEventDestroy(l.ev0);
        if (l.ev1) cudaEventDestroy(l.ev1);
    }
    e->cells.release(); e->st.release(); e->active.release(); e->n_active.release(); e->remaining.release(); e->vec.release(); e->cd.release(); e->gw.release();
    e->kb_cost.release(); e->lm_ys.release(); e->lm_alpha.release(); e->cr.release(); e->gdc.release(); e->gdt.release(); e->dual.release();
    e->hs.release(); e->hy.release(); e->wway.release(); e->wf.release(); e->wt.release(); e->wf_off.release(); e->wway_off.release();
    if (e->h_remaining) cudaFreeHost(e->h_remaining);
    if (e->evA) cudaEventDestroy(e->evA);
    if (e->evB) cudaEventDestroy(e->evB);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static bool any_in_flight(TpEngine *e)
{
    for (auto &l : e->lanes) if (l.in_flight) return true;
    return false;
}

int tp_set_params(TpEngine *e, const ualm_params_t *p, std::string *err)
{
    if (p->int_K < 1 || p->int_K > TP_KB_THREADS - 1) { if (err) *err = "int_K out of range [1,127] on the throughput path"; return UALM_ELIMIT; }
    if (p->mem_size < 1 || p->mem_size > 1024) { if (err) *err = "mem_size out of range [1,1024]"; return UALM_ELIMIT; }
    if (p->past < 0 || p->past > 16) { if (err) *err = "past out of range [0,16]"; return UALM_ELIMIT; }
    if (any_in_flight(e)) { if (err) *err = "ualm_set_params while a batch is in flight"; return UALM_ESTATE; }
    TpParams &d = e->p;
    d.rho_T = p->rho_T; d.rho_ter = p->rho_ter; d.max_vel = p->max_vel; d.max_acc_lon = p->max_acc_lon; d.max_acc_lat = p->max_acc_lat;
    d.max_kap = p->max_kap; d.min_cxi = p->min_cxi; d.max_sig = p->max_sig; d.use_scaling = p->use_scaling; d.rho = p->rho;
    d.beta = p->beta; d.gamma = p->gamma; d.epsilon_con = p->epsilon_con; d.max_iter = p->max_iter; d.g_epsilon = p->g_epsilon;
    d.min_step = p->min_step; d.delta = p->delta; d.inner_max_iter = (int)p->inner_max_iter; d.mem_size = p->mem_size; d.past = p->past;
    d.int_K = p->int_K; d.gravity = p->gravity;
    e->have_params = true;
    // the pool strides depend on int_K and mem_size: it is rebuilt at the next admission; uploaded batches stay valid (they hold
    // problem data only), but nothing may be resident in the pool
    e->capacity = 0; e->free_slots.clear(); e->live = 0;
    for (int g = 0; g < TP_NGROUPS; g++) { e->glive[g] = 0; e->gscale[g] = 0; e->gcompact[g] = false; }
    for (auto &l : e->lanes) { l.collected = false; }
    return UALM_OK;
}

int tp_set_map(TpEngine *e, const ualm_map_geom_t *g, const float *cells, std::string *err)
{
    if (any_in_flight(e)) { if (err) *err = "ualm_set_map while a batch is in flight"; return UALM_ESTATE; }
    TCK(cudaSetDevice(e->device));
    TCK(cudaStreamSynchronize(e->stream));
    const size_t ncell = (size_t)g->voxel_num[0] * g->voxel_num[1] * g->voxel_num[2];
    TCK(e->cells.ensure(ncell));
    TCK(cudaMemcpy(e->cells.p, cells, ncell * sizeof(float4), cudaMemcpyHostToDevice));
    TpMap &m = e->map;
    m.cells = e->cells.p;
    for (int k = 0; k < 3; k++) { m.vn[k] = g->voxel_num[k]; m.origin[k] = g->origin[k]; m.maxb[k] = g->max_boundary[k]; }
    m.xy_res = g->xy_resolution; m.yaw_res = g->yaw_resolution; m.xy_inv = 1.0 / g->xy_resolution; m.yaw_inv = 1.0 / g->yaw_resolution;
    // TMA descriptor of the grid: a 3-D float tensor {4 * Yaw, Y, X} (innermost first), box = 8 yaw layers x 8 x 8 cells
    e->have_tmap = false;
    memset(&e->tmap, 0, sizeof(e->tmap));
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn && qres == cudaDriverEntryPointSuccess) {
        const cuuint64_t dims[3] = {(cuuint64_t)4 * g->voxel_num[2], (cuuint64_t)g->voxel_num[1], (cuuint64_t)g->voxel_num[0]};
        const cuuint64_t strides[2] = {(cuuint64_t)16 * g->voxel_num[2], (cuuint64_t)16 * g->voxel_num[2] * g->voxel_num[1]};
        const cuuint32_t box[3] = {4 * TP_TILE, TP_TILE, TP_TILE};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = ((EncodeFn)fn)(&e->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)e->cells.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        e->have_tmap = (r == CUDA_SUCCESS) && g->voxel_num[2] >= TP_TILE && g->voxel_num[1] >= TP_TILE && g->voxel_num[0] >= TP_TILE;
    }
    if (getenv("UALM_DEBUG")) fprintf(stderr, "[ualm-tp] tensor map %s, TMA staging %s\n", e->have_tmap ? "encoded" : "unavailable", (e->have_tmap && e->use_tma) ? "on" : "off");
    e->have_map = true;
    return UALM_OK;
}

// (re)allocate the pool for `cap` slots with the current parameters; only when nothing is resident
static int pool_alloc(TpEngine *e, int cap, std::string *err)
{
    const int K = e->p.int_K, m = e->p.mem_size;
    const size_t Smax = (size_t)TP_NMAX * (K + 1), es = e->esz();
    TCK(e->st.ensure(cap)); TCK(e->active.ensure((size_t)cap * TP_NGROUPS));
    TCK(e->vec.ensure((size_t)cap * 5 * TP_NVAR)); TCK(e->cd.ensure((size_t)cap * TP_CSTRIDE)); TCK(e->gw.ensure((size_t)cap * TP_CSTRIDE));
    TCK(e->kb_cost.ensure(cap)); TCK(e->lm_ys.ensure((size_t)cap * m)); TCK(e->lm_alpha.ensure((size_t)cap * m));
    if (e->f32()) TCK(e->cr.ensure((size_t)cap * TP_CSTRIDE * es));
    TCK(e->gdc.ensure((size_t)cap * TP_CSTRIDE * es)); TCK(e->gdt.ensure((size_t)cap * TP_TSTRIDE * es));
    TCK(e->dual.ensure((size_t)cap * TP_NDUAL * Smax * es));
    TCK(e->hs.ensure((size_t)cap * m * TP_NVAR * es)); TCK(e->hy.ensure((size_t)cap * m * TP_NVAR * es));
    TCK(cudaMemsetAsync(e->st.p, 0, sizeof(TpState) * cap, e->stream));     // every slot PH_FREE
    TCK(cudaMemsetAsync(e->n_active.p, 0, sizeof(int) * TP_NGROUPS, e->stream));
    TCK(cudaStreamSynchronize(e->stream));
    TpPool &E = e->E;
    E.capacity = cap; E.m = m; E.K = K; E.Smax = (int)Smax; E.use_tma = e->use_tma;
    E.st = e->st.p; E.active = e->active.p; E.n_active = e->n_active.p; E.remaining = e->remaining.p;
    E.vec = e->vec.p; E.cd = e->cd.p; E.gw = e->gw.p; E.cr = e->f32() ? (void *)e->cr.p : (void *)e->cd.p; E.gdc = e->gdc.p; E.gdt = e->gdt.p;
    E.kb_cost = e->kb_cost.p; E.dual = e->dual.p; E.hs = e->hs.p; E.hy = e->hy.p; E.lm_ys = e->lm_ys.p; E.lm_alpha = e->lm_alpha.p;
    E.wf = e->wf.p; E.wf_off = e->wf_off.p; E.wt = e->wt.p; E.wway = e->wway.p; E.wway_off = e->wway_off.p;
    E.prof = nullptr;
    if (e->prof) { TCK(e->kprof.ensure(32)); TCK(cudaMemsetAsync(e->kprof.p, 0, 32 * sizeof(long long), e->stream)); E.prof = e->kprof.p; }
    e->capacity = cap;
    e->free_slots.clear();
    for (int s = cap - 1; s >= 0; s--) e->free_slots.push_back(s);
    e->live = 0;
    for (int g = 0; g < TP_NGROUPS; g++) e->glive[g] = 0;
    if (getenv("UALM_DEBUG")) fprintf(stderr, "[ualm-tp] pool of %d slots allocated (%.2f MB per slot)\n", cap,
                                      (double)((size_t)5 * TP_NVAR * 8 + 3 * TP_CSTRIDE * 8 + TP_NDUAL * Smax * es + 2 * (size_t)m * TP_NVAR * es) / 1e6);
    return UALM_OK;
}

int tp_upload(TpEngine *e, int lane, int B, const int32_t *N, const int32_t *M, const double *bnd, const double *total_time, const double *inner_xy,
              const double *inner_yaw, std::string *err)
{
    if (lane < 0 || lane >= TP_LANES) { if (err) *err = "lane out of range"; return UALM_EINVAL; }
    if (!e->have_params) { if (err) *err = "ualm_set_params must be called before ualm_upload"; return UALM_ESTATE; }
    TpLane &l = e->lanes[lane];
    if (l.in_flight) { if (err) *err = "ualm_upload into a lane whose batch is still in flight"; return UALM_ESTATE; }
    long long need_xy = 0, need_yaw = 0;
    for (int b = 0; b < B; b++) {
        if (N[b] < 1 || M[b] < 1) { if (err) *err = "piece counts must be >= 1"; return UALM_EINVAL; }
        if (!(total_time[b] > 0.0) || !std::isfinite(total_time[b])) { if (err) *err = "total_time must be finite and > 0"; return UALM_EINVAL; }
        need_xy += 2LL * (N[b] - 1); need_yaw += M[b] - 1;
    }
    if ((need_xy > 0 && !inner_xy) || (need_yaw > 0 && !inner_yaw)) { if (err) *err = "inner waypoint arrays are NULL but N > 1 or M > 1"; return UALM_EINVAL; }
    TCK(cudaSetDevice(e->device));
    const int K = e->p.int_K;
    l.have_batch = false; l.collected = false;
    l.B = B; l.N.assign(N, N + B); l.M.assign(M, M + B); l.ad.assign(B, AdmitDesc()); l.gd.assign(B, GatherDesc());
    l.Nmax = l.Mmax = 1;
    std::vector<double> x0;
    std::vector<long long> offs(B + 1, 0);
    long long ox = 0, os = 0, ocx = 0, ocy = 0, oixy = 0, oiyw = 0;
    for (int b = 0; b < B; b++) {
        AdmitDesc &a = l.ad[b];
        GatherDesc &g = l.gd[b];
        const bool skip = N[b] > TP_NMAX || M[b] > TP_MMAX;
        a.slot = -1; a.N = N[b]; a.M = M[b]; a.ticket = lane; a.index = b; a.mode = skip ? -1 : 0; a.off_x = ox; a.total_time = total_time[b];
        for (int k = 0; k < 18; k++) a.bnd[k] = bnd[(size_t)b * 18 + k];
        g.slot = -1; g.N = N[b]; g.M = M[b]; g.off_cxy = ocx; g.off_cyaw = ocy; g.off_x = ox; g.off_s = os;
        offs[b] = os;
        const long long nxy = 2LL * (N[b] - 1), nyw = M[b] - 1;
        ocx += 12LL * N[b]; ocy += 6LL * M[b];
        if (!skip) {
            const double T = total_time[b];
            x0.push_back(T > 1.0 ? (std::sqrt(2.0 * T - 1.0) - 1.0) : (1.0 - std::sqrt(2.0 / T - 1.0)));   // logC2, alm_traj_opt.h:238-241
            for (long long q = 0; q < nxy; q++) x0.push_back(inner_xy[oixy + q]);
            for (long long q = 0; q < nyw; q++) x0.push_back(inner_yaw[oiyw + q]);
            ox += 1 + nxy + nyw; os += (long long)N[b] * (K + 1);
            l.Nmax = std::max(l.Nmax, (int)N[b]); l.Mmax = std::max(l.Mmax, (int)M[b]);
        }
        oixy += nxy; oiyw += nyw;
    }
    offs[B] = os;
    l.tot_x = ox; l.tot_s = os; l.tot_cxy = ocx; l.tot_cyaw = ocy;
    TCK(l.d_ad.ensure(B)); TCK(l.d_gd.ensure(B)); TCK(l.d_x0.ensure(ox)); TCK(l.d_cxy.ensure(ocx)); TCK(l.d_cyaw.ensure(ocy)); TCK(l.d_res.ensure(B));
    TCK(l.d_offs.ensure(B + 1));
    if (B > 0) {
        if (ox > 0) TCK(cudaMemcpyAsync(l.d_x0.p, x0.data(), sizeof(double) * ox, cudaMemcpyHostToDevice, e->stream));
        TCK(cudaMemcpyAsync(l.d_offs.p, offs.data(), sizeof(long long) * (B + 1), cudaMemcpyHostToDevice, e->stream));
        TCK(cudaStreamSynchronize(e->stream));
    }
    l.have_batch = true; l.mode = 0;
    return UALM_OK;
}

static int admit_impl(TpEngine *e, int lane, int mode, bool single_group, std::string *err)
{
    TpLane &l = e->lanes[lane];
    if (!e->have_params || !e->have_map || !l.have_batch) { if (err) *err = "set_params, set_map and upload must precede the solve"; return UALM_ESTATE; }
    if (l.in_flight) { if (err) *err = "lane already in flight"; return UALM_ESTATE; }
    TCK(cudaSetDevice(e->device));
    int need = 0;
    for (int b = 0; b < l.B; b++) if (!(l.N[b] > TP_NMAX || l.M[b] > TP_MMAX)) need++;
    // the pool grows only while nothing is resident; it is sized for TP_LANES batches of this size in flight (fewer when that
    // would take more than ~64 GB of HBM)
    int mult = TP_LANES;
    {
        const double per_slot = 5.0 * TP_NVAR * 8 + 3.0 * TP_CSTRIDE * 8 + (double)TP_NDUAL * TP_NMAX * (e->p.int_K + 1) * e->esz() + 2.0 * e->p.mem_size * TP_NVAR * e->esz();
        while (mult > 1 && per_slot * mult * need > 64e9) mult /= 2;
    }
    if (e->live == 0 && e->capacity < mult * need) {
        int cap = std::max(1024, e->capacity);
        while (cap < mult * need) cap *= 2;
        if (const char *s = getenv("UALM_TP_CAPACITY")) cap = std::max(cap, atoi(s));
        int rc = pool_alloc(e, cap, err);
        if (rc) return rc;
    }
    if ((int)e->free_slots.size() < need) { if (err) *err = "throughput pool exhausted: collect a batch in flight first (the pool grows only when idle)"; return UALM_ELIMIT; }
    l.slots.clear();
    std::vector<ualm_result_t> res0(l.B);
    int k = 0;
    for (int b = 0; b < l.B; b++) {
        memset(&res0[b], 0, sizeof(ualm_result_t));
        const bool skip = l.N[b] > TP_NMAX || l.M[b] > TP_MMAX;
        if (skip) { l.ad[b].slot = -1; l.gd[b].slot = -1; l.ad[b].mode = -1; res0[b].ret_code = UALM_ELIMIT; continue; }
        const int s = e->free_slots.back();
        e->free_slots.pop_back();
        l.slots.push_back(s);
        const int g = lane * TP_SUBGROUPS + (single_group ? 0 : k % e->sg);
        l.ad[b].slot = s; l.ad[b].mode = mode; l.ad[b].group = g; l.gd[b].slot = s;
        e->glive[g]++;
        k++;
    }
    e->live += (int)l.slots.size();
    e->Nmax_live = std::max(e->Nmax_live, l.Nmax); e->Mmax_live = std::max(e->Mmax_live, l.Mmax);
    // compact descriptors of the admitted problems only (skipped ones keep their UALM_ELIMIT record)
    std::vector<AdmitDesc> ad;
    std::vector<long long> offs;
    {
        long long os = 0;
        for (int b = 0; b < l.B; b++) if (l.ad[b].slot >= 0) { ad.push_back(l.ad[b]); offs.push_back(os); os += (long long)l.N[b] * (e->p.int_K + 1); }
    }
    const int BA = (int)ad.size();
    TCK(cudaEventRecord(l.ev0, e->stream));
    l.launches = 0;
    TCK(cudaMemcpyAsync(l.d_res.p, res0.data(), sizeof(ualm_result_t) * l.B, cudaMemcpyHostToDevice, e->stream));
    if (BA < l.B) {
        TCK(cudaMemsetAsync(l.d_cxy.p, 0, sizeof(double) * l.tot_cxy, e->stream));
        TCK(cudaMemsetAsync(l.d_cyaw.p, 0, sizeof(double) * l.tot_cyaw, e->stream));
    }
    if (BA > 0) {
        TCK(cudaMemcpyAsync(l.d_ad.p, ad.data(), sizeof(AdmitDesc) * BA, cudaMemcpyHostToDevice, e->stream));
        TCK(cudaMemcpyAsync(l.d_offs.p, offs.data(), sizeof(long long) * BA, cudaMemcpyHostToDevice, e->stream));
        TCK(cudaMemcpyAsync(l.d_gd.p, l.gd.data(), sizeof(GatherDesc) * l.B, cudaMemcpyHostToDevice, e->stream));
        e->h_remaining[lane] = BA;
        TCK(cudaMemcpyAsync(e->remaining.p + lane, e->h_remaining + lane, sizeof(int), cudaMemcpyHostToDevice, e->stream));
        const double *lam = l.has_lam ? l.d_lam.p : nullptr, *mu = l.has_mu ? l.d_mu.p : nullptr, *scx = l.has_scx ? l.d_scin.p : nullptr,
                     *sfx = l.has_sfx ? l.d_sfin.p : nullptr;
        if (e->f32()) admit_kernel<float><<<BA, 128, 0, e->stream>>>(e->E, l.d_ad.p, l.d_x0.p, BA, e->p.use_scaling, e->p.int_K, lam, mu, scx, sfx, l.rho_eval, l.d_offs.p);
        else admit_kernel<double><<<BA, 128, 0, e->stream>>>(e->E, l.d_ad.p, l.d_x0.p, BA, e->p.use_scaling, e->p.int_K, lam, mu, scx, sfx, l.rho_eval, l.d_offs.p);
        TCK(cudaGetLastError());
        // host-synchronised: ad / offs / res0 are stack-local staging vectors, and the lane's group streams start after the admission
        TCK(cudaStreamSynchronize(e->stream));
        l.launches++;
        for (int q = 0; q < TP_SUBGROUPS; q++) {
            const int g = lane * TP_SUBGROUPS + q;
            if (e->glive[g] > 0) { e->gcompact[g] = true; if ((mode == 0 && e->p.use_scaling) || mode == 2) e->gscale[g] = std::max(e->gscale[g], 1); }
        }
    } else {
        e->h_remaining[lane] = 0;
        TCK(cudaStreamSynchronize(e->stream));
    }
    l.in_flight = true; l.collected = false; l.mode = mode;
    return UALM_OK;
}

int tp_admit(TpEngine *e, int lane, std::string *err)
{
    if (lane < 0 || lane >= TP_LANES) { if (err) *err = "lane out of range"; return UALM_EINVAL; }
    return admit_impl(e, lane, 0, false, err);
}

static size_t kb_smem_bytes(TpEngine *e, bool tma)
{
    const size_t es = e->esz();
    const size_t arrays = 2 * (size_t)(TP_KB_THREADS / 32) * (TP_MAXPPC * 13 + TP_YCAP * 7) * es;      // two ChunkPart<R> tables
    const size_t coef = ((size_t)12 * e->Nmax_live + (size_t)6 * e->Mmax_live * 2 + e->Mmax_live) * es;
    return (tma ? (size_t)TP_MAXPPC * TP_TILE_BYTES : 0) + ((arrays + 15) & ~(size_t)15) + coef + 64;
}
static size_t ks_smem_bytes(TpEngine *e)
{
    const size_t es = e->esz();
    const size_t arrays = (size_t)(TP_KB_THREADS / 32) * (TP_MAXPPC * 13 + TP_YCAP * 7) * es;      // ChunkPart<R>
    const size_t coef = ((size_t)12 * e->Nmax_live * 3 + (size_t)6 * e->Mmax_live * 3 + e->Nmax_live + e->Mmax_live) * es;
    return ((arrays + 15) & ~(size_t)15) + coef + 64;
}
// ka_kernel's shared memory per warp: the C^-1 dcost/dc vector (12 N + 6 M doubles of the largest live problem), aliased by the
// right-hand sides of the forward pass and by the two-loop's history ring (TP_HRING slots x {s, y} x n elements)
static void ka_layout(TpEngine *e)
{
    const int nmax = 1 + 2 * (e->Nmax_live - 1) + (e->Mmax_live - 1);
    const int hstride = (int)(((size_t)nmax * e->esz() + 15) / 16 * 16 / e->esz());
    const size_t col = ((size_t)12 * e->Nmax_live + 6 * e->Mmax_live) * 8, hist = (((size_t)TP_HRING * 2 * hstride * e->esz()) + 15) & ~(size_t)15;
    e->E.ka_hist_bytes = (int)hist;                                             // behind the ring: 1 / ys and alpha of the two-loop (2 m doubles)
    e->E.ka_col_bytes = (int)((std::max(col, hist + 2 * (size_t)e->p.mem_size * 8) + 15) & ~(size_t)15);
    e->E.ka_hist_stride = hstride;
}
static size_t ka_smem_bytes(TpEngine *e) { return (size_t)e->ka_warps * e->E.ka_col_bytes; }

// one round of group g on its stream: ka -> [ks] -> kb
template <class R>
static int launch_round(TpEngine *e, int g, bool with_ks, bool tma, std::string *err)
{
    const int upper = e->glive[g];
    cudaStream_t st = e->gstream[g];
    ka_kernel<R><<<(upper + e->ka_warps - 1) / e->ka_warps, 32 * e->ka_warps, ka_smem_bytes(e), st>>>(e->E, e->p, g);
    if (with_ks) ks_kernel<R><<<upper, TP_KB_THREADS, ks_smem_bytes(e), st>>>(e->E, e->p, e->map, g);
    if (tma) kb_kernel<R, true><<<upper, TP_KB_THREADS, kb_smem_bytes(e, true), st>>>(e->E, e->p, e->map, e->tmap, g);
    else kb_kernel<R, false><<<upper, TP_KB_THREADS, kb_smem_bytes(e, false), st>>>(e->E, e->p, e->map, e->tmap, g);
    TCK(cudaGetLastError());
    return UALM_OK;
}

template <class R>
static int set_attrs(TpEngine *e, std::string *err)
{
    static size_t done_kb[2][2] = {{0, 0}, {0, 0}}, done_ks[2] = {0, 0}, done_ka[2] = {0, 0};
    const int pi = sizeof(R) == 4 ? 0 : 1;
    ka_layout(e);
    const size_t a = kb_smem_bytes(e, true), b = kb_smem_bytes(e, false), c = ks_smem_bytes(e), d = ka_smem_bytes(e);
    if (a > 227 * 1024 || c > 227 * 1024 || d > 227 * 1024) { if (err) *err = "problem too large for shared memory"; return UALM_ELIMIT; }
    if (d > done_ka[pi]) { TCK(cudaFuncSetAttribute(ka_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d)); done_ka[pi] = d; }
    if (a > done_kb[pi][1]) { TCK(cudaFuncSetAttribute(kb_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a)); done_kb[pi][1] = a; }
    if (b > done_kb[pi][0]) { TCK(cudaFuncSetAttribute(kb_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b)); done_kb[pi][0] = b; }
    if (c > done_ks[pi]) { TCK(cudaFuncSetAttribute(ks_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c)); done_ks[pi] = c; }
    return UALM_OK;
}

// `nr` rounds of every live group, interleaved over the groups' streams
static int run_rounds(TpEngine *e, int nr, std::string *err)
{
    if (e->live <= 0) return UALM_OK;
    int rc = e->f32() ? set_attrs<float>(e, err) : set_attrs<double>(e, err);
    if (rc) return rc;
    const bool tma = e->have_tmap && e->use_tma;
    for (int r = 0; r < nr; r++) {
        for (int g = 0; g < TP_NGROUPS; g++) {
            if (e->glive[g] <= 0) continue;
            if (e->gcompact[g] || (e->grounds[g] % e->chunk) == 0) {
                compact_kernel<<<1, 1024, 0, e->gstream[g]>>>(e->E, e->capacity, g);
                e->gcompact[g] = false;
            }
            const bool with_ks = e->gscale[g] > 0;
            rc = e->f32() ? launch_round<float>(e, g, with_ks, tma, err) : launch_round<double>(e, g, with_ks, tma, err);
            if (rc) return rc;
            if (e->gscale[g] > 0) e->gscale[g]--;
            e->grounds[g]++;
            e->lanes[g / TP_SUBGROUPS].launches += 2 + (with_ks ? 1 : 0);
        }
    }
    return UALM_OK;
}

int tp_collect(TpEngine *e, int lane, std::string *err)
{
    if (lane < 0 || lane >= TP_LANES) { if (err) *err = "lane out of range"; return UALM_EINVAL; }
    TpLane &l = e->lanes[lane];
    if (l.collected) return UALM_OK;
    if (!l.in_flight) { if (err) *err = "nothing in flight on this lane"; return UALM_ESTATE; }
    TCK(cudaSetDevice(e->device));
    long long guard = 0;
    while (true) {
        // the lane's own groups are the pace setters: wait for what was queued for them, then look at the lane's counter
        for (int q = 0; q < TP_SUBGROUPS; q++) TCK(cudaStreamSynchronize(e->gstream[lane * TP_SUBGROUPS + q]));
        TCK(cudaMemcpyAsync(e->h_remaining, e->remaining.p, sizeof(int) * TP_LANES, cudaMemcpyDeviceToHost, e->stream));
        TCK(cudaStreamSynchronize(e->stream));
        if (e->h_remaining[lane] <= 0) break;
        int rc = run_rounds(e, e->chunk, err);        // every live group advances, not only this lane's
        if (rc) return rc;
        if ((guard += e->chunk) > 400000) { if (err) *err = "throughput engine: a trajectory did not terminate"; return UALM_ENOCUDA; }
    }
    // gather the batch's results into the lane's packed device outputs, free its slots
    const bool ev = l.mode != 0;
    if (l.B > 0 && !l.slots.empty()) {
        if (ev) {
            TCK(l.d_f.ensure(l.B)); TCK(l.d_grad.ensure(l.tot_x)); TCK(l.d_hx.ensure(l.tot_s)); TCK(l.d_gx.ensure(6 * l.tot_s)); TCK(l.d_sfx.ensure(l.B));
            TCK(l.d_scx.ensure(7 * l.tot_s));
        }
        TCK(l.d_xout.ensure(l.tot_x));
#define GATHER(R)                                                                                                                                          \
        gather_kernel<R><<<l.B, 128, 0, e->stream>>>(e->E, l.d_gd.p, l.B, l.d_res.p, l.d_cxy.p, l.d_cyaw.p, l.d_xout.p, ev ? l.d_f.p : nullptr,            \
                                                     ev ? l.d_grad.p : nullptr, ev ? l.d_hx.p : nullptr, ev ? l.d_gx.p : nullptr, ev ? l.d_sfx.p : nullptr,    \
                                                     ev ? l.d_scx.p : nullptr)
        if (e->f32()) GATHER(float); else GATHER(double);
#undef GATHER
        TCK(cudaGetLastError());
        free_kernel<<<(l.B + 127) / 128, 128, 0, e->stream>>>(e->E, l.d_gd.p, l.B);
        TCK(cudaGetLastError());
        l.launches += 2;
    }
    TCK(cudaEventRecord(l.ev1, e->stream));
    TCK(cudaStreamSynchronize(e->stream));
    for (int s : l.slots) e->free_slots.push_back(s);
    e->live -= (int)l.slots.size();
    l.slots.clear();
    for (int q = 0; q < TP_SUBGROUPS; q++) { const int g = lane * TP_SUBGROUPS + q; e->glive[g] = 0; e->gscale[g] = 0; e->gcompact[g] = false; }
    l.in_flight = false; l.collected = true;
    return UALM_OK;
}

int tp_download(TpEngine *e, int lane, ualm_result_t *results, double *c_xy, double *c_yaw, std::string *err)
{
    TpLane &l = e->lanes[lane];
    if (!l.collected) { int rc = tp_collect(e, lane, err); if (rc) return rc; }
    TCK(cudaSetDevice(e->device));
    if (l.B > 0) {
        if (results) TCK(cudaMemcpyAsync(results, l.d_res.p, sizeof(ualm_result_t) * l.B, cudaMemcpyDeviceToHost, e->stream));
        if (c_xy) TCK(cudaMemcpyAsync(c_xy, l.d_cxy.p, sizeof(double) * l.tot_cxy, cudaMemcpyDeviceToHost, e->stream));
        if (c_yaw) TCK(cudaMemcpyAsync(c_yaw, l.d_cyaw.p, sizeof(double) * l.tot_cyaw, cudaMemcpyDeviceToHost, e->stream));
    }
    TCK(cudaStreamSynchronize(e->stream));
    return UALM_OK;
}

__global__ void tp_pack_records_kernel(const ualm_result_t *res, const GatherDesc *gd, const double *c_xy, const double *c_yaw, int B, double *rec, int stride)
{
    const int b = blockIdx.x;
    if (b >= B) return;
    const GatherDesc g = gd[b];
    double *o = rec + (size_t)b * stride;
    const ualm_result_t r = res[b];
    if (threadIdx.x == 0) {
        o[0] = r.ret_code; o[1] = r.outer_iters; o[2] = r.n_evals; o[3] = r.n_lbfgs_iters; o[4] = r.inner_cost; o[5] = r.jerk_cost;
        o[6] = r.total_T; o[7] = r.res_h; o[8] = r.res_g; o[9] = g.N; o[10] = g.M; o[11] = 0.0;
    }
    const int ncx = 12 * g.N, ncy = 6 * g.M;
    for (int q = threadIdx.x; q < stride - 12; q += blockDim.x) {
        double v = 0.0;
        if (q < ncx) v = c_xy[g.off_cxy + q];
        else if (q < ncx + ncy) v = c_yaw[g.off_cyaw + q - ncx];
        o[12 + q] = v;
    }
}

int tp_pack_records(TpEngine *e, int lane, double *d_records, int stride, std::string *err)
{
    TpLane &l = e->lanes[lane];
    if (!l.collected) { int rc = tp_collect(e, lane, err); if (rc) return rc; }
    if (stride < 12 + 12 * l.Nmax + 6 * l.Mmax) { if (err) *err = "record stride too small"; return UALM_EINVAL; }
    if (l.B > 0) {
        tp_pack_records_kernel<<<l.B, 128, 0, e->stream>>>(l.d_res.p, l.d_gd.p, l.d_cxy.p, l.d_cyaw.p, l.B, d_records, stride);
        TCK(cudaGetLastError());
    }
    TCK(cudaStreamSynchronize(e->stream));
    return UALM_OK;
}

bool tp_lane_in_flight(TpEngine *e, int lane) { return e->lanes[lane].in_flight; }
bool tp_lane_has_batch(TpEngine *e, int lane) { return e->lanes[lane].have_batch; }
bool tp_lane_collected(TpEngine *e, int lane) { return e->lanes[lane].collected; }

int tp_last_solve(TpEngine *e, int lane, float *ms, int *launches)
{
    TpLane &l = e->lanes[lane];
    if (!l.collected) return UALM_ESTATE;
    float t = 0.f;
    cudaEventElapsedTime(&t, l.ev0, l.ev1);
    if (ms) *ms = t;
    if (launches) *launches = l.launches;
    return UALM_OK;
}

int tp_mark_begin(TpEngine *e, std::string *err)
{
    TCK(cudaSetDevice(e->device));
    TCK(cudaEventRecord(e->evA, e->stream));
    return UALM_OK;
}
int tp_mark_end(TpEngine *e, float *ms, std::string *err)
{
    TCK(cudaSetDevice(e->device));
    for (int g = 0; g < TP_NGROUPS; g++) { TCK(cudaEventRecord(e->gev[g], e->gstream[g])); TCK(cudaStreamWaitEvent(e->stream, e->gev[g], 0)); }
    TCK(cudaEventRecord(e->evB, e->stream));
    TCK(cudaEventSynchronize(e->evB));
    TCK(cudaEventElapsedTime(ms, e->evA, e->evB));
    return UALM_OK;
}

static int put(TpEngine *e, Buf<double> &dst, const double *src, size_t n, bool &has, std::string *err)
{
    has = src != nullptr;
    if (!src || n == 0) return UALM_OK;
    TCK(dst.ensure(n));
    TCK(cudaMemcpyAsync(dst.p, src, n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    return UALM_OK;
}

int tp_eval(TpEngine *e, int lane, const double *x, const double *lambda, const double *mu, const double *scale_cx, const double *scale_fx, double rho,
            double *f, double *grad, double *hx, double *gx, double *c_xy, double *c_yaw, std::string *err)
{
    TpLane &l = e->lanes[lane];
    if (!l.have_batch || !e->have_map) { if (err) *err = "upload and set_map first"; return UALM_ESTATE; }
    for (int b = 0; b < l.B; b++) if (l.N[b] > TP_NMAX || l.M[b] > TP_MMAX) { if (err) *err = "batch holds problems over the compiled limits"; return UALM_ELIMIT; }
    TCK(cudaSetDevice(e->device));
    if (x && l.tot_x > 0) TCK(cudaMemcpyAsync(l.d_x0.p, x, sizeof(double) * l.tot_x, cudaMemcpyHostToDevice, e->stream));
    int rc;
    if ((rc = put(e, l.d_lam, lambda, l.tot_s, l.has_lam, err))) return rc;
    if ((rc = put(e, l.d_mu, mu, 6 * l.tot_s, l.has_mu, err))) return rc;
    if ((rc = put(e, l.d_scin, scale_cx, 7 * l.tot_s, l.has_scx, err))) return rc;
    if ((rc = put(e, l.d_sfin, scale_fx, l.B, l.has_sfx, err))) return rc;
    l.rho_eval = rho;
    rc = admit_impl(e, lane, 1, false, err);
    l.has_lam = l.has_mu = l.has_scx = l.has_sfx = false;
    if (rc) return rc;
    if ((rc = tp_collect(e, lane, err))) return rc;
    if (l.B > 0) {
        if (f) TCK(cudaMemcpyAsync(f, l.d_f.p, sizeof(double) * l.B, cudaMemcpyDeviceToHost, e->stream));
        if (grad) TCK(cudaMemcpyAsync(grad, l.d_grad.p, sizeof(double) * l.tot_x, cudaMemcpyDeviceToHost, e->stream));
        if (hx) TCK(cudaMemcpyAsync(hx, l.d_hx.p, sizeof(double) * l.tot_s, cudaMemcpyDeviceToHost, e->stream));
        if (gx) TCK(cudaMemcpyAsync(gx, l.d_gx.p, sizeof(double) * 6 * l.tot_s, cudaMemcpyDeviceToHost, e->stream));
        if (c_xy) TCK(cudaMemcpyAsync(c_xy, l.d_cxy.p, sizeof(double) * l.tot_cxy, cudaMemcpyDeviceToHost, e->stream));
        if (c_yaw) TCK(cudaMemcpyAsync(c_yaw, l.d_cyaw.p, sizeof(double) * l.tot_cyaw, cudaMemcpyDeviceToHost, e->stream));
    }
    TCK(cudaStreamSynchronize(e->stream));
    l.collected = false;           // the lane holds evaluation outputs, not a solve
    return UALM_OK;
}

int tp_init_scaling(TpEngine *e, int lane, double *scale_fx, double *scale_cx, std::string *err)
{
    TpLane &l = e->lanes[lane];
    if (!l.have_batch || !e->have_map) { if (err) *err = "upload and set_map first"; return UALM_ESTATE; }
    for (int b = 0; b < l.B; b++) if (l.N[b] > TP_NMAX || l.M[b] > TP_MMAX) { if (err) *err = "batch holds problems over the compiled limits"; return UALM_ELIMIT; }
    TCK(cudaSetDevice(e->device));
    int rc = admit_impl(e, lane, 2, false, err);
    if (rc) return rc;
    if ((rc = tp_collect(e, lane, err))) return rc;
    if (l.B > 0) {
        if (scale_fx) TCK(cudaMemcpyAsync(scale_fx, l.d_sfx.p, sizeof(double) * l.B, cudaMemcpyDeviceToHost, e->stream));
        if (scale_cx) TCK(cudaMemcpyAsync(scale_cx, l.d_scx.p, sizeof(double) * 7 * l.tot_s, cudaMemcpyDeviceToHost, e->stream));
    }
    TCK(cudaStreamSynchronize(e->stream));
    l.collected = false;
    return UALM_OK;
}

int tp_time_penalty(TpEngine *e, int lane, int reps, int use_tma, float *ms_per_launch, double *algorithmic_bytes, std::string *err)
{
    TpLane &l = e->lanes[lane];
    if (!l.have_batch || !e->have_map) { if (err) *err = "upload and set_map first"; return UALM_ESTATE; }
    if (l.in_flight || e->live > 0) { if (err) *err = "ualm_time_penalty_kernel needs an idle pool"; return UALM_ESTATE; }
    if (ms_per_launch) *ms_per_launch = 0.f;
    if (algorithmic_bytes) *algorithmic_bytes = 0.0;
    if (l.B == 0) return UALM_OK;
    TCK(cudaSetDevice(e->device));
    // admit as single evaluations (initial guess, zero duals, unit scales), run the forward half of a round, then time kb alone
    l.has_lam = l.has_mu = l.has_scx = l.has_sfx = false; l.rho_eval = e->p.rho;
    int rc = admit_impl(e, lane, 1, true, err);
    if (rc) return rc;
    rc = e->f32() ? set_attrs<float>(e, err) : set_attrs<double>(e, err);
    if (rc) return rc;
    const int g = lane * TP_SUBGROUPS;
    cudaStream_t st = e->gstream[g];
    compact_kernel<<<1, 1024, 0, st>>>(e->E, e->capacity, g);
    e->gcompact[g] = false;
    const int upper = e->glive[g];
    const bool tma = e->have_tmap && use_tma;
    if (e->f32()) ka_kernel<float><<<(upper + e->ka_warps - 1) / e->ka_warps, 32 * e->ka_warps, ka_smem_bytes(e), st>>>(e->E, e->p, g);
    else ka_kernel<double><<<(upper + e->ka_warps - 1) / e->ka_warps, 32 * e->ka_warps, ka_smem_bytes(e), st>>>(e->E, e->p, g);
    auto kb = [&]() {
        if (e->f32()) {
            if (tma) kb_kernel<float, true><<<upper, TP_KB_THREADS, kb_smem_bytes(e, true), st>>>(e->E, e->p, e->map, e->tmap, g);
            else kb_kernel<float, false><<<upper, TP_KB_THREADS, kb_smem_bytes(e, false), st>>>(e->E, e->p, e->map, e->tmap, g);
        } else {
            if (tma) kb_kernel<double, true><<<upper, TP_KB_THREADS, kb_smem_bytes(e, true), st>>>(e->E, e->p, e->map, e->tmap, g);
            else kb_kernel<double, false><<<upper, TP_KB_THREADS, kb_smem_bytes(e, false), st>>>(e->E, e->p, e->map, e->tmap, g);
        }
    };
    kb();    // warm-up
    TCK(cudaEventRecord(e->evA, st));
    for (int r = 0; r < reps; r++) kb();
    TCK(cudaEventRecord(e->evB, st));
    TCK(cudaGetLastError());
    TCK(cudaEventSynchronize(e->evB));
    float ms = 0.f;
    TCK(cudaEventElapsedTime(&ms, e->evA, e->evB));
    if (ms_per_launch) *ms_per_launch = ms / reps;
    if (algorithmic_bytes) {
        // per trajectory per evaluation S * 45 e + (25 N + 13 M) e, e = element size of the sample data
        double bytes = 0;
        const double es = (double)e->esz();
        for (int b = 0; b < l.B; b++) bytes += (double)l.N[b] * (e->p.int_K + 1) * 45 * es + (25.0 * l.N[b] + 13.0 * l.M[b]) * es;
        *algorithmic_bytes = bytes;
    }
    // finish the evaluations and release the slots
    rc = tp_collect(e, lane, err);
    l.collected = false;
    return rc;
}

int tp_lane_outputs(TpEngine *e, int lane, const ualm_result_t **d_res, const double **d_cxy, const double **d_cyaw, int *B, const int32_t **N,
                    const int32_t **M)
{
    TpLane &l = e->lanes[lane];
    if (!l.collected) return UALM_ESTATE;
    *d_res = l.d_res.p; *d_cxy = l.d_cxy.p; *d_cyaw = l.d_cyaw.p; *B = l.B; *N = l.N.data(); *M = l.M.data();
    return UALM_OK;
}

} // namespace ualm_tp
