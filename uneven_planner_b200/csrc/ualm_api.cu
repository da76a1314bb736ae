// ualm_api.cu -- the extern "C" layer declared in include/ualm.h: context, uploads, kernel launches.
// Host code is thin: it packs descriptors, sizes shared memory from the batch maxima and launches the
// per-trajectory kernels of ualm_kernels.cuh.  There is NO CPU fallback: every compute entry point needs a CUDA
// device and returns UALM_ENOCUDA otherwise.
#include "ualm_kernels.cuh"
#include "map_prep.h"

#include <algorithm>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

using namespace ualm;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return fail(UALM_ENOCUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

template <class T>
struct DevBuf {
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

struct ualm_ctx {
    int device = 0, precision = 64;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool have_params = false, have_map = false, have_batch = false, solved = false;
    ualm_params_t hp;
    DevParams dp;
    DevMap dm;
    DevBuf<float4> cells;
    // batch
    int B = 0, Nmax = 0, Mmax = 0, nmax = 0, Smax = 0;
    std::vector<ProbDesc> desc;
    std::vector<int> order;
    long long tot_x = 0, tot_s = 0, tot_cxy = 0, tot_cyaw = 0, tot_hist = 0, tot_scr = 0, tot_fac = 0, tot_ws = 0;
    DevBuf<ProbDesc> d_desc;
    DevBuf<int> d_order;
    DevBuf<int4> d_wdesc;
    std::vector<int4> wdesc;
    std::vector<int> group;      // warps per problem (1, 2 or 4)
    int group_mode = 1;
    struct GClass { int G = 1, n_ctas = 0, r0 = 0, r1 = 0, occ = 1; size_t wd_off = 0, smem = 0; SmemLayout L; };
    std::vector<GClass> cls;
    enum { NAUX = 7 };
    cudaStream_t aux[NAUX] = {};
    cudaEvent_t evs[NAUX + 1] = {};
    DevBuf<double> d_x0, d_x, d_lambda, d_mu, d_scale_cx, d_hx, d_gx, d_lm_s, d_lm_y, d_lm_aux, d_fac, d_scr, d_ws, d_cxy, d_cyaw, d_f, d_grad, d_sfx;
    DevBuf<ualm_result_t> d_res;
    DevBuf<long long> d_prof;
    DevBuf<double> d_pieceT, d_feas;
    bool profile = false;
    float last_ms = 0.f;
    int last_launches = 0;
    SmemLayout L;
    size_t smem_bytes = 0;
};

extern "C" const char *ualm_last_error(void) { return g_err.c_str(); }

extern "C" int ualm_create(ualm_ctx_t **out, int device, int precision)
{
    if (!out) return fail(UALM_EINVAL, "ctx out pointer is NULL");
    if (precision != 64) return fail(UALM_EINVAL, "only precision=64 (bit-reproducible double path) is built in this round");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) return fail(UALM_ENOCUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(UALM_EINVAL, "device ordinal out of range");
    CK(cudaSetDevice(device));
    ualm_ctx *c = new ualm_ctx();
    c->device = device; c->precision = precision;
    if (const char *e = getenv("UALM_GROUPS")) c->group_mode = atoi(e);   // 0 = one warp per trajectory everywhere (developer switch)
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (int q = 0; q < ualm_ctx::NAUX; q++) CK(cudaStreamCreateWithFlags(&c->aux[q], cudaStreamNonBlocking));
    for (int q = 0; q < ualm_ctx::NAUX + 1; q++) CK(cudaEventCreateWithFlags(&c->evs[q], cudaEventDisableTiming));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    *out = c;
    return UALM_OK;
}

extern "C" int ualm_destroy(ualm_ctx_t *c)
{
    if (!c) return UALM_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->cells.release(); c->d_desc.release(); c->d_order.release(); c->d_wdesc.release();
    DevBuf<double> *bufs[] = {&c->d_x0, &c->d_x, &c->d_lambda, &c->d_mu, &c->d_scale_cx, &c->d_hx, &c->d_gx, &c->d_lm_s, &c->d_lm_y, &c->d_lm_aux, &c->d_fac,
                              &c->d_scr, &c->d_ws, &c->d_cxy, &c->d_cyaw, &c->d_f, &c->d_grad, &c->d_sfx};
    for (auto *b : bufs) b->release();
    c->d_res.release(); c->d_prof.release(); c->d_pieceT.release(); c->d_feas.release();
    for (int q = 0; q < ualm_ctx::NAUX; q++) cudaStreamDestroy(c->aux[q]);
    for (int q = 0; q < ualm_ctx::NAUX + 1; q++) cudaEventDestroy(c->evs[q]);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->own_stream);
    delete c;
    return UALM_OK;
}

extern "C" int ualm_set_stream(ualm_ctx_t *c, void *s)
{
    if (!c) return fail(UALM_EINVAL, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return UALM_OK;
}

extern "C" int ualm_set_params(ualm_ctx_t *c, const ualm_params_t *p)
{
    if (!c || !p) return fail(UALM_EINVAL, "null argument");
    if (p->int_K < 1 || p->int_K > 128) return fail(UALM_ELIMIT, "int_K out of range [1,128]");
    if (p->mem_size < 1 || p->mem_size > 1024) return fail(UALM_ELIMIT, "mem_size out of range [1,1024]");
    if (p->past < 0 || p->past > 16) return fail(UALM_ELIMIT, "past out of range [0,16]");
    c->hp = *p;
    DevParams &d = c->dp;
    d.rho_T = p->rho_T; d.rho_ter = p->rho_ter; d.max_vel = p->max_vel; d.max_acc_lon = p->max_acc_lon; d.max_acc_lat = p->max_acc_lat;
    d.max_kap = p->max_kap; d.min_cxi = p->min_cxi; d.max_sig = p->max_sig; d.use_scaling = p->use_scaling; d.rho = p->rho;
    d.beta = p->beta; d.gamma = p->gamma; d.epsilon_con = p->epsilon_con; d.max_iter = p->max_iter; d.g_epsilon = p->g_epsilon;
    d.min_step = p->min_step; d.delta = p->delta; d.inner_max_iter = (int)p->inner_max_iter; d.mem_size = p->mem_size; d.past = p->past;
    d.int_K = p->int_K; d.gravity = p->gravity;
    c->have_params = true;
    return UALM_OK;
}

extern "C" int ualm_set_map(ualm_ctx_t *c, const ualm_map_geom_t *g, const float *cells)
{
    if (!c || !g || !cells) return fail(UALM_EINVAL, "null argument");
    CK(cudaSetDevice(c->device));
    const size_t ncell = (size_t)g->voxel_num[0] * g->voxel_num[1] * g->voxel_num[2];
    CK(c->cells.ensure(ncell));
    CK(cudaMemcpyAsync(c->cells.p, cells, ncell * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    DevMap &m = c->dm;
    m.cells = c->cells.p;
    for (int k = 0; k < 3; k++) { m.vn[k] = g->voxel_num[k]; m.origin[k] = g->origin[k]; m.maxb[k] = g->max_boundary[k]; }
    m.xy_res = g->xy_resolution; m.yaw_res = g->yaw_resolution;
    m.xy_inv = 1.0 / g->xy_resolution; m.yaw_inv = 1.0 / g->yaw_resolution; // uneven_map.cpp:104-105
    c->have_map = true;
    return UALM_OK;
}

static int prepare_launch(ualm_ctx *c)
{
    c->L = make_layout(c->Nmax, c->Mmax, c->nmax, c->dp.mem_size, c->dp.past, c->dp.int_K, c->Smax);
    c->smem_bytes = (size_t)c->L.total_doubles * sizeof(double) * UALM_WPB;
    if (c->smem_bytes > 227 * 1024) return fail(UALM_ELIMIT, "problem too large for shared memory (N/M/int_K too big)");
    CK(cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
    CK(cudaFuncSetAttribute(eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
    CK(cudaFuncSetAttribute(scaling_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
    CK(cudaFuncSetAttribute(penalty_only_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
    return UALM_OK;
}

static BatchPtrs batch_ptrs(ualm_ctx *c)
{
    BatchPtrs b;
    b.B = c->B;
    b.wdesc = c->d_wdesc.p;
    b.n_leader_slots = 4;
    b.adopt = getenv("UALM_NOADOPT") ? 0 : 1;
    b.desc = c->d_desc.p; b.order = c->d_order.p; b.x0 = c->d_x0.p; b.x = c->d_x.p;
    b.lambda = c->d_lambda.p; b.mu = c->d_mu.p; b.scale_cx = c->d_scale_cx.p; b.hx = c->d_hx.p; b.gx = c->d_gx.p;
    b.lm_s = c->d_lm_s.p; b.lm_y = c->d_lm_y.p; b.lm_aux = c->d_lm_aux.p; b.fac = c->d_fac.p; b.scratch = c->d_scr.p; b.ws_scaling = c->d_ws.p;
    b.c_xy = c->d_cxy.p; b.c_yaw = c->d_cyaw.p; b.results = c->d_res.p; b.piece_T = c->d_pieceT.p; b.f_out = c->d_f.p; b.grad_out = c->d_grad.p;
    b.scale_fx_io = c->d_sfx.p;
    b.prof = c->profile ? c->d_prof.p : nullptr;
    return b;
}

extern "C" int ualm_upload(ualm_ctx_t *c, int B, const int32_t *N, const int32_t *M, const double *bnd, const double *total_time,
                           const double *inner_xy, const double *inner_yaw)
{
    if (!c || B < 0 || (B > 0 && (!N || !M || !bnd || !total_time))) return fail(UALM_EINVAL, "bad argument");
    if (!c->have_params) return fail(UALM_ESTATE, "ualm_set_params must be called before ualm_upload");
    CK(cudaSetDevice(c->device));
    const int K = c->dp.int_K, m = c->dp.mem_size;
    c->B = B; c->desc.resize(B); c->order.resize(B); c->group.assign(B, 1);
    c->Nmax = c->Mmax = c->nmax = c->Smax = 1;
    for (int b = 0; b < B; b++) {
        ProbDesc &d = c->desc[b];
        if (N[b] < 1 || M[b] < 1 || N[b] > 64 || M[b] > 128) return fail(UALM_ELIMIT, "piece count outside [1,64] x [1,128]");
        d.N = N[b]; d.M = M[b]; d.n = 1 + 2 * (N[b] - 1) + (M[b] - 1); d.S = N[b] * (K + 1);
        c->Nmax = std::max(c->Nmax, d.N); c->Mmax = std::max(c->Mmax, d.M); c->nmax = std::max(c->nmax, d.n); c->Smax = std::max(c->Smax, d.S);
    }
    // launch order: most samples first (longest-processing-time-first keeps the tail short)
    std::iota(c->order.begin(), c->order.end(), 0);
    std::stable_sort(c->order.begin(), c->order.end(), [&](int a, int b2) { return c->desc[a].S > c->desc[b2].S; });
    int rc0 = prepare_launch(c);
    if (rc0) return rc0;
    // Warp groups and size classes.  Problems are sorted by size; the largest get 4 or 2 warps (helpers for the parallel
    // phases) as far as the whole batch still fits on the device at once.  Every (group size, size bucket) class is its own
    // launch with its own shared-memory layout sized by the class maxima: a G=4 CTA holds one trajectory slot + 3 helper
    // rings, a G=2 CTA two slots + 2 rings, a G=1 CTA four slots.  The classes run concurrently on separate streams.
    {
        int dev_sms = 0;
        CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, c->device));
        const int RINGD = 2 * UALM_RINGB * 6 * UALM_FW;
        int occ0 = 1;   // CTAs per SM the register file allows (shared memory is checked per class below)
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, solve_kernel, UALM_THREADS * UALM_WPB, 16 * 1024));
        if (getenv("UALM_DEBUG")) {
            cudaFuncAttributes fa;
            CK(cudaFuncGetAttributes(&fa, solve_kernel));
            fprintf(stderr, "[ualm] solve_kernel: %d regs/thread, %zu B local, %zu B static smem, %d CTAs/SM by registers\n", fa.numRegs, fa.localSizeBytes,
                    fa.sharedSizeBytes, occ0);
        }
        const long long W = 4LL * occ0 * dev_sms;   // warp slots on the device
        std::vector<ualm_ctx::GClass> best;
        for (int attempt = 0; attempt < 32; attempt++) {
            // Policy (measured on B200, tools/gpu_policy_dev.py): helpers shorten the latency of a trajectory without adding to the
            // L2-resident working set, while more resident trajectories than ~600 thrash the 126 MB L2.  So: everything that fits
            // gets 4 warps; beyond that 2 warps per trajectory with the largest tenth at 4, run in waves.
            long long K4 = 0, K2 = 0;
            if (c->group_mode) {
                if (4LL * B <= W) K4 = B;
                else if (2LL * B <= W) { K4 = std::min<long long>(B, (W - 2LL * B) / 2); K2 = (B - K4) & ~1LL; }
                else { K4 = B / 10; K2 = (B - K4) & ~1LL; }
                if (attempt > 0) { K4 = (long long)(K4 * std::pow(0.8, attempt)); K2 = std::min<long long>(B - K4, K2) & ~1LL; }
            }
            const bool forced = getenv("UALM_F4") || getenv("UALM_F2");   // developer override: fractions of the batch
            if (forced) {
                K4 = (long long)(B * (getenv("UALM_F4") ? atof(getenv("UALM_F4")) : 0.0));
                K2 = std::min<long long>(B - K4, (long long)(B * (getenv("UALM_F2") ? atof(getenv("UALM_F2")) : 0.0))) & ~1LL;
            }
            // rank ranges per group size, each split into size buckets (the order is by descending size)
            std::vector<ualm_ctx::GClass> cls;
            const long long lim[4] = {0, K4, K4 + K2, B};
            for (int gi = 0; gi < 3; gi++) {
                const int G = gi == 0 ? 4 : gi == 1 ? 2 : 1;
                const long long a = lim[gi], b2 = lim[gi + 1];
                if (b2 <= a) continue;
                const int nsplit = (b2 - a) >= 96 ? (gi == 2 ? 3 : 2) : 1;
                for (int sp = 0; sp < nsplit; sp++) {
                    ualm_ctx::GClass cl;
                    cl.G = G;
                    long long s0 = a + (b2 - a) * sp / nsplit, s1 = a + (b2 - a) * (sp + 1) / nsplit;
                    const int per = 4 / G;                         // problems per CTA
                    if (sp > 0) s0 = a + ((s0 - a + per - 1) / per) * per;
                    if (sp + 1 < nsplit) s1 = a + ((s1 - a + per - 1) / per) * per;
                    if (s1 <= s0) continue;
                    cl.r0 = (int)s0; cl.r1 = (int)s1;
                    int Nm = 1, Mm = 1, nm = 1, Sm = 1;
                    for (long long r = s0; r < s1; r++) {
                        const ProbDesc &d = c->desc[c->order[r]];
                        Nm = std::max(Nm, d.N); Mm = std::max(Mm, d.M); nm = std::max(nm, d.n); Sm = std::max(Sm, d.S);
                    }
                    cl.L = make_layout(Nm, Mm, nm, c->dp.mem_size, c->dp.past, c->dp.int_K, Sm);
                    cl.smem = ((size_t)per * cl.L.total_doubles + (size_t)(4 - per) * RINGD) * sizeof(double);
                    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cl.occ, solve_kernel, UALM_THREADS * UALM_WPB, cl.smem));
                    cl.n_ctas = (int)((s1 - s0 + per - 1) / per);
                    cls.push_back(cl);
                }
            }
            double sm_need = 0.0;
            for (auto &cl : cls) sm_need += (double)cl.n_ctas / std::max(cl.occ, 1);
            best = cls;
            // multi-wave batches need not fit at once; batches meant to be fully resident must
            if (forced || 2LL * B > W || sm_need <= 0.98 * dev_sms || attempt >= 12) break;
        }
        c->cls = best;
        c->wdesc.clear();
        for (auto &cl : c->cls) {
            cl.wd_off = c->wdesc.size();
            for (int r = cl.r0; r < cl.r1; r++) c->group[c->order[r]] = cl.G;
            int q = cl.r0;
            while (q < cl.r1) {
                if (cl.G == 4) {
                    for (int w = 0; w < 4; w++) c->wdesc.push_back(make_int4(c->order[q], 0, 4, w | ((w ? w - 1 : 0) << 8)));
                    q += 1;
                } else if (cl.G == 2) {
                    // CTA j of the class pairs its j-th largest with its j-th smallest problem: the small one ends early and its
                    // two warps then help the large one (adoption, ualm_kernels.cuh)
                    const int j = (q - cl.r0) / 2, cnt = cl.r1 - cl.r0;
                    const int pa = cl.r0 + j, pb = cl.r1 - 1 - j;
                    const int pr[2] = {c->order[pa], pb > pa ? c->order[pb] : -1};
                    for (int h = 0; h < 2; h++) {
                        if (pr[h] >= 0) {
                            c->wdesc.push_back(make_int4(pr[h], h, 2, 0));
                            c->wdesc.push_back(make_int4(pr[h], h, 2, 1 | (h << 8)));
                        } else { c->wdesc.push_back(make_int4(-1, 0, 1, 0)); c->wdesc.push_back(make_int4(-1, 0, 1, 0)); }
                    }
                    (void)cnt;
                    q += 2;
                } else {
                    for (int w = 0; w < 4; w++) {
                        if (q < cl.r1) { c->wdesc.push_back(make_int4(c->order[q], w, 1, 0)); q += 1; }
                        else c->wdesc.push_back(make_int4(-1, 0, 1, 0));
                    }
                }
            }
            if ((int)((c->wdesc.size() - cl.wd_off) / 4) != cl.n_ctas) return fail(UALM_EINVAL, "internal: class CTA count mismatch");
            if (getenv("UALM_DEBUG"))
                fprintf(stderr, "[ualm] class G=%d: problems [%d,%d) %d CTAs, smem %zu B, %d CTAs/SM\n", cl.G, cl.r0, cl.r1, cl.n_ctas, cl.smem, cl.occ);
        }
    }
    long long ox = 0, os = 0, ocx = 0, ocy = 0, oh = 0, oscr = 0, oixy = 0, oiyaw = 0, ofac = 0, ows = 0;
    std::vector<double> x0;
    for (int b = 0; b < B; b++) {
        ProbDesc &d = c->desc[b];
        d.off_x = ox; d.off_s = os; d.off_cxy = ocx; d.off_cyaw = ocy; d.off_hist = oh; d.off_scr = oscr; d.off_fac = ofac; d.off_ws = ows;
        for (int k = 0; k < 18; k++) d.bnd[k] = bnd[(size_t)b * 18 + k];
        d.total_time = total_time[b];
        // x = [tau | Pxy | Pyaw]  (alm_traj_opt.cpp:205-216); logC2 (alm_traj_opt.h:238-241) is two IEEE ops + sqrt
        const double T = total_time[b];
        x0.push_back(T > 1.0 ? (std::sqrt(2.0 * T - 1.0) - 1.0) : (1.0 - std::sqrt(2.0 / T - 1.0)));
        for (int q = 0; q < 2 * (d.N - 1); q++) x0.push_back(inner_xy[oixy + q]);
        for (int q = 0; q < d.M - 1; q++) x0.push_back(inner_yaw[oiyaw + q]);
        oixy += 2 * (d.N - 1); oiyaw += d.M - 1;
        ox += d.n; os += d.S; ocx += 12 * d.N; ocy += 6 * d.M; oh += (long long)m * d.n; oscr += (long long)UALM_NFIELD * d.S;
        ofac += 1LL * UALM_FW * ((6 * d.N + 2 * UALM_FPAD) + (6 * d.M + 2 * UALM_FPAD));
        ows += (long long)(12 * d.N + 6 * d.M) * 32 * c->group[b];
    }
    c->tot_x = ox; c->tot_s = os; c->tot_cxy = ocx; c->tot_cyaw = ocy; c->tot_hist = oh; c->tot_scr = oscr; c->tot_fac = ofac; c->tot_ws = ows;
    CK(c->d_desc.ensure(B)); CK(c->d_order.ensure(B)); CK(c->d_wdesc.ensure(c->wdesc.size())); CK(c->d_x0.ensure(ox)); CK(c->d_x.ensure(ox)); CK(c->d_grad.ensure(ox));
    CK(c->d_lambda.ensure(os)); CK(c->d_hx.ensure(os)); CK(c->d_mu.ensure(6 * os)); CK(c->d_gx.ensure(6 * os)); CK(c->d_scale_cx.ensure(7 * os));
    CK(c->d_lm_s.ensure(oh)); CK(c->d_lm_y.ensure(oh)); CK(c->d_scr.ensure(oscr));
    CK(c->d_ws.ensure(ows)); CK(c->d_fac.ensure(ofac)); CK(c->d_lm_aux.ensure((size_t)std::max(B, 1) * 3 * m));
    // factor arrays: entries outside the band-in-matrix positions (and the pad rows) are never written and must read 0
    if (ofac > 0) CK(cudaMemsetAsync(c->d_fac.p, 0, sizeof(double) * ofac, c->stream));
    CK(c->d_prof.ensure((size_t)std::max(B, 1) * UALM_NPROF));
    CK(c->d_pieceT.ensure((size_t)std::max(B, 1) * 2)); CK(c->d_feas.ensure((size_t)std::max(B, 1) * 10));
    CK(c->d_cxy.ensure(ocx)); CK(c->d_cyaw.ensure(ocy)); CK(c->d_res.ensure(B)); CK(c->d_f.ensure(B)); CK(c->d_sfx.ensure(B));
    if (B > 0) {
        CK(cudaMemcpyAsync(c->d_desc.p, c->desc.data(), sizeof(ProbDesc) * B, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_order.p, c->order.data(), sizeof(int) * B, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_wdesc.p, c->wdesc.data(), sizeof(int4) * c->wdesc.size(), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d_x0.p, x0.data(), sizeof(double) * ox, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream)); // x0 is a stack-local staging vector
    }
    c->have_batch = true; c->solved = false;
    return UALM_OK;
}

extern "C" int ualm_solve_resident(ualm_ctx_t *c)
{
    if (!c) return fail(UALM_EINVAL, "null ctx");
    if (!c->have_map || !c->have_params || !c->have_batch) return fail(UALM_ESTATE, "set_params, set_map and upload must precede solve");
    CK(cudaSetDevice(c->device));
    c->last_launches = 0;
    CK(cudaEventRecord(c->ev0, c->stream));
    if (c->B > 0) {
        // the classes run concurrently: class 0 on the context stream, the others on auxiliary streams
        const int last = ualm_ctx::NAUX;
        CK(cudaEventRecord(c->evs[last], c->stream));
        for (size_t ci = 0; ci < c->cls.size(); ci++) {
            auto &cl = c->cls[ci];
            if (cl.n_ctas == 0) continue;
            const int ai = (int)((ci - 1) % ualm_ctx::NAUX);
            cudaStream_t st = ci == 0 ? c->stream : c->aux[ai];
            if (ci > 0) CK(cudaStreamWaitEvent(st, c->evs[last], 0));
            BatchPtrs bp = batch_ptrs(c);
            bp.wdesc = c->d_wdesc.p + cl.wd_off;
            bp.n_leader_slots = 4 / cl.G;
            solve_kernel<<<cl.n_ctas, UALM_THREADS * UALM_WPB, cl.smem, st>>>(bp, c->dp, c->dm, cl.L);
            CK(cudaGetLastError());
            c->last_launches++;
            if (ci > 0) { CK(cudaEventRecord(c->evs[ai], st)); CK(cudaStreamWaitEvent(c->stream, c->evs[ai], 0)); }
        }
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    c->solved = true;
    return UALM_OK;
}

extern "C" int ualm_sync(ualm_ctx_t *c)
{
    if (!c) return fail(UALM_EINVAL, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return UALM_OK;
}

extern "C" int ualm_last_solve_ms(ualm_ctx_t *c, float *ms, int *launches)
{
    if (!c || !c->solved) return fail(UALM_ESTATE, "no solve to time");
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    if (ms) *ms = c->last_ms;
    if (launches) *launches = c->last_launches;
    return UALM_OK;
}

extern "C" int ualm_download(ualm_ctx_t *c, ualm_result_t *results, double *c_xy, double *c_yaw)
{
    if (!c || !c->solved) return fail(UALM_ESTATE, "nothing solved");
    CK(cudaSetDevice(c->device));
    if (c->B > 0) {
        if (results) CK(cudaMemcpyAsync(results, c->d_res.p, sizeof(ualm_result_t) * c->B, cudaMemcpyDeviceToHost, c->stream));
        if (c_xy) CK(cudaMemcpyAsync(c_xy, c->d_cxy.p, sizeof(double) * c->tot_cxy, cudaMemcpyDeviceToHost, c->stream));
        if (c_yaw) CK(cudaMemcpyAsync(c_yaw, c->d_cyaw.p, sizeof(double) * c->tot_cyaw, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return UALM_OK;
}

extern "C" int ualm_solve_batch(ualm_ctx_t *c, int B, const int32_t *N, const int32_t *M, const double *bnd, const double *total_time,
                                const double *inner_xy, const double *inner_yaw, ualm_result_t *results, double *c_xy, double *c_yaw)
{
    int rc = ualm_upload(c, B, N, M, bnd, total_time, inner_xy, inner_yaw);
    if (rc) return rc;
    rc = ualm_solve_resident(c);
    if (rc) return rc;
    return ualm_download(c, results, c_xy, c_yaw);
}

extern "C" int ualm_pack_records_device(ualm_ctx_t *c, double *d_records, int stride)
{
    if (!c || !c->solved || !d_records) return fail(UALM_ESTATE, "nothing solved / null buffer");
    if (stride < 12 + 12 * c->Nmax + 6 * c->Mmax) return fail(UALM_EINVAL, "record stride too small");
    CK(cudaSetDevice(c->device));
    if (c->B > 0) {
        pack_records_kernel<<<c->B, 128, 0, c->stream>>>(batch_ptrs(c), c->B, d_records, stride);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(c->stream));
    return UALM_OK;
}

// upload an optional host array (or a constant fill) into a device buffer
static int put(ualm_ctx *c, double *dst, const double *src, size_t n, double fill)
{
    if (n == 0) return UALM_OK;
    if (src) { CK(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream)); }
    else {
        std::vector<double> tmp(n, fill);
        CK(cudaMemcpyAsync(dst, tmp.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return UALM_OK;
}

extern "C" int ualm_eval_batch(ualm_ctx_t *c, const double *x, const double *lambda, const double *mu, const double *scale_cx,
                               const double *scale_fx, double rho, double *f, double *grad, double *hx, double *gx, double *c_xy,
                               double *c_yaw)
{
    if (!c || !c->have_batch || !c->have_map) return fail(UALM_ESTATE, "upload and set_map first");
    CK(cudaSetDevice(c->device));
    int rc;
    if (x) { if ((rc = put(c, c->d_x0.p, x, c->tot_x, 0.0))) return rc; }
    if ((rc = put(c, c->d_lambda.p, lambda, c->tot_s, 0.0))) return rc;
    if ((rc = put(c, c->d_mu.p, mu, 6 * c->tot_s, 0.0))) return rc;
    if ((rc = put(c, c->d_scale_cx.p, scale_cx, 7 * c->tot_s, 1.0))) return rc;
    if ((rc = put(c, c->d_sfx.p, scale_fx, c->B, 1.0))) return rc;
    if (c->B > 0) {
        eval_kernel<<<(c->B + UALM_WPB - 1) / UALM_WPB, UALM_THREADS * UALM_WPB, c->smem_bytes, c->stream>>>(batch_ptrs(c), c->dp, c->dm, c->L, rho);
        CK(cudaGetLastError());
        if (f) CK(cudaMemcpyAsync(f, c->d_f.p, sizeof(double) * c->B, cudaMemcpyDeviceToHost, c->stream));
        if (grad) CK(cudaMemcpyAsync(grad, c->d_grad.p, sizeof(double) * c->tot_x, cudaMemcpyDeviceToHost, c->stream));
        if (hx) CK(cudaMemcpyAsync(hx, c->d_hx.p, sizeof(double) * c->tot_s, cudaMemcpyDeviceToHost, c->stream));
        if (gx) CK(cudaMemcpyAsync(gx, c->d_gx.p, sizeof(double) * 6 * c->tot_s, cudaMemcpyDeviceToHost, c->stream));
        if (c_xy) CK(cudaMemcpyAsync(c_xy, c->d_cxy.p, sizeof(double) * c->tot_cxy, cudaMemcpyDeviceToHost, c->stream));
        if (c_yaw) CK(cudaMemcpyAsync(c_yaw, c->d_cyaw.p, sizeof(double) * c->tot_cyaw, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return UALM_OK;
}

extern "C" int ualm_init_scaling_batch(ualm_ctx_t *c, double *scale_fx, double *scale_cx)
{
    if (!c || !c->have_batch || !c->have_map) return fail(UALM_ESTATE, "upload and set_map first");
    CK(cudaSetDevice(c->device));
    if (c->B > 0) {
        scaling_kernel<<<(c->B + UALM_WPB - 1) / UALM_WPB, UALM_THREADS * UALM_WPB, c->smem_bytes, c->stream>>>(batch_ptrs(c), c->dp, c->dm, c->L);
        CK(cudaGetLastError());
        if (scale_fx) CK(cudaMemcpyAsync(scale_fx, c->d_sfx.p, sizeof(double) * c->B, cudaMemcpyDeviceToHost, c->stream));
        if (scale_cx) CK(cudaMemcpyAsync(scale_cx, c->d_scale_cx.p, sizeof(double) * 7 * c->tot_s, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return UALM_OK;
}

extern "C" int ualm_time_penalty_kernel(ualm_ctx_t *c, int reps, float *ms_per_launch, double *algorithmic_bytes)
{
    if (!c || !c->have_batch || !c->have_map) return fail(UALM_ESTATE, "upload and set_map first");
    if (reps < 1) return fail(UALM_EINVAL, "reps < 1");
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = put(c, c->d_lambda.p, nullptr, c->tot_s, 0.0))) return rc;
    if ((rc = put(c, c->d_mu.p, nullptr, 6 * c->tot_s, 0.0))) return rc;
    if ((rc = put(c, c->d_scale_cx.p, nullptr, 7 * c->tot_s, 1.0))) return rc;
    penalty_only_kernel<<<(c->B + UALM_WPB - 1) / UALM_WPB, UALM_THREADS * UALM_WPB, c->smem_bytes, c->stream>>>(batch_ptrs(c), c->dp, c->dm, c->L, 1); // warm-up
    CK(cudaEventRecord(c->ev0, c->stream));
    penalty_only_kernel<<<(c->B + UALM_WPB - 1) / UALM_WPB, UALM_THREADS * UALM_WPB, c->smem_bytes, c->stream>>>(batch_ptrs(c), c->dp, c->dm, c->L, reps);
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaGetLastError());
    CK(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (ms_per_launch) *ms_per_launch = ms / reps;
    if (algorithmic_bytes) {
        // per trajectory per evaluation S*45e + (25N + 13M)e
        double bytes = 0;
        for (auto &d : c->desc) bytes += (double)d.S * 45 * 8 + (25.0 * d.N + 13.0 * d.M) * 8;
        *algorithmic_bytes = bytes;
    }
    return UALM_OK;
}

// developer aid: per-phase SM cycle counters of the last solve (thread 0 of every CTA), summed over the batch.
// enable != 0 turns collection on for subsequent solves; out16 (may be NULL) receives the sums of the last solve.
extern "C" int ualm_profile(ualm_ctx_t *c, int enable, long long *out16)
{
    if (!c) return fail(UALM_EINVAL, "null ctx");
    CK(cudaSetDevice(c->device));
    if (out16 && c->profile && c->solved && c->B > 0) {
        std::vector<long long> h((size_t)c->B * UALM_NPROF);
        CK(cudaMemcpyAsync(h.data(), c->d_prof.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (enable == 2) { // raw: out16 has room for B x 16 values, rows in launch order
            for (size_t q = 0; q < h.size(); q++) out16[q] = h[q];
        } else {
            for (int q = 0; q < UALM_NPROF; q++) out16[q] = 0;
            for (int b = 0; b < c->B; b++) for (int q = 0; q < UALM_NPROF; q++) out16[q] += h[(size_t)b * UALM_NPROF + q];
        }
    }
    c->profile = enable != 0;
    return UALM_OK;
}

extern "C" int ualm_feasibility_batch(ualm_ctx_t *c, double dt, double *out10)
{
    if (!c || !out10 || !(dt > 0.0)) return fail(UALM_EINVAL, "null argument or dt <= 0");
    if (!c->have_map || !c->have_params || !c->have_batch || !c->solved) return fail(UALM_ESTATE, "ualm_feasibility_batch needs a solved resident batch");
    CK(cudaSetDevice(c->device));
    if (c->B == 0) return UALM_OK;
    const BatchPtrs bp = batch_ptrs(c);
    feasibility_kernel<<<(c->B + UALM_WPB - 1) / UALM_WPB, UALM_THREADS * UALM_WPB, 0, c->stream>>>(bp, c->dp, c->dm, dt, c->d_feas.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out10, c->d_feas.p, sizeof(double) * 10 * c->B, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return UALM_OK;
}

// ---------------------------------------------------------------------------------------------
// UnevenMap construction on the device.  HBM-bound gather: one thread per cell, yaw fastest so that the 64
// threads of a CTA share (x, y) and read the same few cloud bins through L1.
// ---------------------------------------------------------------------------------------------
namespace ualm {
__global__ void __launch_bounds__(64) map_build_kernel(UalmMapPrep prep, ualm_map_geom_t geom, float4 *cells)
{
    const int W = geom.voxel_num[2], Y = geom.voxel_num[1];
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)geom.voxel_num[0] * Y * W;
    if (cell >= total) return;
    const int w = (int)(cell % W), y = (int)((cell / W) % Y), x = (int)(cell / ((long long)W * Y));
    float o[4];
    ualm_map_cell(prep, geom, x, y, w, o);
    cells[cell] = make_float4(o[0], o[1], o[2], o[3]);
}
} // namespace ualm

extern "C" int ualm_map_build_device(ualm_ctx_t *c, const float *pin, int64_t npts, const ualm_map_geom_t *g, double ex, double ey, double ez,
                                     int iter_num, float *cells, float *kernel_ms)
{
    if (!c || !pin || !g || !cells || npts < 0) return fail(UALM_EINVAL, "null argument");
    CK(cudaSetDevice(c->device));
    UalmMapHostPrep prep;
    ualm_map_preprocess(pin, npts, ex, ey, ez, prep);
    UalmMapPrep view = prep.view(ex, ey, ez, iter_num);
    const long long total = (long long)g->voxel_num[0] * g->voxel_num[1] * g->voxel_num[2];
    // scratch buffers and events are released on every exit path (CK returns on the first CUDA error)
    struct Scratch {
        DevBuf<float> pts; DevBuf<int> start; DevBuf<float4> cells;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ~Scratch() { pts.release(); start.release(); cells.release(); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
    } sc;
    CK(sc.pts.ensure(std::max<size_t>(prep.pts.size(), 3)));
    CK(sc.start.ensure(prep.start.size()));
    CK(sc.cells.ensure((size_t)total));
    CK(cudaEventCreate(&sc.e0)); CK(cudaEventCreate(&sc.e1));
    if (!prep.pts.empty()) CK(cudaMemcpyAsync(sc.pts.p, prep.pts.data(), sizeof(float) * prep.pts.size(), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(sc.start.p, prep.start.data(), sizeof(int) * prep.start.size(), cudaMemcpyHostToDevice, c->stream));
    view.pts = sc.pts.p; view.start = sc.start.p;
    CK(cudaEventRecord(sc.e0, c->stream));
    map_build_kernel<<<(unsigned)((total + 63) / 64), 64, 0, c->stream>>>(view, *g, sc.cells.p);
    CK(cudaGetLastError());
    CK(cudaEventRecord(sc.e1, c->stream));
    CK(cudaMemcpyAsync(cells, sc.cells.p, sizeof(float4) * total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, sc.e0, sc.e1));
    return UALM_OK;
}
