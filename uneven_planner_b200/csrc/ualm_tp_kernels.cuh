// ualm_tp_kernels.cuh -- the THROUGHPUT path of the batched MINCO / PHR-ALM / L-BFGS optimizer (precision 32 and 65 of include/ualm.h).
//
// Same algorithm as the parity path (ualm_kernels.cuh; reference: back_end/src/alm_traj_opt.cpp:168-347, 349-661, 663-991,
// utils/lbfgs.hpp:276-722, utils/se2traj.hpp:595-816, uneven_map.h:258-377), laid out for throughput instead of for bit parity:
//
//   * LOCKSTEP EVALUATION ROUNDS with continuous batching.  A pool of trajectory slots holds every problem in flight (several
//     batches).  One round = one cost/gradient evaluation of every active trajectory, as three bulk kernels:
//       ka_kernel  one WARP per trajectory: finishes the previous evaluation (adjoint banded solve, gradient assembly), runs the
//                  per-trajectory control flow up to the next evaluation request (Lewis-Overton line search, L-BFGS two-loop on the
//                  coalesced [slot][n] history, ALM dual update / convergence) and the MINCO forward solve for the new point;
//       ks_kernel  (rounds after an admission only) initScaling, one CTA per new trajectory, one thread per constraint sample;
//       kb_kernel  the per-constraint-sample penalty cost + gradient (calConstrainCostGrad): one CTA per trajectory, one thread per
//                  sample, the UnevenMap tiles of the pieces staged into shared memory by TMA (cp.async.bulk.tensor + mbarrier),
//                  duals / scales / constraint values streamed as coalesced SoA arrays, gradients reduced onto the control
//                  points in shared memory.
//     Finished trajectories leave the active list, new ones join between rounds: no trajectory waits for a batch mate.
//   * The MINCO system is NONDIMENSIONALISED: piece durations are uniform (alm_traj_opt.h:257-261), so A(T) = R(T) A(1) C(T) with
//     diagonal R, C; A(1) depends only on the piece count and is LU-factored ONCE per piece count at engine creation.  An
//     evaluation only runs the triangular sweeps.  initScaling uses the waypoint rows of A(1)^-T (dense, precomputed per piece
//     count) and one extra forward solve instead of one adjoint solve per constraint.
//   * Mixed precision: ka_kernel (short serial chains, latency bound) always computes in double; the penalty kernel computes
//     in R = float (precision 32) or double (precision 65); duals, constraint values, exchanged gradients and the L-BFGS history are
//     stored in R.
// FMA contraction is on, libm sincos/atan2 are CUDA's, reductions are re-associated: results are NOT bit-comparable with the oracle;
// tests/test_gpu_tp.py bounds the difference per evaluation and reports the end-to-end distribution (BASELINE config 5's sweep).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ualm.h"

namespace ualm_tp {

#define TP_NMAX 64
#define TP_MMAX 128
#define TP_NVAR 256                     // >= 1 + 2 (NMAX - 1) + (MMAX - 1)
#define TP_CSTRIDE (12 * TP_NMAX + 6 * TP_MMAX)   // per-slot coefficient block: c_xy (6N x 2 column-major) at 0, c_yaw at 12 * TP_NMAX
#define TP_CYAW (12 * TP_NMAX)
#define TP_TSTRIDE (TP_NMAX + TP_MMAX)  // per-slot time-gradient block: gdT_xy at 0, gdT_yaw at TP_NMAX
#define TP_NDUAL 21                     // dual block fields (each S long): 0 lambda | 1..6 mu | 7..13 scale_cx | 14 hx | 15..20 gx
#define TP_FW 14                        // doubles per LU factor row: 13 band entries + 1 / diagonal
#define TP_FPAD 6                       // zero rows before and after each factor table
#define TP_KA_WARPS 4
#define TP_KB_THREADS 128
#define TP_TILE 8                       // map tile = 8 x 8 cells x 8 yaw layers of float4 = 8 KB
#define TP_TILE_BYTES (TP_TILE * TP_TILE * TP_TILE * 16)
#define TP_MAXPPC 7                     // pieces per sample chunk (<= TP_KB_THREADS / (K + 1))
#define TP_MAX_TICKETS 64
#define TP_SUBGROUPS 4                  // every batch in flight is split into up to this many independently advancing groups
#define TP_MAXLANES 16                  // batches in flight (= UALM_MAX_LANES of ualm_api.cu)
#define TP_NGROUPS (TP_MAXLANES * TP_SUBGROUPS)   // lanes x subgroups: each group has its own stream, active list and round sequence

#define TP_DELTA_SIGL 0.01
#define TP_CUR_SCALE 10.0
#define TP_SIG_SCALE 1000.0
#define TP_SCALE_TRICK_JERK 1000.0

enum { PH_FREE = 0, PH_NEW, PH_REQ_FIRST, PH_REQ_LS, PH_REQ_EVALONLY, PH_DONE };

// lbfgs return codes (lbfgs.hpp:135-184)
enum {
    LB_CONVERGENCE = 0, LB_STOP, LB_CANCELED,
    LBERR_UNKNOWNERROR = -1024, LBERR_INVALID_N, LBERR_INVALID_MEMSIZE, LBERR_INVALID_GEPSILON, LBERR_INVALID_TESTPERIOD, LBERR_INVALID_DELTA,
    LBERR_INVALID_MINSTEP, LBERR_INVALID_MAXSTEP, LBERR_INVALID_FDECCOEFF, LBERR_INVALID_SCURVCOEFF, LBERR_INVALID_MACHINEPREC,
    LBERR_INVALID_MAXLINESEARCH, LBERR_INVALID_FUNCVAL, LBERR_MINIMUMSTEP, LBERR_MAXIMUMSTEP, LBERR_MAXIMUMLINESEARCH, LBERR_MAXIMUMITERATION,
    LBERR_WIDTHTOOSMALL, LBERR_INVALIDPARAMETERS, LBERR_INCREASEGRADIENT,
};

struct TpParams {
    double rho_T, rho_ter, max_vel, max_acc_lon, max_acc_lat, max_kap, min_cxi, max_sig;
    int use_scaling;
    double rho, beta, gamma, epsilon_con, max_iter, g_epsilon, min_step, delta;
    int inner_max_iter, mem_size, past, int_K;
    double gravity;
};

struct TpMap {
    const float4 *cells;   // {z, sigma, zbx, zby}
    int vn[3];
    double origin[3], maxb[3], xy_res, yaw_res, xy_inv, yaw_inv;
};

struct TpState {
    int phase, N, M, n, S, ticket, index, need_scale;
    int mode, group, padm1, padm2;   // 0 = solve, 1 = one evaluation at the given duals (kernel-level parity), 2 = initScaling only
    int n_evals, iters_total, outer_iter, last_ret, ret_code, max_bound, sum_bound, pad0;
    int k, end, bound, ls_count, brackt, touched, pad1, pad2;
    double fx, step, stp, ls_mu, ls_nu, dginit, finit, dgtest, dstest;
    double pf[16];
    double rho, scale_fx, inner_cost, res_h, res_g;
    double tau, T, Tx, Ty, jerk_raw;          // of the last forward solve (the state getTraj() / the dual update see)
    double f_last;
    double bnd[18];
};

struct TpPool {
    int capacity, m, K, Smax, use_tma, ka_col_bytes, ka_hist_stride, ka_hist_bytes;
    TpState *st;
    int *active;            // [TP_NGROUPS][capacity] active slot list per group
    int *n_active;          // [TP_NGROUPS]
    int *remaining;         // per ticket: trajectories not yet done
    double *vec;            // [cap][5][TP_NVAR]  x | g | xp | gp | d
    double *cd;             // [cap][TP_CSTRIDE]  coefficients of the last forward solve (double)
    double *gw;             // [cap][TP_CSTRIDE]  adjoint workspace / the z vectors of initScaling
    void *cr;               // [cap][TP_CSTRIDE]  R copy of the coefficients for the sample kernels (== cd when R is double)
    void *gdc;              // [cap][TP_CSTRIDE]  R  constraint part of dcost/dc from kb_kernel
    void *gdt;              // [cap][TP_TSTRIDE]  R  constraint part of dcost/dT
    double *kb_cost;        // [cap]              constraint cost
    void *dual;             // [cap][TP_NDUAL * Smax]  R  (field f of slot s at s * 21 * Smax + f * S_slot)
    void *hs, *hy;          // [cap][m][TP_NVAR]  R  L-BFGS history, vector j of slot s at (s * m + j) * TP_NVAR
    double *lm_ys, *lm_alpha;   // [cap][m]
    const double *lu;       // factor tables of A(1): table of P pieces at lu_off[P] (row 0; TP_FPAD zero rows on both sides)
    const int *lu_off;      // [TP_MMAX + 1]
    const void *wway;       // R  waypoint rows of A(1)^-T: (P - 1) x 6P at wway_off[P]
    const long long *wway_off;
    long long *prof;        // developer profile: [16] SM-cycle sums per ka phase (lane 0 of every warp), or null
};
enum { KP_FIN_PRE = 0, KP_FIN_SWEEP, KP_FIN_POST, KP_ADV_LS, KP_ADV_POST, KP_TWOLOOP, KP_ALM, KP_FWD_PRE, KP_FWD_SWEEP, KP_FWD_POST, KP_SCALEZ, KP_WARPS, KP_N };
struct KProf {
    long long *p; long long last;
    __device__ __forceinline__ void start(long long *pp, int lane) { p = lane == 0 ? pp : nullptr; if (p) last = clock64(); }
    __device__ __forceinline__ void mark(int ph) { if (p) { const long long c = clock64(); atomicAdd((unsigned long long *)&p[ph], (unsigned long long)(c - last)); last = c; } }
};

// ---------------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double expC2(double tau) { return tau > 0.0 ? ((0.5 * tau + 1.0) * tau + 1.0) : 1.0 / ((0.5 * tau - 1.0) * tau + 1.0); } // alm_traj_opt.h:232-235
__device__ __forceinline__ double dTdtau(double tau)   // getTtoTauGrad, alm_traj_opt.h:244-253
{
    if (tau > 0) return tau + 1.0;
    const double den = (0.5 * tau - 1.0) * tau + 1.0;
    return (1.0 - tau) / (den * den);
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <class T>
__device__ __forceinline__ double vdot(const double *a, const T *b, int n, int lane)
{
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += a[i] * (double)b[i];
    return warp_sum(s);
}
__device__ __forceinline__ double vabsmax(const double *a, int n, int lane)
{
    double m = 0.0;
    for (int i = lane; i < n; i += 32) m = fmax(m, fabs(a[i]));
    return warp_max(m);
}

// derivative order of row r of the MINCO system of P pieces (se2traj.hpp:609-674): the row scale of A(T) = R(T) A(1) C(T) is T^-ord
__device__ __host__ __forceinline__ int row_ord(int r, int n6)
{
    if (r < 3) return r;
    if (r >= n6 - 3) return r - (n6 - 3);
    const int tt = (r - 3) % 6;
    return tt == 0 ? 3 : tt == 1 ? 4 : tt == 4 ? 1 : tt == 5 ? 2 : 0;
}

// entry A(r, c) of the P-piece system at T = 1 (se2traj.hpp:609-674)
__device__ __host__ inline double a1_entry(int P, int r, int c)
{
    const int n6 = 6 * P;
    if (c < 0 || c >= n6 || r < 0 || r >= n6) return 0.0;
    if (r < 3) return c == r ? (r == 2 ? 2.0 : 1.0) : 0.0;
    if (r >= n6 - 3) {
        const int e = c - (n6 - 6), tr = r - (n6 - 3);
        if (e < 0) return 0.0;
        const double pos[6] = {1, 1, 1, 1, 1, 1}, vel[6] = {0, 1, 2, 3, 4, 5}, acc[6] = {0, 0, 2, 6, 12, 20};
        return tr == 0 ? pos[e] : tr == 1 ? vel[e] : acc[e];
    }
    const int i = (r - 3) / 6, tt = (r - 3) - 6 * i, e = c - 6 * i;
    if (e < 0 || e > 11) return 0.0;
    switch (tt) {
    case 0: return e == 3 ? 6.0 : e == 4 ? 24.0 : e == 5 ? 60.0 : e == 9 ? -6.0 : 0.0;          // jerk continuity
    case 1: return e == 4 ? 24.0 : e == 5 ? 120.0 : e == 10 ? -24.0 : 0.0;                     // snap continuity
    case 2: return e <= 5 ? 1.0 : 0.0;                                                          // waypoint
    case 3: return e <= 5 ? 1.0 : e == 6 ? -1.0 : 0.0;                                          // position continuity
    case 4: return (e >= 1 && e <= 5) ? (double)e : e == 7 ? -1.0 : 0.0;                        // velocity continuity
    default: return e == 2 ? 2.0 : e == 3 ? 6.0 : e == 4 ? 12.0 : e == 5 ? 20.0 : e == 8 ? -2.0 : 0.0;   // acceleration continuity
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// engine set-up: LU of A(1) for every piece count (banded_system.hpp:66-91, no pivoting), and the waypoint rows of A(1)^-T
// ---------------------------------------------------------------------------------------------------------------------
__global__ void lu_tables_kernel(double *lu, const int *lu_off, int Pmax)
{
    const int P = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (P > Pmax) return;
    const int n6 = 6 * P;
    double *F = lu + lu_off[P];
    for (int r = -TP_FPAD; r < n6 + TP_FPAD; r++)
        for (int q = 0; q < TP_FW; q++) F[r * TP_FW + q] = (r >= 0 && r < n6 && q < 13) ? a1_entry(P, r, r - 6 + q) : 0.0;
    for (int k = 0; k < n6; k++) {
        const double piv = F[k * TP_FW + 6];
        const int last = min(k + 6, n6 - 1);
        for (int i = k + 1; i <= last; i++) {
            double &l = F[i * TP_FW + 6 - (i - k)];
            if (l == 0.0) continue;
            l = l / piv;
            for (int j = k + 1; j <= last; j++) {
                const double u = F[k * TP_FW + 6 + (j - k)];
                if (u != 0.0) F[i * TP_FW + 6 + (j - i)] -= l * u;
            }
        }
    }
    for (int r = 0; r < n6; r++) F[r * TP_FW + 13] = 1.0 / F[r * TP_FW + 6];
}

// A(1) x = b in place (one thread, plain loops; set-up only)
__device__ inline void solve_a1_serial(const double *F, int n6, double *b)
{
    for (int r = 0; r < n6; r++) {
        double v = b[r];
        for (int d = 1; d <= 6 && r - d >= 0; d++) v -= F[r * TP_FW + 6 - d] * b[r - d];
        b[r] = v;
    }
    for (int r = n6 - 1; r >= 0; r--) {
        double v = b[r];
        for (int d = 1; d <= 6 && r + d < n6; d++) v -= F[r * TP_FW + 6 + d] * b[r + d];
        b[r] = v * F[r * TP_FW + 13];
    }
}
// Wway[P][w][c] = (A(1)^-T)(6w+5, c) = (A(1)^-1 e_{6w+5})(c)
template <class R>
__global__ void wway_tables_kernel(const double *lu, const int *lu_off, R *wway, const long long *wway_off, double *scratch, int Pmax)
{
    const int P = blockIdx.y + 2;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (P > Pmax || w >= P - 1) return;
    const int n6 = 6 * P;
    double *b = scratch + ((size_t)(blockIdx.y * gridDim.x * blockDim.x) + w) * (6 * TP_MMAX);
    for (int r = 0; r < n6; r++) b[r] = 0.0;
    b[6 * w + 5] = 1.0;
    solve_a1_serial(lu + lu_off[P], n6, b);
    R *o = wway + wway_off[P] + (size_t)w * n6;
    for (int r = 0; r < n6; r++) o[r] = (R)b[r];
}

// ---------------------------------------------------------------------------------------------------------------------
// triangular sweeps of the prefactored A(1): lanes 0 / 1 solve the x / y columns of the xy system, lane 2 the yaw column,
// in lockstep (rolling 6-entry register window, static indices after unrolling)
// ---------------------------------------------------------------------------------------------------------------------
// One pass over the prefactored A(1) of both systems.  The factor rows stream through a per-warp shared-memory ring by cp.async
// (all 32 lanes stage; prefetch distance two blocks of six rows), lanes 0 / 1 / 2 run the dependent chains of the x / y / yaw
// columns with a rolling six-entry register window (static indices after unrolling); the right-hand sides of the next block are
// fetched while the current block is solved, so no global-memory latency sits on the chain.
//   KIND 0: L y = b      ascending,   L(r, r-d)  = F[r][6-d]
//   KIND 1: U x = y      descending,  U(r, r+d)  = F[r][6+d], then * F[r][13] (= 1 / U(r, r))
//   KIND 2: U^T y = b    ascending,   U(r-d, r)  = F[r-d][6+d], then * F[r][13]
//   KIND 3: L^T x = y    descending,  L(r+d, r)  = F[r+d][6-d]
#define TP_RING 5
#define TP_BLK (6 * TP_FW)
// Static non-zero structure of the LU factors of the MINCO matrix (symbolic elimination; period 6 in the row index, the last
// block -- tail position / velocity / acceleration rows -- is dense for the L-based kinds): bit d-1 of the mask of row type
// t = r mod 6 says whether the term with offset d can be non-zero.  Supersets of the numeric pattern, so skipped terms are exact zeros.
__host__ __device__ constexpr unsigned long long tp_pack6(int a, int b, int c, int d, int e, int f)
{
    return (unsigned long long)a | ((unsigned long long)b << 6) | ((unsigned long long)c << 12) | ((unsigned long long)d << 18) |
           ((unsigned long long)e << 24) | ((unsigned long long)f << 30);
}
__host__ __device__ constexpr int tp_sweep_mask(int kind, int t)
{
    return (int)(((kind == 0 ? tp_pack6(0x3f, 0x3e, 0x3c, 0x00, 0x00, 0x1f)      // L rows
                 : kind == 1 ? tp_pack6(0x0c, 0x06, 0x03, 0x23, 0x21, 0x18)      // U rows
                 : kind == 2 ? tp_pack6(0x00, 0x00, 0x00, 0x2f, 0x3f, 0x03)      // U columns
                             : tp_pack6(0x30, 0x38, 0x3c, 0x1e, 0x0f, 0x07))     // L columns
                  >> (6 * t)) & 0x3f);
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NP>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NP) : "memory"); }

struct SweepSys {
    const double *Fxy, *Fyaw;   // row 0 of the factor tables
    int N, M;                   // blocks (= pieces) of the xy / yaw system
    double *col;                // this lane's column (lanes 0, 1: xy; lane 2: yaw), else unused
    double *ring;               // this warp's ring: [2][TP_RING][TP_BLK]
};

// six rows of one block.  Dependent-chain discipline (an fp64 FMA that waits for its operand costs a full pipeline latency): per
// row only ONE FMA may wait for the previous row's result -- terms with older neighbours (d = 3..6) are summed first on two
// accumulators, then d = 2, and the newest neighbour (d = 1) enters last; divisions are folded into pre-scaled factors off the chain.
template <int KIND, bool FULL>
__device__ __forceinline__ void block_rows(const double *f, const double *g, bool hasg, const double (&rhs)[6], double (&w)[6], double (&out)[6])
{
    constexpr bool ASC = (KIND == 0 || KIND == 2);
#pragma unroll
    for (int tt = 0; tt < 6; tt++) {
        const int t = ASC ? tt : 5 - tt;
        const int mask = FULL ? 0x3f : tp_sweep_mask(KIND, t);
        double fs[7];
#pragma unroll
        for (int d = 1; d <= 6; d++) {
            double fv = 0.0;
            if ((mask >> (d - 1)) & 1) {
                if (KIND == 0) fv = f[t * TP_FW + 6 - d];
                else if (KIND == 1) fv = f[t * TP_FW + 6 + d];
                else if (KIND == 2) fv = (t - d >= 0) ? f[(t - d) * TP_FW + 6 + d] : (hasg ? g[(t - d + 6) * TP_FW + 6 + d] : 0.0);
                else fv = (t + d <= 5) ? f[(t + d) * TP_FW + 6 - d] : (hasg ? g[(t + d - 6) * TP_FW + 6 - d] : 0.0);
            }
            fs[d] = fv;
        }
        double r0 = rhs[t];
        if (KIND == 1 || KIND == 2) {
            const double rd = f[t * TP_FW + 13];
            r0 *= rd;
#pragma unroll
            for (int d = 1; d <= 6; d++) if ((mask >> (d - 1)) & 1) fs[d] *= rd;
        }
#define TP_W(d) (ASC ? w[(t - (d) + 6) % 6] : w[(t + (d)) % 6])
#define TP_TERM(d) (((mask >> ((d) - 1)) & 1) ? fs[d] * TP_W(d) : 0.0)
        const double pa = r0 - TP_TERM(6) - TP_TERM(4);
        const double pb = -TP_TERM(5) - TP_TERM(3);
        double v = (pa + pb) - TP_TERM(2);
        v -= TP_TERM(1);
#undef TP_TERM
#undef TP_W
        w[t] = v;
        out[t] = v;
    }
}

template <int KIND>
__device__ __forceinline__ void sweep_pass(const SweepSys &s, int lane)
{
    constexpr bool ASC = (KIND == 0 || KIND == 2);
    const int nb = max(s.N, s.M);
    const int sys = lane == 2 ? 1 : 0, P = sys ? s.M : s.N;
    const bool consumer = lane < 3;
    const unsigned ring0 = (unsigned)__cvta_generic_to_shared(s.ring);
    auto issue = [&](int c) {
        if (c < s.N) {
            const int blk = ASC ? c : s.N - 1 - c;
            const double *src = s.Fxy + (size_t)blk * TP_BLK;
            const unsigned dst = ring0 + 8u * (unsigned)((c % TP_RING) * TP_BLK);
            for (int q = lane; q < TP_BLK / 2; q += 32) cp_async16(dst + 16u * q, src + 2 * q);
        }
        if (c < s.M) {
            const int blk = ASC ? c : s.M - 1 - c;
            const double *src = s.Fyaw + (size_t)blk * TP_BLK;
            const unsigned dst = ring0 + 8u * (unsigned)((TP_RING + c % TP_RING) * TP_BLK);
            for (int q = lane; q < TP_BLK / 2; q += 32) cp_async16(dst + 16u * q, src + 2 * q);
        }
        cp_async_commit();
    };
    const double *myring = s.ring + (size_t)sys * TP_RING * TP_BLK;
    double w[6] = {0, 0, 0, 0, 0, 0}, rhs[6] = {0, 0, 0, 0, 0, 0}, nrhs[6] = {0, 0, 0, 0, 0, 0};
    issue(0);
    issue(1);
    issue(2);
    if (consumer && 0 < P) {
        const int blk = ASC ? 0 : P - 1;
#pragma unroll
        for (int t = 0; t < 6; t++) rhs[t] = s.col[6 * blk + t];
    }
    for (int c = 0; c < nb; c++) {
        issue(c + 3);
        cp_async_wait<3>();
        __syncwarp();
        const bool on = consumer && c < P;
        const int blk = ASC ? c : P - 1 - c;
        if (consumer && c + 1 < P) {          // right-hand sides of the next block
            const int nblk = ASC ? c + 1 : P - 2 - c;
#pragma unroll
            for (int t = 0; t < 6; t++) nrhs[t] = s.col[6 * nblk + t];
        }
        if (on) {
            const double *f = myring + (size_t)(c % TP_RING) * TP_BLK;               // this block's six factor rows
            const double *g = myring + (size_t)((c + TP_RING - 1) % TP_RING) * TP_BLK;   // the block processed just before
            const bool hasg = c > 0;
            double out[6];
            if ((KIND == 0 || KIND == 3) && blk == P - 1) block_rows<KIND, true>(f, g, hasg, rhs, w, out);
            else block_rows<KIND, false>(f, g, hasg, rhs, w, out);
#pragma unroll
            for (int t = 0; t < 6; t++) s.col[6 * blk + t] = out[t];
        }
#pragma unroll
        for (int t = 0; t < 6; t++) rhs[t] = nrhs[t];
        __syncwarp();
    }
    cp_async_wait<0>();
    __syncwarp();
}
// A(1) x = b and A(1)^T x = b, in place
__device__ __forceinline__ void sweep_forward(const SweepSys &s, int lane) { sweep_pass<0>(s, lane); sweep_pass<1>(s, lane); }
__device__ __forceinline__ void sweep_adjoint(const SweepSys &s, int lane) { sweep_pass<2>(s, lane); sweep_pass<3>(s, lane); }

// jerk energy of one piece and its T-derivative (se2traj.hpp:702-707, 739-744); a = first column block, b = second (or null)
__device__ __forceinline__ void jerk_piece(const double *a, const double *b, double T1, double T2, double T3, double T4, double T5, double &e, double &gt)
{
    double d33 = a[3] * a[3], d43 = a[4] * a[3], d44 = a[4] * a[4], d53 = a[5] * a[3], d54 = a[5] * a[4], d55 = a[5] * a[5];
    if (b) { d33 += b[3] * b[3]; d43 += b[4] * b[3]; d44 += b[4] * b[4]; d53 += b[5] * b[3]; d54 += b[5] * b[4]; d55 += b[5] * b[5]; }
    e = 36.0 * d33 * T1 + 144.0 * d43 * T2 + 192.0 * d44 * T3 + 240.0 * d53 * T3 + 720.0 * d54 * T4 + 720.0 * d55 * T5;
    gt = 36.0 * d33 + 288.0 * d43 * T1 + 576.0 * d44 * T2 + 720.0 * d53 * T2 + 2880.0 * d54 * T3 + 3600.0 * d55 * T4;
}
__device__ __forceinline__ double jerk_gc(const double *c6, int k, double T1, double T2, double T3, double T4, double T5)   // se2traj.hpp:719-737
{
    const double c3 = c6[3], c4 = c6[4], c5 = c6[5];
    if (k == 5) return 240.0 * c3 * T3 + 720.0 * c4 * T4 + 1440.0 * c5 * T5;
    if (k == 4) return 144.0 * c3 * T2 + 384.0 * c4 * T3 + 720.0 * c5 * T4;
    if (k == 3) return 72.0 * c3 * T1 + 144.0 * c4 * T2 + 240.0 * c5 * T3;
    return 0.0;
}
// the B1 / B2 vectors of calGradCTtoQT (se2traj.hpp:763-814): coefficient k of row j of piece i's time-gradient contraction
__device__ __forceinline__ void time_b(const double *cc, double T1, double T2, double T3, double T4, double &nv, double &na, double &nj, double &ns, double &nc)
{
    nv = -(cc[1] + 2.0 * T1 * cc[2] + 3.0 * T2 * cc[3] + 4.0 * T3 * cc[4] + 5.0 * T4 * cc[5]);
    na = -(2.0 * cc[2] + 6.0 * T1 * cc[3] + 12.0 * T2 * cc[4] + 20.0 * T3 * cc[5]);
    nj = -(6.0 * cc[3] + 24.0 * T1 * cc[4] + 60.0 * T2 * cc[5]);
    ns = -(24.0 * cc[4] + 120.0 * T1 * cc[5]);
    nc = -120.0 * cc[5];
}

struct SlotView {
    TpState *st;
    double *x, *g, *xp, *gp, *d;
    double *cd, *gw;
    const double *Fxy, *Fyaw;
    double *ring;              // this warp's factor ring in shared memory
    double *sm;                // this warp's column buffer in shared memory (12N + 6M doubles; aliased by the two-loop's history ring)
    int N, M, n, S, slot;
};

__device__ __forceinline__ SlotView slot_view(const TpPool &E, int slot, double *sm, double *ring)
{
    SlotView v;
    v.slot = slot;
    v.sm = sm;
    v.ring = ring;
    v.st = E.st + slot;
    v.N = v.st->N; v.M = v.st->M; v.n = v.st->n; v.S = v.st->S;
    double *vb = E.vec + (size_t)slot * 5 * TP_NVAR;
    v.x = vb; v.g = vb + TP_NVAR; v.xp = vb + 2 * TP_NVAR; v.gp = vb + 3 * TP_NVAR; v.d = vb + 4 * TP_NVAR;
    v.cd = E.cd + (size_t)slot * TP_CSTRIDE;
    v.gw = E.gw + (size_t)slot * TP_CSTRIDE;
    v.Fxy = E.lu + E.lu_off[v.N];
    v.Fyaw = E.lu + E.lu_off[v.M];
    return v;
}

// The three columns live in the warp's shared-memory column buffer during a solve: x at [0, 6N), y at [6N, 12N), yaw at [12N, 12N + 6M)
// (global coefficient blocks keep the yaw column at TP_CYAW).  sweeps in place on v.sm.
__device__ __forceinline__ void solve_sm(const SlotView &v, bool adjoint, int lane)
{
    SweepSys sc;
    sc.Fxy = v.Fxy; sc.Fyaw = v.Fyaw; sc.N = v.N; sc.M = v.M; sc.ring = v.ring;
    sc.col = v.sm + (lane == 2 ? 12 * v.N : (lane == 1 ? 6 * v.N : 0));
    __syncwarp();
    if (adjoint) sweep_adjoint(sc, lane); else sweep_forward(sc, lane);
    __syncwarp();
}

// MINCO forward for the decision vector in v.x (alm_traj_opt.cpp:293-299 + se2traj.hpp:595-680, nondimensionalised):
// leaves the coefficients in v.cd (and their R copy), T / Tx / Ty / jerk_raw in the state
template <class R>
__device__ void minco_forward(const TpPool &E, const SlotView &v, int lane, KProf &kp)
{
    TpState *st = v.st;
    const int N = v.N, M = v.M, nx = 6 * N, ny = 6 * M;
    const double tau = v.x[0];
    const double T = expC2(tau), Tx = T / (double)N, Ty = T / (double)M;
    double *sm = v.sm;
    for (int q = lane; q < 2 * nx + ny; q += 32) sm[q] = 0.0;
    __syncwarp();
    const double *bnd = st->bnd;
    const double *Pxy = v.x + 1, *Pyaw = v.x + 1 + 2 * (N - 1);
    if (lane < 2) {       // head / tail position, velocity, acceleration rows scaled by T^ord (se2traj.hpp:615-617, 672-674)
        const int d = lane;
        sm[0 + d * nx] = bnd[d]; sm[1 + d * nx] = bnd[d + 2] * Tx; sm[2 + d * nx] = bnd[d + 4] * Tx * Tx;
        sm[nx - 3 + d * nx] = bnd[6 + d]; sm[nx - 2 + d * nx] = bnd[6 + d + 2] * Tx; sm[nx - 1 + d * nx] = bnd[6 + d + 4] * Tx * Tx;
    }
    if (lane == 2) {
        double *y = sm + 2 * nx;
        y[0] = bnd[12]; y[1] = bnd[13] * Ty; y[2] = bnd[14] * Ty * Ty;
        y[ny - 3] = bnd[15]; y[ny - 2] = bnd[16] * Ty; y[ny - 1] = bnd[17] * Ty * Ty;
    }
    for (int i = lane; i < N - 1; i += 32) { sm[6 * i + 5] = Pxy[2 * i]; sm[6 * i + 5 + nx] = Pxy[2 * i + 1]; }
    for (int i = lane; i < M - 1; i += 32) sm[2 * nx + 6 * i + 5] = Pyaw[i];
    kp.mark(KP_FWD_PRE);
    solve_sm(v, false, lane);
    kp.mark(KP_FWD_SWEEP);
    // c_k = c^_k T^-k: one lane per 6-coefficient block; out to global (double + the sample kernels' R copy); jerk energy on the way
    // (se2traj.hpp:697-710, 852-855: per piece 36 c3^2 T + 144 c3 c4 T^2 + 192 c4^2 T^3 + 240 c3 c5 T^3 + 720 c4 c5 T^4 + 720 c5^2 T^5)
    double ix[6], iy[6];
    ix[0] = iy[0] = 1.0;
#pragma unroll
    for (int k = 1; k < 6; k++) { ix[k] = ix[k - 1] / Tx; iy[k] = iy[k - 1] / Ty; }
    const double X1 = Tx, X2 = Tx * Tx, X3 = X2 * Tx, X4 = X2 * X2, X5 = X4 * Tx;
    const double Y1 = Ty, Y2 = Ty * Ty, Y3 = Y2 * Ty, Y4 = Y2 * Y2, Y5 = Y4 * Ty;
    double *c = v.cd;
    R *cr = (R *)E.cr + (size_t)v.slot * TP_CSTRIDE;
    double e = 0.0;
    for (int blk = lane; blk < 2 * N + M; blk += 32) {
        const bool isy = blk >= 2 * N;
        const int off = isy ? TP_CYAW + 6 * (blk - 2 * N) : 6 * blk;
        double c6[6];
#pragma unroll
        for (int k = 0; k < 6; k++) c6[k] = sm[6 * blk + k] * (isy ? iy[k] : ix[k]);
#pragma unroll
        for (int k = 0; k < 6; k++) c[off + k] = c6[k];
        if (sizeof(R) != sizeof(double)) {
#pragma unroll
            for (int k = 0; k < 6; k++) cr[off + k] = (R)c6[k];
        }
        const double T1 = isy ? Y1 : X1, T2 = isy ? Y2 : X2, T3 = isy ? Y3 : X3, T4 = isy ? Y4 : X4, T5 = isy ? Y5 : X5;
        e += 36.0 * c6[3] * c6[3] * T1 + 144.0 * c6[4] * c6[3] * T2 + 192.0 * c6[4] * c6[4] * T3 + 240.0 * c6[5] * c6[3] * T3 + 720.0 * c6[5] * c6[4] * T4 +
             720.0 * c6[5] * c6[5] * T5;
    }
    e = warp_sum(e);
    if (lane == 0) { st->tau = tau; st->T = T; st->Tx = Tx; st->Ty = Ty; st->jerk_raw = e; }
    __syncwarp();
    kp.mark(KP_FWD_POST);
}

// initScaling support: z = A(T)^-1 u with u the B1 / B2 contraction vectors of calGradCTtoQT, so that for any dcost/dc vector g
// sum_i (B . adjoint)(i) = z . g  (one extra forward solve instead of one adjoint solve per constraint).  z -> v.gw
__device__ void scaling_z(const SlotView &v, int lane)
{
    const TpState *st = v.st;
    const int N = v.N, M = v.M, nx = 6 * N, ny = 6 * M;
    const double Tx = st->Tx, Ty = st->Ty;
    double *sm = v.sm;
    const double *c = v.cd;
    for (int q = lane; q < 2 * nx + ny; q += 32) sm[q] = 0.0;
    __syncwarp();
    for (int q = lane; q < 2 * N + M; q += 32) {
        const bool isy = q >= 2 * N;
        const int P = isy ? M : N, i = isy ? q - 2 * N : q % N, d = isy ? 0 : q / N;
        const double T1 = isy ? Ty : Tx, T2 = T1 * T1, T3 = T2 * T1, T4 = T2 * T2;
        const double *cc = c + (isy ? TP_CYAW : d * nx) + 6 * i;
        double *u = sm + (isy ? 2 * nx : d * nx);
        double nv, na, nj, ns, nc;
        time_b(cc, T1, T2, T3, T4, nv, na, nj, ns, nc);
        // rows 6i+3..6i+8 (orders 3, 4, 0, 0, 1, 2) resp. the last three rows (orders 0, 1, 2); right-hand side scaled by T^ord
        if (i < P - 1) {
            u[6 * i + 3] = ns * T3; u[6 * i + 4] = nc * T4; u[6 * i + 5] = nv; u[6 * i + 6] = nv; u[6 * i + 7] = na * T1; u[6 * i + 8] = nj * T2;
        } else {
            u[6 * P - 3] = nv; u[6 * P - 2] = na * T1; u[6 * P - 1] = nj * T2;
        }
    }
    solve_sm(v, false, lane);
    double ix[6], iy[6];
    ix[0] = iy[0] = 1.0;
#pragma unroll
    for (int k = 1; k < 6; k++) { ix[k] = ix[k - 1] / Tx; iy[k] = iy[k - 1] / Ty; }
    double *z = v.gw;
    for (int q = lane; q < 2 * nx; q += 32) z[q] = sm[q] * ix[q % 6];
    for (int q = lane; q < ny; q += 32) z[TP_CYAW + q] = sm[2 * nx + q] * iy[q % 6];
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------------
// ka_kernel pieces
// ---------------------------------------------------------------------------------------------------------------------
// finish the evaluation whose sample part kb_kernel left in gdc / gdt / kb_cost: f -> st->f_last, gradient -> v.g
// (alm_traj_opt.cpp:318-346, se2traj.hpp:751-816)
template <class R>
__device__ void finish_eval(const TpPool &E, const TpParams &p, const SlotView &v, int lane, KProf &kp)
{
    TpState *st = v.st;
    const int N = v.N, M = v.M, nx = 6 * N, ny = 6 * M;
    const double Tx = st->Tx, Ty = st->Ty, scale_fx = st->scale_fx;
    const double js = (p.use_scaling ? TP_SCALE_TRICK_JERK : 1.0) * scale_fx;
    const double X1 = Tx, X2 = Tx * Tx, X3 = X2 * Tx, X4 = X2 * X2, X5 = X4 * Tx;
    const double Y1 = Ty, Y2 = Ty * Ty, Y3 = Y2 * Ty, Y4 = Y2 * Y2, Y5 = Y4 * Ty;
    const double *c = v.cd;
    double *sm = v.sm;
    const R *gdc = (const R *)E.gdc + (size_t)v.slot * TP_CSTRIDE;
    const R *gdt = (const R *)E.gdt + (size_t)v.slot * TP_TSTRIDE;
    double ix[6], iy[6];
    ix[0] = iy[0] = 1.0;
#pragma unroll
    for (int k = 1; k < 6; k++) { ix[k] = ix[k - 1] / Tx; iy[k] = iy[k - 1] / Ty; }
    // dcost/dc = jerk part + constraint part, scaled by T^-k for the nondimensional adjoint (alm_traj_opt.cpp:322-332);
    // one lane per 6-coefficient block: the jerk gradient needs c3..c5 of the block only (se2traj.hpp:719-737)
    for (int blk = lane; blk < 2 * N + M; blk += 32) {
        const bool isy = blk >= 2 * N;
        const int off = isy ? TP_CYAW + 6 * (blk - 2 * N) : 6 * blk;
        const double T1 = isy ? Y1 : X1, T2 = isy ? Y2 : X2, T3 = isy ? Y3 : X3, T4 = isy ? Y4 : X4, T5 = isy ? Y5 : X5;
        const double c3 = c[off + 3], c4 = c[off + 4], c5 = c[off + 5];
        double g6[6];
#pragma unroll
        for (int k = 0; k < 6; k++) g6[k] = (double)gdc[off + k];
        g6[3] += (72.0 * c3 * T1 + 144.0 * c4 * T2 + 240.0 * c5 * T3) * js;
        g6[4] += (144.0 * c3 * T2 + 384.0 * c4 * T3 + 720.0 * c5 * T4) * js;
        g6[5] += (240.0 * c3 * T3 + 720.0 * c4 * T4 + 1440.0 * c5 * T5) * js;
#pragma unroll
        for (int k = 0; k < 6; k++) sm[6 * blk + k] = g6[k] * (isy ? iy[k] : ix[k]);
    }
    kp.mark(KP_FIN_PRE);
    solve_sm(v, true, lane);
    kp.mark(KP_FIN_SWEEP);
    // adjoint row r of the dimensional system = w_r T^ord(r).  Time gradients (se2traj.hpp:763-814) and their sums
    double sx = 0.0, sy = 0.0;
    for (int q = lane; q < N + M; q += 32) {
        const bool isy = q >= N;
        const int P = isy ? M : N, i = isy ? q - N : q;
        const double T1 = isy ? Ty : Tx, T2 = T1 * T1, T3 = T2 * T1, T4 = T2 * T2, T5 = T4 * T1;
        double e, gj;
        jerk_piece(c + (isy ? TP_CYAW : 0) + 6 * i, isy ? nullptr : c + nx + 6 * i, T1, T2, T3, T4, T5, e, gj);
        double gt = gj * js + (double)gdt[isy ? TP_NMAX + i : i];
        for (int d = 0; d < (isy ? 1 : 2); d++) {
            const double *cc = c + (isy ? TP_CYAW : d * nx) + 6 * i;
            const double *a = sm + (isy ? 2 * nx : d * nx);
            double nv, na, nj, ns, nc;
            time_b(cc, T1, T2, T3, T4, nv, na, nj, ns, nc);
            if (i < P - 1) gt += ns * T3 * a[6 * i + 3] + nc * T4 * a[6 * i + 4] + nv * (a[6 * i + 5] + a[6 * i + 6]) + na * T1 * a[6 * i + 7] + nj * T2 * a[6 * i + 8];
            else gt += nv * a[6 * P - 3] + na * T1 * a[6 * P - 2] + nj * T2 * a[6 * P - 1];
        }
        if (isy) sy += gt; else sx += gt;
    }
    sx = warp_sum(sx); sy = warp_sum(sy);
    for (int i = lane; i < N - 1; i += 32) { v.g[1 + 2 * i] = sm[6 * i + 5]; v.g[2 + 2 * i] = sm[6 * i + 5 + nx]; }
    for (int i = lane; i < M - 1; i += 32) v.g[1 + 2 * (N - 1) + i] = sm[2 * nx + 6 * i + 5];
    if (lane == 0) {
        const double tau = st->tau;
        v.g[0] = (p.rho_T * scale_fx + sx / (double)N + sy / (double)M) * dTdtau(tau);
        st->f_last = st->jerk_raw * js + E.kb_cost[v.slot] + p.rho_T * st->T * scale_fx;
    }
    __syncwarp();
    kp.mark(KP_FIN_POST);
}

template <class H>
__device__ __forceinline__ H *hist(void *base, const TpPool &E, int slot, int j) { return (H *)base + ((size_t)slot * E.m + j) * TP_NVAR; }

// start of a line search from (x, fx, g, d, step): lbfgs.hpp:276-316.  Returns 0 and leaves the first trial point in v.x, or the
// (negative) error code
__device__ int ls_begin(const SlotView &v, int lane)
{
    TpState *st = v.st;
    const int n = v.n;
    for (int q = lane; q < n; q += 32) { v.xp[q] = v.x[q]; v.gp[q] = v.g[q]; }
    __syncwarp();
    const double dginit = vdot(v.gp, v.d, n, lane);
    const double stp = st->step;
    if (!(stp > 0.0)) return LBERR_INVALIDPARAMETERS;
    if (0.0 < dginit) return LBERR_INCREASEGRADIENT;
    if (lane == 0) {
        st->ls_count = 0; st->brackt = 0; st->touched = 0; st->stp = stp; st->ls_mu = 0.0; st->ls_nu = 1.0e20;
        st->dginit = dginit; st->finit = st->fx; st->dgtest = 1.0e-4 * dginit; st->dstest = 0.9 * dginit;
    }
    for (int q = lane; q < n; q += 32) v.x[q] = v.xp[q] + stp * v.d[q];
    __syncwarp();
    return 0;
}

// one ka step for a trajectory whose evaluation at v.x has just been finished: runs the optimizer's control flow up to the next
// evaluation request (returns the new phase) or to the end of the solve (PH_DONE).  alm_traj_opt.cpp:234-271, lbfgs.hpp:439-722
template <class R>
__device__ int advance(const TpPool &E, const TpParams &p, const SlotView &v, int lane, KProf &kp)
{
    TpState *st = v.st;
    const int n = v.n, m = p.mem_size;
    int ph = st->phase;
    const double fnew = st->f_last;
    int lret = 0;                  // lbfgs_optimize return code once it ends
    bool ended = false;
    if (ph == PH_REQ_FIRST) {      // lbfgs.hpp:523-550
        if (lane == 0) { st->fx = fnew; st->pf[0] = fnew; }
        for (int q = lane; q < n; q += 32) v.d[q] = -v.g[q];
        __syncwarp();
        const double gn = vabsmax(v.g, n, lane), xn = vabsmax(v.x, n, lane);
        if (gn / fmax(1.0, xn) < p.g_epsilon) { lret = LB_CONVERGENCE; ended = true; }
        else {
            const double dd = vdot(v.d, v.d, n, lane);
            if (lane == 0) { st->step = 1.0 / sqrt(dd); st->k = 1; st->end = 0; st->bound = 0; }
            __syncwarp();
            const int r = ls_begin(v, lane);
            if (r < 0) { lret = r; ended = true; for (int q = lane; q < n; q += 32) { v.x[q] = v.xp[q]; v.g[q] = v.gp[q]; } }
            else return PH_REQ_LS;
        }
    } else {                       // PH_REQ_LS: line_search_lewisoverton after the evaluation of a trial point (lbfgs.hpp:318-388)
        const double fx = fnew, finit = st->finit;
        int count = st->ls_count + 1;
        double stp = st->stp, mu = st->ls_mu, nu = st->ls_nu;
        int brackt = st->brackt, touched = st->touched;
        int ls = 0;
        bool done = false;
        if (isinf(fx) || isnan(fx)) { ls = LBERR_INVALID_FUNCVAL; done = true; }
        else if (p.past > 0 && fabs(finit - fx) / (fabs(finit) + 1.0) < p.delta / (double)p.past) { ls = count; done = true; }   // lbfgs.hpp:327-330
        else {
            if (fx > finit + stp * st->dgtest) { nu = stp; brackt = 1; }
            else {
                const double dg = vdot(v.g, v.d, n, lane);
                if (dg < st->dstest) mu = stp;
                else { ls = count; done = true; }
            }
            if (!done) {
                if (64 <= count) { ls = LBERR_MAXIMUMLINESEARCH; done = true; }
                else if (brackt && (nu - mu) < 1.0e-16 * nu) { ls = LBERR_WIDTHTOOSMALL; done = true; }
                else {
                    stp = brackt ? 0.5 * (mu + nu) : stp * 2.0;
                    if (stp < p.min_step) { ls = LBERR_MINIMUMSTEP; done = true; }
                    else if (stp > 1.0e20) {
                        if (touched) { ls = LBERR_MAXIMUMSTEP; done = true; }
                        else { touched = 1; stp = 1.0e20; }
                    }
                }
            }
        }
        if (!done) {               // next trial point
            if (lane == 0) { st->ls_count = count; st->stp = stp; st->ls_mu = mu; st->ls_nu = nu; st->brackt = brackt; st->touched = touched; }
            for (int q = lane; q < n; q += 32) v.x[q] = v.xp[q] + stp * v.d[q];
            __syncwarp();
            return PH_REQ_LS;
        }
        if (lane == 0) { st->fx = fx; st->step = stp; }
        __syncwarp();
        if (ls < 0) {              // revert to the previous point (lbfgs.hpp:575-582)
            for (int q = lane; q < n; q += 32) { v.x[q] = v.xp[q]; v.g[q] = v.gp[q]; }
            __syncwarp();
            lret = ls; ended = true;
        } else {
            int k = st->k;
            if (lane == 0) st->iters_total++;
            if (k > 1000) { lret = LB_CANCELED; ended = true; }        // earlyExit, alm_traj_opt.cpp:1016
            if (!ended) {
                const double gn = vabsmax(v.g, n, lane), xn = vabsmax(v.x, n, lane);
                if (gn / fmax(1.0, xn) < p.g_epsilon) { lret = LB_CONVERGENCE; ended = true; }
            }
            if (!ended && p.past > 0) {
                if (p.past <= k) {
                    const double rate = fabs(st->pf[k % p.past] - fx) / fmax(1.0, fabs(fx));
                    if (rate < p.delta) { lret = LB_STOP; ended = true; }
                }
                if (!ended) { __syncwarp(); if (lane == 0) st->pf[k % p.past] = fx; __syncwarp(); }
            }
            if (!ended && p.inner_max_iter != 0 && p.inner_max_iter <= k) { lret = LBERR_MAXIMUMITERATION; ended = true; }
            if (!ended) {          // history update, cautious test, two-loop recursion (lbfgs.hpp:640-711)
                ++k;
                int end = st->end, bound = st->bound;
                R *sE = hist<R>(E.hs, E, v.slot, end), *yE = hist<R>(E.hy, E, v.slot, end);
                double ys = 0.0, yy = 0.0, ss = 0.0, gpn = 0.0;
                for (int q = lane; q < n; q += 32) {
                    const double s_ = v.x[q] - v.xp[q], y_ = v.g[q] - v.gp[q];
                    sE[q] = (R)s_; yE[q] = (R)y_;
                    v.d[q] = -v.g[q];
                    ys += y_ * s_; yy += y_ * y_; ss += s_ * s_; gpn += v.gp[q] * v.gp[q];
                }
                ys = warp_sum(ys); yy = warp_sum(yy); ss = warp_sum(ss); gpn = warp_sum(gpn);
                double *lys = E.lm_ys + (size_t)v.slot * E.m, *lal = E.lm_alpha + (size_t)v.slot * E.m;
                if (lane == 0) lys[end] = 1.0 / ys;        // only ever used as a divisor: keep the reciprocal
                __syncwarp();
                kp.mark(KP_ADV_POST);
                if (ys > ss * sqrt(gpn) * 1.0e-6) {
                    ++bound;
                    bound = m < bound ? m : bound;
                    end = (end + 1) % m;
                    // Each lane owns elements lane, lane + 32, ... of d in registers.  The 2 * bound history steps (newest -> oldest,
                    // then oldest -> newest) stream through a four-slot shared-memory ring by cp.async, three steps ahead of the
                    // dependent chain (dot product -> warp reduction -> axpy), so no global-memory latency sits on it.
                    constexpr int NR = TP_NVAR / 32;
                    constexpr int HR = 4;
                    const int hstride = E.ka_hist_stride;                 // elements per vector slot (n rounded up to 16 bytes)
                    R *hring = (R *)v.sm;                                 // [HR][2][hstride], aliases the column buffer (idle here)
                    const unsigned hring0 = (unsigned)__cvta_generic_to_shared(hring);
                    const int nchunk = (n * (int)sizeof(R) + 15) / 16;
                    const int nsteps = 2 * bound, ne = (n + 31) >> 5;
                    // 1 / (y_j . s_j) of the steps and the alphas of the first loop sit in shared memory behind the ring (no global
                    // load on the chain)
                    double *rys_s = (double *)((char *)v.sm + E.ka_hist_bytes), *al_s = rys_s + m;
                    auto jof = [&](int t) { return t < bound ? (end + m - 1 - t % m + m) % m : (end - bound + (t - bound) + 2 * m) % m; };
                    auto hissue = [&](int t) {
                        if (t < nsteps) {
                            const int j = jof(t);
                            const R *sj = hist<R>(E.hs, E, v.slot, j), *yj = hist<R>(E.hy, E, v.slot, j);
                            const unsigned dst = hring0 + (unsigned)((t % HR) * 2 * hstride * (int)sizeof(R));
                            for (int q = lane; q < nchunk; q += 32) {
                                cp_async16(dst + 16u * q, (const char *)sj + 16 * q);
                                cp_async16(dst + (unsigned)(hstride * (int)sizeof(R)) + 16u * q, (const char *)yj + 16 * q);
                            }
                        }
                        cp_async_commit();
                    };
                    double dreg[NR];
#pragma unroll
                    for (int e = 0; e < NR; e++) { const int q = lane + 32 * e; dreg[e] = q < n ? v.d[q] : 0.0; }
                    for (int t = lane; t < bound; t += 32) rys_s[t] = lys[(end + m - 1 - t % m + m) % m];
                    __syncwarp();
                    hissue(0); hissue(1); hissue(2);
                    const double scl = ys / yy;
                    for (int t = 0; t < nsteps; t++) {
                        hissue(t + 3);
                        cp_async_wait<3>();
                        __syncwarp();
                        const int j = jof(t);
                        const R *sv = hring + (size_t)(t % HR) * 2 * hstride, *yv = sv + hstride;
                        const bool first = t < bound;
                        const int tf = first ? t : nsteps - 1 - t;       // the first-loop step that handled this history vector
                        const double rysj = rys_s[tf];
                        const double alj = first ? 0.0 : al_s[tf];
                        double pacc = 0.0;
                        double ax[NR];
#pragma unroll
                        for (int e = 0; e < NR; e++) {
                            ax[e] = 0.0;
                            if (e < ne) {                 // warp-uniform: only the 32-element groups the vector really has
                                const int q = lane + 32 * e;
                                const double sq = q < n ? (double)sv[q] : 0.0, yq = q < n ? (double)yv[q] : 0.0;
                                pacc += (first ? sq : yq) * dreg[e];
                                ax[e] = first ? yq : sq;
                            }
                        }
                        pacc = warp_sum(pacc);
                        double cf;
                        if (first) { cf = -(pacc * rysj); if (lane == 0) al_s[tf] = -cf; }
                        else cf = alj - pacc * rysj;
#pragma unroll
                        for (int e = 0; e < NR; e++) if (e < ne) dreg[e] += cf * ax[e];
                        if (t == bound - 1) {
#pragma unroll
                            for (int e = 0; e < NR; e++) dreg[e] *= scl;
                        }
                        __syncwarp();
                    }
                    cp_async_wait<0>();
                    __syncwarp();
#pragma unroll
                    for (int e = 0; e < NR; e++) { const int q = lane + 32 * e; if (q < n) v.d[q] = dreg[e]; }
                    if (lane == 0) { if (bound > st->max_bound) st->max_bound = bound; st->sum_bound += bound; }
                }
                if (lane == 0) { st->k = k; st->end = end; st->bound = bound; st->step = 1.0; }
                __syncwarp();
                kp.mark(KP_TWOLOOP);
                const int r = ls_begin(v, lane);
                if (r < 0) { lret = r; ended = true; for (int q = lane; q < n; q += 32) { v.x[q] = v.xp[q]; v.g[q] = v.gp[q]; } }
                else return PH_REQ_LS;
            }
        }
    }
    // ---- lbfgs_optimize ended with lret: the ALM loop (alm_traj_opt.cpp:236-270) ----
    __syncwarp();
    kp.mark(KP_ADV_LS);
    if (lane == 0) { st->inner_cost = st->fx; st->last_ret = lret; }
    if (!(lret == LB_CONVERGENCE || lret == LB_CANCELED || lret == LB_STOP || lret == LBERR_MAXIMUMITERATION || lret == LBERR_MAXIMUMLINESEARCH)) {
        if (lane == 0) st->ret_code = 1;
        return PH_DONE;
    }
    // updateDualVars with hx / gx of the LAST evaluation (Q1), judgeConvergence with the updated rho (alm_traj_opt.h:132-151)
    {
        const int S = v.S;
        R *du = (R *)E.dual + (size_t)v.slot * TP_NDUAL * E.Smax;
        const double rho = st->rho, rho_new = fmin((1.0 + p.gamma) * rho, p.beta);
        double mh = 0.0, mg = 0.0;
        for (int q = lane; q < S; q += 32) {
            const double h = (double)du[14 * S + q];
            du[q] = (R)((double)du[q] + rho * h);
            mh = fmax(mh, fabs(h));
        }
        for (int q = lane; q < 6 * S; q += 32) {
            const double gq = (double)du[15 * S + q];
            const double mq = fmax((double)du[S + q] + rho * gq, 0.0);
            du[S + q] = (R)mq;
            mg = fmax(mg, fabs(fmax(gq, -mq / rho_new)));
        }
        mh = warp_max(mh); mg = warp_max(mg);
        int iter = st->outer_iter;
        if (lane == 0) { st->rho = rho_new; st->res_h = mh; st->res_g = mg; }
        if (fmax(mh, mg) < p.epsilon_con) return PH_DONE;
        ++iter;
        if (lane == 0) st->outer_iter = iter;
        if ((double)iter > p.max_iter) { if (lane == 0) st->ret_code = 2; return PH_DONE; }
    }
    __syncwarp();
    return PH_REQ_FIRST;           // next lbfgs_optimize starts with an evaluation at the current x
}

template <class R>
__global__ void __launch_bounds__(32 * TP_KA_WARPS) ka_kernel(const __grid_constant__ TpPool E, const __grid_constant__ TpParams p, int group)
{
    const int lane = threadIdx.x & 31, widx = blockIdx.x * TP_KA_WARPS + (threadIdx.x >> 5);
    if (widx >= E.n_active[group]) return;
    extern __shared__ __align__(16) unsigned char ka_smem[];     // per warp: column buffer (ka_col_bytes) | factor ring
    unsigned char *wbase = ka_smem + (size_t)(threadIdx.x >> 5) * (E.ka_col_bytes + 2 * TP_RING * TP_BLK * 8);
    const int slot = E.active[(size_t)group * E.capacity + widx];
    SlotView v = slot_view(E, slot, (double *)wbase, (double *)(wbase + E.ka_col_bytes));
    TpState *st = v.st;
    const int ph = st->phase;
    if (ph == PH_DONE || ph == PH_FREE) return;
    KProf kp;
    kp.start(E.prof, lane);
    if (kp.p) atomicAdd((unsigned long long *)&kp.p[KP_WARPS], 1ull);
    int next;
    if (ph == PH_NEW) {
        const int mode = st->mode;
        if (lane == 0) {
            if (mode != 1) { st->rho = p.rho; st->scale_fx = 1.0; }
            st->outer_iter = 0; st->n_evals = 0; st->iters_total = 0; st->ret_code = 0; st->last_ret = 0;
            st->max_bound = 0; st->sum_bound = 0; st->inner_cost = 0.0; st->res_h = 0.0; st->res_g = 0.0;
        }
        next = mode == 0 ? PH_REQ_FIRST : PH_REQ_EVALONLY;
    } else {
        finish_eval<R>(E, p, v, lane, kp);
        if (lane == 0) st->n_evals++;
        __syncwarp();
        next = (ph == PH_REQ_EVALONLY) ? PH_DONE : advance<R>(E, p, v, lane, kp);
        kp.mark(next == PH_REQ_LS ? KP_ADV_LS : KP_ALM);
    }
    __syncwarp();
    if (next == PH_DONE) {
        if (lane == 0) { st->phase = PH_DONE; atomicSub(E.remaining + st->ticket, 1); }
        return;
    }
    minco_forward<R>(E, v, lane, kp);
    if (ph == PH_NEW && st->need_scale) { scaling_z(v, lane); kp.mark(KP_SCALEZ); }
    if (lane == 0) st->phase = next;
}

// rebuild the active list of one group from the slot phases (one CTA); slots in PH_NEW join here
__global__ void compact_kernel(const __grid_constant__ TpPool E, int hi, int group)
{
    __shared__ int s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    int *list = E.active + (size_t)group * E.capacity;
    for (int s0 = 0; s0 < hi; s0 += blockDim.x) {
        const int s = s0 + threadIdx.x;
        bool live = false;
        if (s < hi) { const int ph = E.st[s].phase; live = ph != PH_FREE && ph != PH_DONE && E.st[s].group == group; }
        const unsigned bal = __ballot_sync(0xffffffffu, live);
        int base = 0;
        if ((threadIdx.x & 31) == 0 && bal) base = atomicAdd(&s_count, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (live) list[base + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u))] = s;
        __syncthreads();
    }
    if (threadIdx.x == 0) E.n_active[group] = s_count;
}

// ---------------------------------------------------------------------------------------------------------------------
// admission: problem b of a submitted batch -> slot
// ---------------------------------------------------------------------------------------------------------------------
struct AdmitDesc { int slot, N, M, ticket, index, mode, group, pad1; long long off_x; double total_time; double bnd[18]; };

template <class R>
__global__ void admit_kernel(const __grid_constant__ TpPool E, const AdmitDesc *ad, const double *x0, int B, int use_scaling, int K,
                             const double *lam, const double *mu, const double *scx, const double *sfx, double rho_eval, const long long *off_s)
{
    const int b = blockIdx.x;
    if (b >= B) return;
    const AdmitDesc a = ad[b];
    TpState *st = E.st + a.slot;
    const int n = 1 + 2 * (a.N - 1) + (a.M - 1), S = a.N * (K + 1);
    double *x = E.vec + (size_t)a.slot * 5 * TP_NVAR;
    for (int q = threadIdx.x; q < n; q += blockDim.x) x[q] = x0[a.off_x + q];
    R *du = (R *)E.dual + (size_t)a.slot * TP_NDUAL * E.Smax;
    if (a.mode != 1) {             // solve: duals 0, scales 1 (alm_traj_opt.cpp:193-203)
        for (int q = threadIdx.x; q < TP_NDUAL * S; q += blockDim.x) du[q] = (q >= 7 * S && q < 14 * S) ? (R)1.0 : (R)0.0;
    } else {                       // single evaluation at caller-provided duals (reference index layout: mu[6 s + t], scale_cx[7 s + t])
        const long long os = off_s[b];
        for (int q = threadIdx.x; q < S; q += blockDim.x) {
            du[q] = (R)(lam ? lam[os + q] : 0.0);
            for (int t = 0; t < 6; t++) du[(1 + t) * S + q] = (R)(mu ? mu[6 * (os + q) + t] : 0.0);
            for (int t = 0; t < 7; t++) du[(7 + t) * S + q] = (R)(scx ? scx[7 * (os + q) + t] : 1.0);
            for (int t = 0; t < 7; t++) du[(14 + t) * S + q] = (R)0.0;
        }
    }
    if (threadIdx.x == 0) {
        st->N = a.N; st->M = a.M; st->n = n; st->S = S; st->ticket = a.ticket; st->index = a.index;
        for (int k = 0; k < 18; k++) st->bnd[k] = a.bnd[k];
        st->mode = a.mode; st->group = a.group;
        st->need_scale = ((a.mode == 0 && use_scaling) || a.mode == 2) ? 1 : 0;
        st->phase = PH_NEW;
        if (a.mode == 1) { st->rho = rho_eval; st->scale_fx = sfx ? sfx[b] : 1.0; }
    }
}

} // namespace ualm_tp
