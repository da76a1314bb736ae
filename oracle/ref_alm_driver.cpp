/* oracle/ref_alm_driver.cpp -- drives the reference's OWN ALMTrajOpt, i.e. back_end/src/alm_traj_opt.cpp compiled UNMODIFIED from
 * /root/reference (with back_end/include/back_end/alm_traj_opt.h, utils/{se2traj,banded_system,lbfgs}.hpp and
 * uneven_map/include/uneven_map/uneven_map.h), against oracle/shim: a minimal Eigen stand-in, no-op ROS / message / PCL /
 * OpenCV headers, a stub front_end/kino_astar.h and a stub utils/root_finder.hpp (none of which the optimizeSE2Traj path
 * executes).  Two things the path needs live in uneven_map.cpp, which cannot be compiled here (PCL, Eigen::EigenSolver):
 *   - the map grid itself: filled in from the caller's cells through the (private) members UnevenMap::init would set
 *     (uneven_map.cpp:96-122), using `#define private public` on the reference header;
 *   - UnevenMap::normSO2 (uneven_map.cpp:64-71, six lines) and calYawFromR (odometry callback only): defined below.
 * Everything else that runs -- optimizeSE2Traj, innerCallback, calConstrainCostGrad, initScaling, earlyExit, the dual update,
 * getAllWithGrad / getTerrainWithGradI, MINCO, the banded solver, L-BFGS -- is the reference's own source text.
 * sin / cos / atan2 are taken from include/ualm_detmath.h (macro below) like in the oracle: glibc's and CUDA's libm differ in the
 * last bit, so the oracle, this build and the CUDA path all share that one implementation.
 * TEST INFRASTRUCTURE ONLY: built by `make -C oracle ref` into oracle/_ref/libref.so, used by tests/test_ref_pin.py. */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#define private public
#define protected public
#include "back_end/alm_traj_opt.h"
#undef private
#undef protected

namespace uneven_planner {
void UnevenMap::normSO2(double &yaw) /* uneven_map.cpp:64-71 */
{
    while (yaw < -M_PI) yaw += 2 * M_PI;
    while (yaw > M_PI) yaw -= 2 * M_PI;
}
double UnevenMap::calYawFromR(Eigen::Matrix3d) { return 0.0; } /* odometry callback only */
} // namespace uneven_planner

using namespace uneven_planner;

extern "C" {

struct ref_params_t {   /* same field order as orc_params_t / ualm_params_t */
    double rho_T, rho_ter, max_vel, max_acc_lon, max_acc_lat, max_kap, min_cxi, max_sig;
    int use_scaling;
    double rho, beta, gamma, epsilon_con, max_iter, g_epsilon, min_step, inner_max_iter, delta;
    int mem_size, past, int_K;
    double gravity;
};

/* cells: X*Y*W x 4 doubles {z, sigma, zbx, zby}; geometry as ualm_map_geom_t.  Outputs: ret code, c_xy (6N x 2 col-major),
 * c_yaw (6M), piece durations, lambda[S], mu[6S], hx[S], gx[6S], scale_fx, scale_cx[7S], rho at exit, and feas7 = the reference's own
 * post-solve report {getMaxVxAxAyCurAttSig(getTraj()) [6], getTraj().getNonHolError()} (alm_traj_opt.h:170-229, se2traj.hpp:551-561). */
/* the reference's UnevenMap object, filled in once and reused (building 2.5 M RXS2 cells dominates a single call otherwise) */
void *ref_map_create(const double *cells, const int *voxel_num, const double *origin, const double *max_boundary, double xy_res, double yaw_res,
                     double gravity)
{
    UnevenMap::Ptr *h = new UnevenMap::Ptr(new UnevenMap());
    UnevenMap &map = **h;
    for (int k = 0; k < 3; k++) {
        map.voxel_num(k) = voxel_num[k];
        map.min_boundary(k) = origin[k]; map.map_origin(k) = origin[k]; map.max_boundary(k) = max_boundary[k];
        map.min_idx(k) = 0; map.max_idx(k) = voxel_num[k] - 1;
    }
    map.xy_resolution = xy_res; map.yaw_resolution = yaw_res;
    map.xy_resolution_inv = 1.0 / xy_res; map.yaw_resolution_inv = 1.0 / yaw_res;   /* uneven_map.cpp:104-105 */
    map.gravity = gravity;
    const size_t ncell = (size_t)voxel_num[0] * voxel_num[1] * voxel_num[2];
    map.map_buffer.resize(ncell);
    for (size_t i = 0; i < ncell; i++) map.map_buffer[i] = RXS2(cells[4 * i], cells[4 * i + 1], Eigen::Vector2d(cells[4 * i + 2], cells[4 * i + 3]));
    map.map_ready = true;
    return h;
}
void ref_map_destroy(void *h) { delete (UnevenMap::Ptr *)h; }

int ref_alm_solve_h(const ref_params_t *p, void *map_handle, int N, int M, const double *bnd18, double total_time, const double *inner_xy,
                    const double *inner_yaw, double *c_xy, double *c_yaw, double *piece_T, double *lambda, double *mu, double *hx, double *gx,
                    double *scale_fx, double *scale_cx, double *rho_out, double *feas7)
{
    UnevenMap::Ptr map = *(UnevenMap::Ptr *)map_handle;
    ALMTrajOpt opt;
    opt.rho_T = p->rho_T; opt.rho_ter = p->rho_ter; opt.max_vel = p->max_vel; opt.max_acc_lon = p->max_acc_lon; opt.max_acc_lat = p->max_acc_lat;
    opt.max_kap = p->max_kap; opt.min_cxi = p->min_cxi; opt.max_sig = p->max_sig; opt.use_scaling = p->use_scaling != 0; opt.rho = p->rho;
    opt.beta = p->beta; opt.gamma = p->gamma; opt.epsilon_con = p->epsilon_con; opt.max_iter = p->max_iter; opt.g_epsilon = p->g_epsilon;
    opt.min_step = p->min_step; opt.inner_max_iter = p->inner_max_iter; opt.delta = p->delta; opt.mem_size = p->mem_size; opt.past = p->past;
    opt.int_K = p->int_K; opt.in_test = false; opt.in_debug = false;
    opt.setEnvironment(map);

    Eigen::MatrixXd initXY(2, 3), endXY(2, 3), innerXY(2, N - 1);
    Eigen::VectorXd initYaw(3), endYaw(3), innerYaw(M - 1);
    for (int j = 0; j < 3; j++) for (int d = 0; d < 2; d++) { initXY(d, j) = bnd18[d + 2 * j]; endXY(d, j) = bnd18[6 + d + 2 * j]; }
    for (int j = 0; j < 3; j++) { initYaw(j) = bnd18[12 + j]; endYaw(j) = bnd18[15 + j]; }
    for (int j = 0; j < N - 1; j++) for (int d = 0; d < 2; d++) innerXY(d, j) = inner_xy[d + 2 * (size_t)j];
    for (int j = 0; j < M - 1; j++) innerYaw(j) = inner_yaw[j];
    int ret = opt.optimizeSE2Traj(initXY, endXY, innerXY, initYaw, endYaw, innerYaw, total_time);

    const Eigen::MatrixXd &cxy = opt.minco_se2.pos_minco.getCoeffs();
    const Eigen::MatrixXd &cyaw = opt.minco_se2.yaw_minco.getCoeffs();
    for (int d = 0; d < 2; d++) for (int i = 0; i < 6 * N; i++) c_xy[i + (size_t)d * 6 * N] = cxy(i, d);
    for (int i = 0; i < 6 * M; i++) c_yaw[i] = cyaw(i, 0);
    piece_T[0] = opt.minco_se2.pos_minco.T1(0); piece_T[1] = opt.minco_se2.yaw_minco.T1(0);
    const int S = (int)opt.lambda.size();
    for (int i = 0; i < S; i++) { lambda[i] = opt.lambda(i); hx[i] = opt.hx(i); }
    for (int i = 0; i < 6 * S; i++) { mu[i] = opt.mu(i); gx[i] = opt.gx(i); }
    for (int i = 0; i < 7 * S; i++) scale_cx[i] = opt.scale_cx(i);
    *scale_fx = opt.scale_fx;
    *rho_out = opt.rho;
    if (feas7) {
        SE2Trajectory tr = opt.getTraj();
        std::vector<double> mx = opt.getMaxVxAxAyCurAttSig(tr);
        for (int k = 0; k < 6; k++) feas7[k] = mx[k];
        feas7[6] = tr.getNonHolError();
    }
    return ret;
}

/* UnevenMap::getAllWithGrad of the reference (uneven_map.h:318-377 -> getTerrainWithGradI :258-315, posToIndex / boundIndex / isInMap):
 * values[7], grads[21] (row-major 7 x 3) */
void ref_map_query(void *map_handle, const double *pos, double *values, double *grads)
{
    UnevenMap::Ptr map = *(UnevenMap::Ptr *)map_handle;
    std::vector<double> v;
    std::vector<Eigen::Vector3d> g;
    map->getAllWithGrad(Eigen::Vector3d(pos[0], pos[1], pos[2]), v, g);
    for (int i = 0; i < 7; i++) { values[i] = v[i]; for (int k = 0; k < 3; k++) grads[3 * i + k] = g[i](k); }
}

/* One ALMTrajOpt::calConstrainCostGrad call (alm_traj_opt.cpp:663-991) of the reference at a caller-given decision vector x =
 * [tau | Pxy | Pyaw], multipliers, scales and rho: the object state is prepared the way optimizeSE2Traj (:179-203) and innerCallback
 * (:284-299) prepare it, with the reference's own members and calTfromTau / MINCO_SE2::generate.  Outputs: cost, hx[S], gx[6S] and the
 * (C, T) gradients before the adjoint. */
int ref_alm_constrain(const ref_params_t *p, void *map_handle, int N, int M, const double *bnd18, const double *x, const double *lambda,
                      const double *mu, const double *scale_cx, double scale_fx, double rho, double *cost, double *hx, double *gx, double *gdCxy,
                      double *gdTxy, double *gdCyaw, double *gdTyaw)
{
    UnevenMap::Ptr map = *(UnevenMap::Ptr *)map_handle;
    ALMTrajOpt opt;
    opt.rho_T = p->rho_T; opt.rho_ter = p->rho_ter; opt.max_vel = p->max_vel; opt.max_acc_lon = p->max_acc_lon; opt.max_acc_lat = p->max_acc_lat;
    opt.max_kap = p->max_kap; opt.min_cxi = p->min_cxi; opt.max_sig = p->max_sig; opt.use_scaling = p->use_scaling != 0; opt.rho = rho;
    opt.beta = p->beta; opt.gamma = p->gamma; opt.epsilon_con = p->epsilon_con; opt.max_iter = p->max_iter; opt.g_epsilon = p->g_epsilon;
    opt.min_step = p->min_step; opt.inner_max_iter = p->inner_max_iter; opt.delta = p->delta; opt.mem_size = p->mem_size; opt.past = p->past;
    opt.int_K = p->int_K; opt.in_test = false; opt.in_debug = false;
    opt.setEnvironment(map);
    opt.piece_xy = N; opt.piece_yaw = M; opt.dim_T = 1;
    opt.minco_se2.reset(N, M);
    opt.init_xy.resize(2, 3); opt.end_xy.resize(2, 3); opt.init_yaw.resize(1, 3); opt.end_yaw.resize(1, 3);
    for (int j = 0; j < 3; j++) for (int d = 0; d < 2; d++) { opt.init_xy(d, j) = bnd18[d + 2 * j]; opt.end_xy(d, j) = bnd18[6 + d + 2 * j]; }
    for (int j = 0; j < 3; j++) { opt.init_yaw(0, j) = bnd18[12 + j]; opt.end_yaw(0, j) = bnd18[15 + j]; }
    const int S = N * (p->int_K + 1);
    opt.equal_num = S; opt.non_equal_num = 6 * S;
    opt.hx.resize(S); opt.hx.setZero(); opt.gx.resize(6 * S); opt.gx.setZero();
    opt.lambda.resize(S); opt.mu.resize(6 * S); opt.scale_cx.resize(7 * S);
    for (int i = 0; i < S; i++) opt.lambda(i) = lambda[i];
    for (int i = 0; i < 6 * S; i++) opt.mu(i) = mu[i];
    for (int i = 0; i < 7 * S; i++) opt.scale_cx(i) = scale_cx[i];
    opt.scale_fx = scale_fx;
    Eigen::MatrixXd Pxy(2, N - 1), Pyaw(1, M - 1);
    for (int j = 0; j < N - 1; j++) for (int d = 0; d < 2; d++) Pxy(d, j) = x[1 + d + 2 * (size_t)j];
    for (int j = 0; j < M - 1; j++) Pyaw(0, j) = x[1 + 2 * (N - 1) + j];
    Eigen::VectorXd Txy(N), Tyaw(M);
    opt.calTfromTau(x[0], Txy);
    opt.calTfromTau(x[0], Tyaw);
    opt.minco_se2.generate(opt.init_xy, opt.end_xy, Pxy, Txy, opt.init_yaw, opt.end_yaw, Pyaw, Tyaw);
    double c = 0.0;
    Eigen::MatrixXd gCxy, gCyaw;
    Eigen::VectorXd gTxy, gTyaw;
    opt.calConstrainCostGrad(c, gCxy, gTxy, gCyaw, gTyaw);
    *cost = c;
    for (int i = 0; i < S; i++) hx[i] = opt.hx(i);
    for (int i = 0; i < 6 * S; i++) gx[i] = opt.gx(i);
    for (int d = 0; d < 2; d++) for (int i = 0; i < 6 * N; i++) gdCxy[i + (size_t)d * 6 * N] = gCxy(i, d);
    for (int i = 0; i < N; i++) gdTxy[i] = gTxy(i);
    for (int i = 0; i < 6 * M; i++) gdCyaw[i] = gCyaw(i, 0);
    for (int i = 0; i < M; i++) gdTyaw[i] = gTyaw(i);
    return 0;
}

int ref_alm_solve(const ref_params_t *p, const double *cells, const int *voxel_num, const double *origin, const double *max_boundary,
                  double xy_res, double yaw_res, int N, int M, const double *bnd18, double total_time, const double *inner_xy,
                  const double *inner_yaw, int scaling_only, double *c_xy, double *c_yaw, double *piece_T, double *lambda, double *mu, double *hx,
                  double *gx, double *scale_fx, double *scale_cx, double *rho_out, double *feas7)
{
    (void)scaling_only;
    void *h = ref_map_create(cells, voxel_num, origin, max_boundary, xy_res, yaw_res, p->gravity);
    const int ret = ref_alm_solve_h(p, h, N, M, bnd18, total_time, inner_xy, inner_yaw, c_xy, c_yaw, piece_T, lambda, mu, hx, gx, scale_fx, scale_cx,
                                    rho_out, feas7);
    ref_map_destroy(h);
    return ret;
}

} /* extern "C" */
