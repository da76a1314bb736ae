/* oracle/ref_pm_driver.cpp -- drives the reference's OWN PlanManager::rcvWpsCallBack (plan_manager/src/plan_manager.cpp:43-188,
 * compiled unmodified from /root/reference): front-end path -> yaw unwrapping -> arc-length resampling (pm.cpp:62-122) ->
 * ALMTrajOpt::optimizeSE2Traj -> SE2Traj message for the MPC (pm.cpp:151-185).  The front-end is the stub of
 * oracle/shim/front_end/kino_astar.h, whose plan() returns the polyline the test injects; publishers keep the last message
 * (oracle/shim/ros/ros.h).  Map and parameters are filled in as in ref_alm_driver.cpp.  TEST INFRASTRUCTURE ONLY. */
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
#include <string.h>
#include <time.h>

#define private public
#define protected public
#include "plan_manager/plan_manager.h"
#undef private
#undef protected

namespace uneven_planner {
void UnevenMap::init(ros::NodeHandle &) {}   /* uneven_map.cpp (PCL) is not part of this build; PlanManager::init is never called */
} // namespace uneven_planner

using namespace uneven_planner;

extern "C" {

struct ref_params_t {
    double rho_T, rho_ter, max_vel, max_acc_lon, max_acc_lat, max_kap, min_cxi, max_sig;
    int use_scaling;
    double rho, beta, gamma, epsilon_con, max_iter, g_epsilon, min_step, inner_max_iter, delta;
    int mem_size, past, int_K;
    double gravity;
};

/* path: n x 3 (x, y, yaw) row-major.  manager5 = {piece_len, yaw_piece_times, mean_vel, init_time_times, init_sig_vel}.
 * Outputs: counts[2] = {#pos_pts, #angle_pts}; pos_pts (x, y interleaved), posT_pts, angle_pts, angleT_pts (caller-sized). */
int ref_pm_plan(const ref_params_t *p, const double *cells, const int *voxel_num, const double *origin, const double *max_boundary, double xy_res,
                double yaw_res, const double *manager5, const double *path, int n, int *counts, double *pos_pts, double *posT_pts,
                double *angle_pts, double *angleT_pts)
{
    PlanManager pm;
    pm.piece_len = manager5[0]; pm.yaw_piece_times = manager5[1]; pm.mean_vel = manager5[2]; pm.init_time_times = manager5[3];
    pm.init_sig_vel = manager5[4];
    pm.uneven_map.reset(new UnevenMap);
    UnevenMap &map = *pm.uneven_map;
    for (int k = 0; k < 3; k++) {
        map.voxel_num(k) = voxel_num[k];
        map.min_boundary(k) = origin[k]; map.map_origin(k) = origin[k]; map.max_boundary(k) = max_boundary[k];
        map.min_idx(k) = 0; map.max_idx(k) = voxel_num[k] - 1;
    }
    map.xy_resolution = xy_res; map.yaw_resolution = yaw_res;
    map.xy_resolution_inv = 1.0 / xy_res; map.yaw_resolution_inv = 1.0 / yaw_res;
    map.gravity = p->gravity;
    const size_t ncell = (size_t)voxel_num[0] * voxel_num[1] * voxel_num[2];
    map.map_buffer.resize(ncell);
    for (size_t i = 0; i < ncell; i++) map.map_buffer[i] = RXS2(cells[4 * i], cells[4 * i + 1], Eigen::Vector2d(cells[4 * i + 2], cells[4 * i + 3]));
    map.map_ready = true;
    pm.kino_astar.reset(new KinoAstar);
    for (int i = 0; i < n; i++) pm.kino_astar->injected_path.push_back(Eigen::Vector3d(path[3 * i], path[3 * i + 1], path[3 * i + 2]));
    ALMTrajOpt &opt = pm.traj_opt;
    opt.rho_T = p->rho_T; opt.rho_ter = p->rho_ter; opt.max_vel = p->max_vel; opt.max_acc_lon = p->max_acc_lon; opt.max_acc_lat = p->max_acc_lat;
    opt.max_kap = p->max_kap; opt.min_cxi = p->min_cxi; opt.max_sig = p->max_sig; opt.use_scaling = p->use_scaling != 0; opt.rho = p->rho;
    opt.beta = p->beta; opt.gamma = p->gamma; opt.epsilon_con = p->epsilon_con; opt.max_iter = p->max_iter; opt.g_epsilon = p->g_epsilon;
    opt.min_step = p->min_step; opt.inner_max_iter = p->inner_max_iter; opt.delta = p->delta; opt.mem_size = p->mem_size; opt.past = p->past;
    opt.int_K = p->int_K; opt.in_test = false; opt.in_debug = false;
    opt.setFrontend(pm.kino_astar);
    opt.setEnvironment(pm.uneven_map);
    pm.odom_pos = Eigen::Vector3d(path[0], path[1], path[2]);

    ros::shim_last_message<mpc_controller::SE2Traj>() = mpc_controller::SE2Traj();
    geometry_msgs::PoseStamped goal;
    goal.pose.position.x = path[3 * (n - 1)]; goal.pose.position.y = path[3 * (n - 1) + 1];
    goal.pose.orientation.w = 1.0;
    pm.rcvWpsCallBack(goal);

    const mpc_controller::SE2Traj &msg = ros::shim_last_message<mpc_controller::SE2Traj>();
    counts[0] = (int)msg.pos_pts.size(); counts[1] = (int)msg.angle_pts.size();
    for (size_t i = 0; i < msg.pos_pts.size(); i++) { pos_pts[2 * i] = msg.pos_pts[i].x; pos_pts[2 * i + 1] = msg.pos_pts[i].y; }
    for (size_t i = 0; i < msg.posT_pts.size(); i++) posT_pts[i] = msg.posT_pts[i];
    for (size_t i = 0; i < msg.angle_pts.size(); i++) angle_pts[i] = msg.angle_pts[i].x;
    for (size_t i = 0; i < msg.angleT_pts.size(); i++) angleT_pts[i] = msg.angleT_pts[i];
    return 0;
}

} /* extern "C" */
