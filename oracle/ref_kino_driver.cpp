/* oracle/ref_kino_driver.cpp -- drives the reference's OWN KinoAstar::plan (front_end/src/kino_astar.cpp:67-236 with the helpers of
 * front_end/include/front_end/kino_astar.h), compiled UNMODIFIED from /root/reference against oracle/shim.  The UnevenMap it searches is
 * filled here from a given cell grid (the members UnevenMap::init would set, uneven_map.cpp:96-122) and the occupancy rule of
 * uneven_map.cpp:169-179; OMPL's DubinsStateSpace is the shim's (oracle/shim/ompl, = csrc/dubins.h).  Plain libm trigonometry on both
 * sides (this is host code on both sides: no detmath redirect).  TEST INFRASTRUCTURE ONLY: oracle/_ref/librefkino.so, used by
 * tests/test_ref_pin.py::test_kino_astar_*. */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#define private public
#define protected public
#include "uneven_map/uneven_map.h"
#include "front_end/kino_astar.h"
#undef private
#undef protected

using namespace uneven_planner;

extern "C" int ref_kino_plan(const double *cells, double map_size_x, double map_size_y, double xy_res, double yaw_res, double min_cnormal, double max_rho,
                             const double *kp /* 13 values in the order of ualm_astar_params_t */, const double *start, const double *goal, double *path_out,
                             int max_pts, unsigned char *occ3_out, unsigned char *occ2_out)
{
    UnevenMap::Ptr mp(new UnevenMap());
    UnevenMap &map = *mp;
    map.map_size[0] = map_size_x; map.map_size[1] = map_size_y;
    map.xy_resolution = xy_res; map.yaw_resolution = yaw_res;
    map.min_cnormal = min_cnormal; map.max_rho = max_rho;
    /* uneven_map.cpp:96-122 */
    map.map_size[2] = 2.0 * M_PI + 5e-2;
    map.min_boundary = -map.map_size / 2.0;
    map.max_boundary = map.map_size / 2.0;
    map.map_origin = map.min_boundary;
    map.xy_resolution_inv = 1.0 / map.xy_resolution;
    map.yaw_resolution_inv = 1.0 / map.yaw_resolution;
    map.voxel_num(0) = ceil(map.map_size(0) / map.xy_resolution);
    map.voxel_num(1) = ceil(map.map_size(1) / map.xy_resolution);
    map.voxel_num(2) = ceil(map.map_size(2) / map.yaw_resolution);
    map.min_idx = Eigen::Vector3i::Zero();
    map.max_idx = map.voxel_num - Eigen::Vector3i::Ones();
    const int buffer_size = (int)map.voxel_num(0) * (int)map.voxel_num(1) * (int)map.voxel_num(2);
    map.map_buffer = std::vector<RXS2>(buffer_size, RXS2());
    map.c_buffer = std::vector<double>(buffer_size, 1.0);
    map.occ_buffer = std::vector<char>(buffer_size, 0);
    map.occ_r2_buffer = std::vector<char>(map.getXYNum(), 0);
    for (int i = 0; i < buffer_size; i++) {
        map.map_buffer[i] = RXS2(cells[4 * i], cells[4 * i + 1], Eigen::Vector2d(cells[4 * i + 2], cells[4 * i + 3]));
        map.c_buffer[i] = map.map_buffer[i].getC();          /* uneven_map.cpp:385, 390 */
    }
    /* occ map, uneven_map.cpp:169-179 */
    for (int x = 0; x < map.voxel_num[0]; x++)
        for (int y = 0; y < map.voxel_num[1]; y++)
            for (int yaw = 0; yaw < map.voxel_num[2]; yaw++)
                if (map.c_buffer[map.toAddress(x, y, yaw)] < map.min_cnormal || map.map_buffer[map.toAddress(x, y, yaw)].sigma > map.max_rho) {
                    map.occ_buffer[map.toAddress(x, y, yaw)] = 1;
                    map.occ_r2_buffer[x * map.voxel_num(1) + y] = 1;
                }
    map.map_ready = true;
    if (occ3_out) std::memcpy(occ3_out, map.occ_buffer.data(), buffer_size);
    if (occ2_out) std::memcpy(occ2_out, map.occ_r2_buffer.data(), map.getXYNum());

    KinoAstar ka;
    /* KinoAstar::init, kino_astar.cpp:5-44, without the node handle */
    ka.yaw_resolution = kp[0]; ka.lambda_heu = kp[1]; ka.weight_r2 = kp[2]; ka.weight_so2 = kp[3]; ka.weight_v_change = kp[4];
    ka.weight_delta_change = kp[5]; ka.weight_sigma = kp[6]; ka.time_interval = kp[7]; ka.collision_interval = kp[8]; ka.oneshot_range = kp[9];
    ka.wheel_base = kp[10]; ka.max_steer = kp[11]; ka.max_vel = kp[12];
    ka.in_test = false;
    ka.yaw_resolution_inv = 1.0 / ka.yaw_resolution;
    ka.shot_finder = std::make_shared<ompl::base::DubinsStateSpace>(ka.wheel_base / tan(ka.max_steer));
    ka.setEnvironment(mp);
    std::streambuf *keep = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    std::vector<Eigen::Vector3d> path = ka.plan(Eigen::Vector3d(start[0], start[1], start[2]), Eigen::Vector3d(goal[0], goal[1], goal[2]));
    std::cout.rdbuf(keep);
    if ((int)path.size() > max_pts) return -1;
    for (size_t i = 0; i < path.size(); i++) { path_out[3 * i] = path[i](0); path_out[3 * i + 1] = path[i](1); path_out[3 * i + 2] = path[i](2); }
    return (int)path.size();
}
