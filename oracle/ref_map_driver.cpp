/* oracle/ref_map_driver.cpp -- drives the reference's OWN UnevenMap::constructMap + UnevenMap::filter (uneven_map/src/uneven_map.cpp:317-398,
 * 5-43), compiled UNMODIFIED from /root/reference against oracle/shim (Eigen stand-in with a Jacobi EigenSolver, a bin-search
 * KdTreeFLANN with PCL's radiusSearch / nearestKSearch semantics, no-op ROS).  UnevenMap::init is NOT run (its PCD reader, CropBox and
 * VoxelGrid are PCL's): the members init would set (uneven_map.cpp:96-122, 147-162) are filled here from an already preprocessed cloud,
 * then constructMap() runs as written.  sin / cos come from include/ualm_detmath.h like everywhere else (shim/detmath_redirect.h).
 * TEST INFRASTRUCTURE ONLY: oracle/_ref/librefmap.so, used by tests/test_ref_pin.py::test_reference_constructMap_*. */
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
#undef private
#undef protected

using namespace uneven_planner;

extern "C" int ref_map_construct(const float *pts, int npts, double map_size_x, double map_size_y, double xy_res, double yaw_res, double ex, double ey,
                                 double ez, int iter_num, double *cells_out, int *voxel_num_out)
{
    UnevenMap map;
    map.iter_num = iter_num;
    map.map_size[0] = map_size_x; map.map_size[1] = map_size_y;
    map.ellipsoid_x = ex; map.ellipsoid_y = ey; map.ellipsoid_z = ez;
    map.xy_resolution = xy_res; map.yaw_resolution = yaw_res;
    map.map_file = "/dev/null";
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
    map.occ_r2_buffer = std::vector<char>((size_t)map.voxel_num(0) * (size_t)map.voxel_num(1), 0);
    /* uneven_map.cpp:147-162 with the given cloud */
    map.world_cloud.reset(new pcl::PointCloud<pcl::PointXYZ>());
    map.world_cloud_plane.reset(new pcl::PointCloud<pcl::PointXY>());
    for (int i = 0; i < npts; i++) {
        map.world_cloud->points.push_back(pcl::PointXYZ(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
        pcl::PointXY p;
        p.x = pts[3 * i]; p.y = pts[3 * i + 1];
        map.world_cloud_plane->points.emplace_back(p);
    }
    map.kd_tree.setInputCloud(map.world_cloud);
    map.kd_tree_plane.setInputCloud(map.world_cloud_plane);
    /* the reference prints its progress to stdout */
    std::streambuf *keep = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    map.constructMap();
    std::cout.rdbuf(keep);
    for (int k = 0; k < 3; k++) voxel_num_out[k] = (int)map.voxel_num(k);
    for (int i = 0; i < buffer_size; i++) {
        const RXS2 &r = map.map_buffer[i];
        cells_out[4 * i] = r.z; cells_out[4 * i + 1] = r.sigma; cells_out[4 * i + 2] = r.zb.x(); cells_out[4 * i + 3] = r.zb.y();
    }
    return 0;
}
