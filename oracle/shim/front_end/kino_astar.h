// oracle/shim/front_end/kino_astar.h -- shadows front_end/include/front_end/kino_astar.h (OMPL, not on the optimizeSE2Traj path):
// the back-end only holds a KinoAstar::Ptr and calls plan() from its interactive test callback (alm_traj_opt.cpp:69).
#pragma once
#include <memory>
#include <vector>
#include <Eigen/Eigen>
#include "uneven_map/uneven_map.h"
namespace uneven_planner {
class KinoAstar {
public:
    typedef std::shared_ptr<KinoAstar> Ptr;
    std::vector<Eigen::Vector3d> injected_path;    // the pin tests supply the front-end's polyline (x, y, yaw) here
    void init(ros::NodeHandle &) {}
    void setEnvironment(const UnevenMap::Ptr &) {}
    std::vector<Eigen::Vector3d> plan(const Eigen::Vector3d &, const Eigen::Vector3d &) { return injected_path; }
};
}
