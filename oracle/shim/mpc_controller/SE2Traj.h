// oracle/shim/mpc_controller/SE2Traj.h -- the message type catkin would generate from mpc_controller/msg/SE2Traj.msg
// (time start_time; geometry_msgs/Point[] pos_pts, angle_pts; geometry_msgs/Vector3 init_v, init_a; float64[] posT_pts, angleT_pts)
#pragma once
#include <vector>
#include <ros/ros.h>
#include <geometry_msgs/Twist.h>
namespace mpc_controller {
struct SE2Traj {
    ros::Time start_time;
    std::vector<geometry_msgs::Point> pos_pts, angle_pts;
    geometry_msgs::Vector3 init_v, init_a;
    std::vector<double> posT_pts, angleT_pts;
};
}
