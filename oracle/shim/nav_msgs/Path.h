#pragma once
#include <vector>
#include <geometry_msgs/Twist.h>
namespace nav_msgs {
struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; };
struct Odometry { std_msgs::Header header; geometry_msgs::PoseWithCovariance pose; geometry_msgs::TwistWithCovariance twist; typedef std::shared_ptr<const Odometry> ConstPtr; };
typedef std::shared_ptr<const Odometry> OdometryConstPtr;
}
