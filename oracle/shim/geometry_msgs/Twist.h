#pragma once
#include <std_msgs/Header.h>
#include <memory>
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; typedef std::shared_ptr<const PoseStamped> ConstPtr; };
struct Twist { Vector3 linear, angular; };
struct PoseWithCovariance { Pose pose; };
struct TwistWithCovariance { Twist twist; };
typedef std::shared_ptr<const PoseStamped> PoseStampedConstPtr;
}
