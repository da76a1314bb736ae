#pragma once
#include <string>
#include <ros/ros.h>
namespace std_msgs { struct Header { ros::Time stamp; std::string frame_id; unsigned seq = 0; }; struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; }; }
