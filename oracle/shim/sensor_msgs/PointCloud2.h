#pragma once
#include <std_msgs/Header.h>
namespace sensor_msgs { struct PointCloud2 { std_msgs::Header header; }; struct PointCloud { std_msgs::Header header; }; }
