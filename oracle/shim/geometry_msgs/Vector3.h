#pragma once
#include <geometry_msgs/Twist.h>
