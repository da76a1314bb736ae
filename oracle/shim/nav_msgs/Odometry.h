#pragma once
#include <nav_msgs/Path.h>
