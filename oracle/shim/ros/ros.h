// oracle/shim/ros/ros.h -- empty stand-in: back_end/include/utils/se2traj.hpp includes <ros/ros.h> but uses nothing from it.
#pragma once
