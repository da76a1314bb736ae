// oracle/shim/ompl/geometric/SimpleSetup.h -- included by kino_astar.h, nothing of it is used
#pragma once
namespace ompl { namespace geometric {} }
