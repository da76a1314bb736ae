// oracle/shim/ros/ros.h -- stand-in for the few ROS names the reference's back-end translation unit mentions (node handle,
// publishers, time, log macros).  Nothing here does anything: the pin build only calls optimizeSE2Traj and its callees, never
// init(nh) / the callbacks / the visualisation.  TEST INFRASTRUCTURE ONLY (oracle/shim/Eigen/Eigen explains the purpose).
#pragma once
#include <iostream>
#include <sstream>
#include <string>
#include <memory>
#include <boost_shim.h>
namespace ros {
struct Duration { Duration(double = 0.0) {} double toSec() const { return 0.0; } };
struct Time { static Time now() { return Time(); } double toSec() const { return 0.0; } Duration operator-(const Time &) const { return Duration(); } };
struct TimerEvent {};
struct Timer {};
// publish() keeps the last message of every type so the pin tests can read what the reference would have sent
template <class M> M &shim_last_message() { static M m; return m; }
struct Publisher { template <class M> void publish(const M &m) const { shim_last_message<M>() = m; } };
struct Subscriber {};
struct Rate { Rate(double) {} void sleep() {} };
struct TransportHints { TransportHints &tcpNoDelay() { return *this; } };
struct NodeHandle {
    NodeHandle() {}
    NodeHandle(const std::string &) {}
    template <class T> bool getParam(const std::string &, T &) const { return false; }
    template <class T, class D> bool param(const std::string &, T &, const D &) const { return false; }
    template <class M> Publisher advertise(const std::string &, int, bool = false) { return Publisher(); }
    template <class M, class C, class A> Subscriber subscribe(const std::string &, int, void (C::*)(A), C *, const TransportHints & = TransportHints()) { return Subscriber(); }
    template <class C> Timer createTimer(Duration, void (C::*)(const TimerEvent &), C *) { return Timer(); }
};
inline bool ok() { return true; }
} // namespace ros
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { } while (0)
#define ROS_INFO_STREAM(x) do { std::ostringstream ros_shim_os; ros_shim_os << x; } while (0)
#define ROS_WARN_STREAM(x) do { std::ostringstream ros_shim_os; ros_shim_os << x; } while (0)
#define ROS_ERROR_STREAM(x) do { std::ostringstream ros_shim_os; ros_shim_os << x; } while (0)
