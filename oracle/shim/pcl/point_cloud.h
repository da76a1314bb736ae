// oracle/shim/pcl -- just enough of the PCL type names for UnevenMap's member declarations (uneven_map.h:91-100) to compile; the pin
// build never constructs a map from a cloud (that is uneven_map.cpp, not compiled here).
#pragma once
#include <memory>
#include <vector>
#include <string>
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; PointXYZ() {} PointXYZ(float a, float b, float c) : x(a), y(b), z(c) {} };
struct PointXY { float x = 0, y = 0; };
template <class P> struct PointCloud { typedef std::shared_ptr<PointCloud<P>> Ptr; std::vector<P> points; size_t size() const { return points.size(); } void push_back(const P &p) { points.push_back(p); } };
template <class P> struct KdTreeFLANN {
    void setInputCloud(const typename PointCloud<P>::Ptr &) {}
    int nearestKSearch(const P &, int, std::vector<int> &, std::vector<float> &) const { return 0; }
    int radiusSearch(const P &, double, std::vector<int> &, std::vector<float> &) const { return 0; }
};
}
