// oracle/shim/pcl -- the PCL names UnevenMap uses (uneven_map.h:91-100, uneven_map.cpp:127-163, 351, 363), enough to COMPILE the reference's
// uneven_map.cpp unmodified and to RUN UnevenMap::constructMap + filter on a cloud the test injects (oracle/ref_map_driver.cpp).
// KdTreeFLANN is a uniform-bin search with the results PCL's radiusSearch / nearestKSearch return (all points with squared float
// distance <= r^2, sorted by distance; the nearest point); PCDReader / CropBox / VoxelGrid / toROSMsg only exist so UnevenMap::init
// compiles -- the pin never calls init (its preprocessing is PCL's; the test feeds the already preprocessed cloud).
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <algorithm>
#include <cmath>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; PointXYZ() {} PointXYZ(float a, float b, float c) : x(a), y(b), z(c) {} };
struct PointXY { float x = 0, y = 0; };
struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };
struct PCLHeader { std::string frame_id; };
template <class P> struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    std::vector<P> points;
    unsigned width = 0, height = 0;
    bool is_dense = true;
    PCLHeader header;
    size_t size() const { return points.size(); }
    void push_back(const P &p) { points.push_back(p); }
    void emplace_back(const P &p) { points.push_back(p); }
    void clear() { points.clear(); }
    Ptr makeShared() const { return Ptr(new PointCloud<P>(*this)); }
};
namespace detail {
inline float px(const PointXYZ &p, int k) { return k == 0 ? p.x : k == 1 ? p.y : p.z; }
inline float px(const PointXY &p, int k) { return k == 0 ? p.x : k == 1 ? p.y : 0.f; }
template <class P> struct Dim { static const int n = 3; };
template <> struct Dim<PointXY> { static const int n = 2; };
} // namespace detail
template <class P> struct KdTreeFLANN {
    typename PointCloud<P>::Ptr cloud;
    double bin = 0.25;
    std::unordered_map<long long, std::vector<int>> bins;
    static long long key(long long a, long long b, long long c) { return ((a + 100000) * 400000 + (b + 100000)) * 400000 + (c + 100000); }
    long long cell(float v) const { return (long long)std::floor((double)v / bin); }
    void setInputCloud(const typename PointCloud<P>::Ptr &c)
    {
        cloud = c;
        bins.clear();
        for (size_t i = 0; i < c->points.size(); i++) {
            const P &p = c->points[i];
            bins[key(cell(detail::px(p, 0)), cell(detail::px(p, 1)), detail::Dim<P>::n == 3 ? cell(detail::px(p, 2)) : 0)].push_back((int)i);
        }
    }
    static float d2(const P &a, const P &b)
    {
        float s = 0.f;
        for (int k = 0; k < detail::Dim<P>::n; k++) { const float d = detail::px(a, k) - detail::px(b, k); s += d * d; }
        return s;
    }
    int radiusSearch(const P &q, double radius, std::vector<int> &idx, std::vector<float> &sq) const
    {
        idx.clear(); sq.clear();
        if (!cloud) return 0;
        const int reach = (int)std::ceil(radius / bin);
        const float r2 = (float)(radius * radius);
        std::vector<std::pair<float, int>> hit;
        const long long c0 = cell(detail::px(q, 0)), c1 = cell(detail::px(q, 1)), c2 = detail::Dim<P>::n == 3 ? cell(detail::px(q, 2)) : 0;
        for (long long a = c0 - reach; a <= c0 + reach; a++)
            for (long long b = c1 - reach; b <= c1 + reach; b++)
                for (long long c = (detail::Dim<P>::n == 3 ? c2 - reach : 0); c <= (detail::Dim<P>::n == 3 ? c2 + reach : 0); c++) {
                    auto it = bins.find(key(a, b, c));
                    if (it == bins.end()) continue;
                    for (int i : it->second) { const float d = d2(cloud->points[i], q); if (d <= r2) hit.push_back(std::make_pair(d, i)); }
                }
        std::sort(hit.begin(), hit.end());
        for (auto &h : hit) { idx.push_back(h.second); sq.push_back(h.first); }
        return (int)idx.size();
    }
    int nearestKSearch(const P &q, int k, std::vector<int> &idx, std::vector<float> &sq) const
    {
        idx.clear(); sq.clear();
        if (!cloud || cloud->points.empty() || k != 1) return 0;
        for (double radius = bin; radius < 1e4; radius *= 2.0) {     // grow the search ball until it holds a point
            std::vector<int> i2; std::vector<float> s2;
            if (radiusSearch(q, radius, i2, s2) > 0) { idx.push_back(i2[0]); sq.push_back(s2[0]); return 1; }
        }
        return 0;
    }
};
struct PCDReader { template <class P> int read(const std::string &, PointCloud<P> &) { return -1; } };
template <class P> struct CropBox {
    template <class V> void setMin(const V &) {}
    template <class V> void setMax(const V &) {}
    void setInputCloud(const typename PointCloud<P>::Ptr &c) { in = c; }
    void filter(PointCloud<P> &out) { if (in) out = *in; }
    typename PointCloud<P>::Ptr in;
};
template <class P> struct VoxelGrid {
    void setLeafSize(float, float, float) {}
    void setInputCloud(const typename PointCloud<P>::Ptr &c) { in = c; }
    void filter(PointCloud<P> &out) { if (in) out = *in; }
    typename PointCloud<P>::Ptr in;
};
template <class C, class M> void toROSMsg(const C &, M &) {}
}
