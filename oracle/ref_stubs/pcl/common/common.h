#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT> void getMinMax3D(const PointCloud<PointT> &c, PointT &mn, PointT &mx) {
    for (size_t i = 0; i < c.points.size(); ++i) {
        const PointT &p = c.points[i];
        if (i == 0) { mn = p; mx = p; continue; }
        if (p.x < mn.x) mn.x = p.x; if (p.y < mn.y) mn.y = p.y; if (p.z < mn.z) mn.z = p.z;
        if (p.x > mx.x) mx.x = p.x; if (p.y > mx.y) mx.y = p.y; if (p.z > mx.z) mx.z = p.z;
    }
}
template <typename A, typename B> void copyPointCloud(const PointCloud<A> &in, PointCloud<B> &out) { out.points.assign(in.points.begin(), in.points.end()); out.width = in.width; out.height = in.height; }
template <typename A> void copyPointCloud(const PointCloud<A> &in, const std::vector<int> &idx, PointCloud<A> &out) { out.points.clear(); for (size_t i = 0; i < idx.size(); ++i) out.points.push_back(in.points[idx[i]]); }
}
