#pragma once
// stub of <pcl/io/pcd_io.h> (TEST INFRASTRUCTURE ONLY): the matching harness never loads a file; also carries the two
// PCL helpers matching.cpp reaches through PCL's transitive includes (pcl::isFinite, pcl::removeNaNFromPointCloud)
#include <cmath>
#include <string>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl {
namespace io { template <typename CloudT> int loadPCDFile(const std::string &, CloudT &) { return -1; } }
template <typename PointT> inline bool isFinite(const PointT &p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }
template <typename PointT>
void removeNaNFromPointCloud(const PointCloud<PointT> &in, PointCloud<PointT> &out, std::vector<int> &index) {
    std::vector<PointT, Eigen::aligned_allocator<PointT> > keep;
    index.clear();
    for (size_t i = 0; i < in.points.size(); ++i)
        if (isFinite(in.points[i])) { keep.push_back(in.points[i]); index.push_back((int)i); }
    out.points.assign(keep.begin(), keep.end());
    out.width = (unsigned)out.points.size(); out.height = 1; out.is_dense = true;
}
}
