#pragma once
#include <pcl/point_cloud.h>
#include <eigen3/Eigen/Dense>
namespace pcl {
// PCL 1.7 transformPointCloud (dense branch), float arithmetic evaluated left to right
template <typename PointT>
void transformPointCloud(const PointCloud<PointT> &in, PointCloud<PointT> &out, const Eigen::Matrix4f &t) {
    if (&in != &out) { out.points = in.points; out.width = in.width; out.height = in.height; out.is_dense = in.is_dense; }
    for (size_t i = 0; i < out.points.size(); ++i) {
        const float x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
        out.points[i].x = static_cast<float>(t(0, 0) * x + t(0, 1) * y + t(0, 2) * z + t(0, 3));
        out.points[i].y = static_cast<float>(t(1, 0) * x + t(1, 1) * y + t(1, 2) * z + t(1, 3));
        out.points[i].z = static_cast<float>(t(2, 0) * x + t(2, 1) * y + t(2, 2) * z + t(2, 3));
    }
}
}
