#pragma once
#include <pcl/point_types.h>
#include <pcl/search/kdtree.h>
namespace pcl {
// no-op: cluster_by_distance is never called on the registration path
template <typename PointT> class EuclideanClusterExtraction {
  public:
    void setInputCloud(const typename PointCloud<PointT>::Ptr &) {}
    void setClusterTolerance(double) {}
    void setMinClusterSize(int) {}
    void setMaxClusterSize(int) {}
    void setSearchMethod(const typename search::KdTree<PointT>::Ptr &) {}
    void extract(std::vector<PointIndices> &out) { out.clear(); }
};
}
