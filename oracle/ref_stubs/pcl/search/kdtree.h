#pragma once
#include <pcl/point_cloud.h>
namespace pcl { namespace search {
template <typename PointT> class KdTree {
  public:
    typedef boost::shared_ptr<KdTree<PointT> > Ptr;
    void setInputCloud(const typename PointCloud<PointT>::Ptr &) {}
};
}}
