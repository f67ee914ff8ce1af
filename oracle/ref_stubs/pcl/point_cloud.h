#pragma once
#include <boost/shared_ptr.hpp>
#include <vector>
#include <eigen3/Eigen/StdVector>
namespace pcl {
struct PCLHeader { unsigned seq = 0; unsigned long stamp = 0; };
template <typename PointT>
class PointCloud {
  public:
    typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
    typedef boost::shared_ptr<const PointCloud<PointT> > ConstPtr;
    std::vector<PointT, Eigen::aligned_allocator<PointT> > points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    PCLHeader header;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = height = 0; }
    void push_back(const PointT &p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
    PointCloud &operator+=(const PointCloud &o) { points.insert(points.end(), o.points.begin(), o.points.end()); width = (uint32_t)points.size(); height = 1; return *this; }
    PointT &operator[](size_t i) { return points[i]; }
    const PointT &operator[](size_t i) const { return points[i]; }
};
}
