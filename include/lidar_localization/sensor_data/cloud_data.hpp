// cloud_data.hpp -- point / cloud types crossing the plug-in boundary.
//
// Mirrors lidar_localization/include/lidar_localization/sensor_data/cloud_data.hpp:33-48 of the reference
// (CloudData::POINT = pcl::PointXYZI, CLOUD = pcl::PointCloud<POINT>, CLOUD_PTR = CLOUD::Ptr).
// With -DB2_WITH_PCL the real PCL / Eigen headers are used (the drop-in build inside the reference's
// catkin workspace).  Without it (this repository: no PCL, no Eigen, no boost in the image) a minimal
// layout-compatible shim is provided: PointXYZI is the same 32-byte, 16-aligned record
// {x,y,z,1 | intensity,pad[3]}, PointCloud has points/width/height/is_dense, Ptr is std::shared_ptr.
#ifndef LIDAR_LOCALIZATION_SENSOR_DATA_CLOUD_DATA_HPP_
#define LIDAR_LOCALIZATION_SENSOR_DATA_CLOUD_DATA_HPP_

#ifdef B2_WITH_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <Eigen/Dense>
#else
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

namespace pcl {
struct alignas(16) PointXYZI {
    union { float data[4]; struct { float x, y, z; }; };
    union { float data_c[4]; struct { float intensity; }; };
    PointXYZI() : data{0.f, 0.f, 0.f, 1.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};
static_assert(sizeof(PointXYZI) == 32, "PointXYZI must be 32 bytes like pcl::PointXYZI");

template <typename PointT>
class PointCloud {
  public:
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = height = 0; }
    void push_back(const PointT &p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
    PointT &operator[](std::size_t i) { return points[i]; }
    const PointT &operator[](std::size_t i) const { return points[i]; }
};
}  // namespace pcl

namespace Eigen {
// column-major 4x4 float, the memory layout of Eigen::Matrix4f
struct Matrix4f {
    float m[16];
    Matrix4f() : m{0} {}
    static Matrix4f Identity() { Matrix4f r; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.f; return r; }
    float &operator()(int r, int c) { return m[c * 4 + r]; }
    float operator()(int r, int c) const { return m[c * 4 + r]; }
    float *data() { return m; }
    const float *data() const { return m; }
};
}  // namespace Eigen
#endif  // B2_WITH_PCL

namespace lidar_localization {
class CloudData {
  public:
    using POINT = pcl::PointXYZI;
    using CLOUD = pcl::PointCloud<POINT>;
    using CLOUD_PTR = CLOUD::Ptr;

  public:
    CloudData() : cloud_ptr(new CLOUD()) {}

  public:
    double time = 0.0;
    CLOUD_PTR cloud_ptr;
};
}  // namespace lidar_localization
#endif
