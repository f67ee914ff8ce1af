#pragma once
// stub of <pcl/point_types.h>: PointXYZI memory layout of PCL (32 bytes, 16-aligned)
#include <cstdint>
#include <vector>
#include <eigen3/Eigen/Dense>
#define PCL_ADD_POINT4D union { float data[4]; struct { float x; float y; float z; }; };
#define PCL_ADD_INTENSITY union { float data_c[4]; struct { float intensity; }; };
#define POINT_CLOUD_REGISTER_POINT_STRUCT(name, fseq)
namespace pcl {
struct EIGEN_ALIGN16 PointXYZI {
    PCL_ADD_POINT4D
    PCL_ADD_INTENSITY
    PointXYZI() { x = y = z = 0.f; data[3] = 1.f; intensity = 0.f; data_c[1] = data_c[2] = data_c[3] = 0.f; }
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};
struct PointXYZ { float x, y, z, pad; };
struct PointIndices { std::vector<int> indices; };
}
