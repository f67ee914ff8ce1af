// voxel_filter.hpp -- B200 drop-in for the reference's VoxelFilter
// (lidar_localization/include/lidar_localization/models/cloud_filter/voxel_filter.hpp:13-25,
//  src/models/cloud_filter/voxel_filter.cpp:12-41).  The pcl::VoxelGrid member is replaced by a b2vf
// handle.  Filter(in, out) may be called with in == out (matching.cpp:158, loop_closing.cpp:293,305,
// viewer.cpp:207) and with an empty pre-allocated out (front_end.cpp:106-107).
#ifndef LIDAR_LOCALIZATION_MODELS_CLOUD_FILTER_VOXEL_FILTER_HPP_
#define LIDAR_LOCALIZATION_MODELS_CLOUD_FILTER_VOXEL_FILTER_HPP_

#include <vector>

#include "b2ndt.h"
#include "lidar_localization/models/cloud_filter/cloud_filter_interface.hpp"

namespace lidar_localization {

// What Filter computes (pcl::VoxelGrid semantics, reproduced bit for bit on voxel indices and point counts):
//   * bounding box of the finite points, inverse leaf = 1.0f / leaf, voxel (i,j,k) = floor(p * inverse leaf) - min;
//   * one output point per occupied voxel = float mean of x, y, z and intensity of its members (summed in input
//     order), voxels emitted in ascending linear index i + j*dx + k*dx*dy;
//   * "leaf size too small" (more than 2^31 cells): the input is returned unchanged, as PCL does.
// The work runs on the GPU behind b2vf_filter (radix sort by voxel key + per-voxel reduction); FilterDevice-style
// use with resident clouds goes through b2vf_filter_cloud (include/b2ndt.h).
class VoxelFilter : public CloudFilterInterface {
  public:
#ifdef B2_WITH_YAML
    VoxelFilter(const YAML::Node& node);
#endif
    VoxelFilter(float leaf_size_x, float leaf_size_y, float leaf_size_z);
    ~VoxelFilter() override;
    VoxelFilter(const VoxelFilter&) = delete;
    VoxelFilter& operator=(const VoxelFilter&) = delete;

    bool Filter(const CloudData::CLOUD_PTR& input_cloud_ptr, CloudData::CLOUD_PTR& filtered_cloud_ptr) override;
    // device-resident variant (input == output allowed)
    bool FilterDevice(b2cloud* input, b2cloud* output);

  private:
    bool SetFilterParam(float leaf_size_x, float leaf_size_y, float leaf_size_z);

  private:
    b2vf* vf_ = nullptr;
    decltype(CloudData::CLOUD().points) tmp_;   // output staging (same vector type / allocator as the cloud), reused across calls
};
}  // namespace lidar_localization
#endif
