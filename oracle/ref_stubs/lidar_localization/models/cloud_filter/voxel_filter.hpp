#pragma once
// Stand-in (TEST INFRASTRUCTURE ONLY), see ../registration/ndt_registration.hpp in this directory tree.
#include <yaml-cpp/yaml.h>
#include "lidar_localization/models/cloud_filter/cloud_filter_interface.hpp"
namespace lidar_localization {
class VoxelFilter : public CloudFilterInterface {
  public:
    VoxelFilter(const YAML::Node &) {}
    VoxelFilter(float, float, float) {}
    bool Filter(const CloudData::CLOUD_PTR &in, CloudData::CLOUD_PTR &out) override { if (in.get() != out.get()) *out = *in; return true; }
};
}
