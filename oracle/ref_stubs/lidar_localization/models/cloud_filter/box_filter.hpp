#pragma once
// Stand-in (TEST INFRASTRUCTURE ONLY), see ../registration/ndt_registration.hpp in this directory tree.
#include <vector>
#include <yaml-cpp/yaml.h>
#include "lidar_localization/models/cloud_filter/cloud_filter_interface.hpp"
namespace lidar_localization {
class BoxFilter : public CloudFilterInterface {
  public:
    BoxFilter(const YAML::Node &) {}
    bool Filter(const CloudData::CLOUD_PTR &in, CloudData::CLOUD_PTR &out) override { if (in.get() != out.get()) *out = *in; return true; }
    void SetSize(std::vector<float>) {}
    void SetOrigin(std::vector<float>) {}
    std::vector<float> GetEdge() { return std::vector<float>(6, 0.f); }
};
}
