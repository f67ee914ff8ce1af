// voxel_filter.cpp -- VoxelFilter over the b2vf C ABI (see voxel_filter.hpp).
#include "lidar_localization/models/cloud_filter/voxel_filter.hpp"

#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace lidar_localization {
namespace {
int DefaultDevice() {
    const char* e = std::getenv("B2NDT_DEVICE");
    return e ? std::atoi(e) : 0;
}
constexpr std::size_t kStride = sizeof(CloudData::POINT);
constexpr std::size_t kIntensityOffset = 16;
}  // namespace

#ifdef B2_WITH_YAML
VoxelFilter::VoxelFilter(const YAML::Node& node) {
    SetFilterParam(node["leaf_size"][0].as<float>(), node["leaf_size"][1].as<float>(), node["leaf_size"][2].as<float>());
}
#endif

VoxelFilter::VoxelFilter(float leaf_size_x, float leaf_size_y, float leaf_size_z) {
    SetFilterParam(leaf_size_x, leaf_size_y, leaf_size_z);
}

VoxelFilter::~VoxelFilter() { b2vf_destroy(vf_); }

bool VoxelFilter::SetFilterParam(float leaf_size_x, float leaf_size_y, float leaf_size_z) {
    if (vf_) { b2vf_destroy(vf_); vf_ = nullptr; }
    if (b2vf_create(leaf_size_x, leaf_size_y, leaf_size_z, DefaultDevice(), &vf_) != B2_OK) {
        // no CPU fallback: a filter without an engine would return untouched clouds, so construction fails hard
        vf_ = nullptr;
        const std::string why = std::string("[VoxelFilter] ") + b2_last_error();
        std::cerr << why << std::endl;
        throw std::runtime_error(why);
    }
    std::cout << "Voxel Filter params: " << leaf_size_x << ", " << leaf_size_y << ", " << leaf_size_z << std::endl;
    return true;
}

bool VoxelFilter::Filter(const CloudData::CLOUD_PTR& input_cloud_ptr, CloudData::CLOUD_PTR& filtered_cloud_ptr) {
    const std::size_t n = input_cloud_ptr->points.size();
    // pcl::Filter::filter computes into a temporary when output aliases the input; same here.  The temporary is a
    // member that only grows: constructing n points (3.7 MB for a raw HDL-64 frame) per call costs more than the filter
    if (tmp_.size() < n) tmp_.resize(n);
    std::size_t m = 0;
    if (!vf_ || b2vf_filter(vf_, input_cloud_ptr->points.data(), n, kStride, kIntensityOffset, tmp_.data(), n, kStride,
                            kIntensityOffset, &m, nullptr, nullptr) != B2_OK) {
        // engine failure: PCL cannot fail here; define the output (a copy of the input, what pcl::VoxelGrid itself
        // emits when it gives up) and report false
        std::cerr << "[VoxelFilter::Filter] " << b2_last_error() << std::endl;
        if (filtered_cloud_ptr.get() != input_cloud_ptr.get()) *filtered_cloud_ptr = *input_cloud_ptr;
        return false;
    }
    CloudData::CLOUD& out = *filtered_cloud_ptr;
    out.points.assign(tmp_.begin(), tmp_.begin() + m);
    out.width = static_cast<uint32_t>(m);
    out.height = 1;
    out.is_dense = true;
    return true;
}
bool VoxelFilter::FilterDevice(b2cloud* input, b2cloud* output) {
    if (!vf_ || b2vf_filter_cloud(vf_, input, output) != B2_OK) {
        std::cerr << "[VoxelFilter::FilterDevice] " << b2_last_error() << std::endl;
        return false;
    }
    return true;
}
}  // namespace lidar_localization
