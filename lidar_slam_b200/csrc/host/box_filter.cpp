// box_filter.cpp -- BoxFilter over the b2cloud C ABI (see box_filter.hpp).
#include "lidar_localization/models/cloud_filter/box_filter.hpp"

#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>

namespace lidar_localization {
namespace {
int DefaultDevice() {
    const char* e = std::getenv("B2NDT_DEVICE");
    return e ? std::atoi(e) : 0;
}
constexpr std::size_t kStride = sizeof(CloudData::POINT);
constexpr std::size_t kIntensityOffset = 16;
}  // namespace

BoxFilter::BoxFilter() : origin_(3, 0.f), size_(6, 0.f), edge_(6, 0.f) {
    if (b2cloud_create(DefaultDevice(), &in_) != B2_OK || b2cloud_create(DefaultDevice(), &out_) != B2_OK) {
        // no CPU fallback: construction fails hard instead of leaving a filter that returns untouched clouds
        const std::string why = std::string("[BoxFilter] ") + b2_last_error();
        std::cerr << why << std::endl;
        b2cloud_destroy(in_); b2cloud_destroy(out_);
        in_ = out_ = nullptr;
        throw std::runtime_error(why);
    }
}

BoxFilter::BoxFilter(const std::vector<float>& size) : BoxFilter() { SetSize(size); }

#ifdef B2_WITH_YAML
BoxFilter::BoxFilter(YAML::Node node) : BoxFilter() {
    std::vector<float> size(6);
    for (size_t i = 0; i < size.size(); i++) size.at(i) = node["box_filter_size"][i].as<float>();
    SetSize(size);
}
#endif

BoxFilter::~BoxFilter() {
    b2cloud_destroy(in_);
    b2cloud_destroy(out_);
}

bool BoxFilter::Filter(const CloudData::CLOUD_PTR& input_cloud_ptr, CloudData::CLOUD_PTR& output_cloud_ptr) {
    const std::size_t n = input_cloud_ptr->points.size();
    if (tmp_.size() < n) tmp_.resize(n);           // staging that only grows (no per-call construction of n points)
    std::size_t m = 0;
    if (!in_ || !out_ || b2cloud_upload(in_, input_cloud_ptr->points.data(), n, kStride, kIntensityOffset) != B2_OK ||
        b2cloud_box_filter(in_, edge_.data(), out_) != B2_OK ||
        b2cloud_download(out_, tmp_.data(), n, kStride, kIntensityOffset, &m) != B2_OK) {
        std::cerr << "[BoxFilter::Filter] " << b2_last_error() << std::endl;
        output_cloud_ptr->points.clear();          // engine failure: defined (empty) output, reported
        output_cloud_ptr->width = 0; output_cloud_ptr->height = 1;
        return false;
    }
    CloudData::CLOUD& out = *output_cloud_ptr;     // output_cloud_ptr->clear() + filter, as the reference
    out.points.assign(tmp_.begin(), tmp_.begin() + m);
    out.width = static_cast<uint32_t>(m);
    out.height = 1;
    out.is_dense = true;
    return true;
}

bool BoxFilter::FilterDevice(b2cloud* input, b2cloud* output) {
    if (b2cloud_box_filter(input, edge_.data(), output) != B2_OK) {
        std::cerr << "[BoxFilter::FilterDevice] " << b2_last_error() << std::endl;
        return false;
    }
    return true;
}

void BoxFilter::SetSize(std::vector<float> size) {
    size_ = size;
    std::cout << "Box Filter size: min_x: " << size.at(0) << ", max_x: " << size.at(1) << ", min_y: " << size.at(2)
              << ", max_y: " << size.at(3) << ", min_z: " << size.at(4) << ", max_z: " << size.at(5) << std::endl;
    CalculateEdge();
}

void BoxFilter::SetOrigin(std::vector<float> origin) {
    origin_ = origin;
    CalculateEdge();
}

void BoxFilter::CalculateEdge() {
    for (size_t i = 0; i < origin_.size(); ++i) {
        edge_.at(2 * i) = size_.at(2 * i) + origin_.at(i);
        edge_.at(2 * i + 1) = size_.at(2 * i + 1) + origin_.at(i);
    }
}

std::vector<float> BoxFilter::GetEdge() { return edge_; }
}  // namespace lidar_localization
