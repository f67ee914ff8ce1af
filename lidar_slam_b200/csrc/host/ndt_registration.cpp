// ndt_registration.cpp -- NDTRegistration over the b2ndt C ABI (see ndt_registration.hpp).
#include "lidar_localization/models/registration/ndt_registration.hpp"

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>

namespace lidar_localization {
namespace {
int g_default_device = -1;
int DefaultDevice() {
    if (g_default_device >= 0) return g_default_device;
    const char* e = std::getenv("B2NDT_DEVICE");
    return e ? std::atoi(e) : 0;
}
constexpr std::size_t kStride = sizeof(CloudData::POINT);          // 32
constexpr std::size_t kIntensityOffset = 16;                       // data_c[0]
}  // namespace

void NDTRegistration::SetDefaultDevice(int device) { g_default_device = device; }

#ifdef B2_WITH_YAML
NDTRegistration::NDTRegistration(const YAML::Node& node) {
    SetRegistrationParam(node["res"].as<float>(), node["step_size"].as<float>(), node["trans_eps"].as<float>(),
                         node["max_iter"].as<int>());
}
#endif

NDTRegistration::NDTRegistration(float res, float step_size, float trans_eps, int max_iter) {
    SetRegistrationParam(res, step_size, trans_eps, max_iter);
}

NDTRegistration::~NDTRegistration() { b2ndt_destroy(ndt_); }

bool NDTRegistration::SetRegistrationParam(float res, float step_size, float trans_eps, int max_iter) {
    b2ndt_params p;
    b2ndt_params_default(&p);
    p.res = res;
    p.step_size = step_size;   // float -> double exactly as pcl::NDT::setStepSize(double) receives it
    p.trans_eps = trans_eps;
    p.max_iter = max_iter;
    if (ndt_) { b2ndt_destroy(ndt_); ndt_ = nullptr; }
    if (b2ndt_create(&p, DefaultDevice(), &ndt_) != B2_OK) {
        // there is no CPU fallback behind this class: an object without an engine would hand stale poses to the
        // front end, so construction fails hard (the reference's PCL constructor cannot fail this way)
        ndt_ = nullptr;
        const std::string why = std::string("[NDTRegistration] ") + b2_last_error();
        std::cerr << why << std::endl;
        throw std::runtime_error(why);
    }
    std::cout << "NDT params: res: " << res << ", step_size: " << step_size << ", trans_eps: " << trans_eps
              << ", max_iter: " << max_iter << std::endl;
    return true;
}

bool NDTRegistration::SetInputTarget(const CloudData::CLOUD_PTR& input_target) {
    if (!ndt_ || b2ndt_set_target(ndt_, input_target->points.data(), input_target->points.size(), kStride, kIntensityOffset) != B2_OK) {
        std::cerr << "[NDTRegistration::SetInputTarget] " << b2_last_error() << std::endl;
        return false;   // engine failure (device lost, out of memory): PCL cannot fail here, so this is reported
    }
    return true;
}

bool NDTRegistration::ScanMatch(const CloudData::CLOUD_PTR& input_source, const Eigen::Matrix4f& predict_pose,
                                CloudData::CLOUD_PTR& result_cloud_ptr, Eigen::Matrix4f& result_pose) {
    const std::size_t n = input_source->points.size();
    float pose[16];
    std::memcpy(pose, predict_pose.data(), sizeof(pose));
    // a large (unfiltered) source: let the library fill the result cloud from the device copy of the source
    const bool device_fill = result_cloud_ptr && n >= 16384;
    if (device_fill && result_cloud_ptr.get() != input_source.get()) {
        result_cloud_ptr->points.resize(n);
        result_cloud_ptr->width = input_source->width; result_cloud_ptr->height = input_source->height; result_cloud_ptr->is_dense = input_source->is_dense;
    }
    const bool ok = ndt_ && (device_fill
        ? b2ndt_align_ex(ndt_, input_source->points.data(), n, kStride, kIntensityOffset, predict_pose.data(), pose, &last_,
                         result_cloud_ptr->points.data(), kStride, kIntensityOffset)
        : b2ndt_align(ndt_, input_source->points.data(), n, kStride, kIntensityOffset, predict_pose.data(), pose, &last_)) == B2_OK;
    if (ok && device_fill) {
        std::memcpy(result_pose.data(), pose, sizeof(pose));
        return true;
    }
    if (!ok) {
        // engine failure: the outputs are still defined (pose = prediction, cloud = source moved by it) and the call
        // reports false -- the reference's `return true` (ndt_registration.cpp:60) holds for PCL, which cannot fail
        std::cerr << "[NDTRegistration::ScanMatch] " << b2_last_error() << std::endl;
        std::memcpy(pose, predict_pose.data(), sizeof(pose));
        std::memset(&last_, 0, sizeof(last_));
    }
    std::memcpy(result_pose.data(), pose, sizeof(pose));
    // align(output): the source transformed by the final pose, float arithmetic of pcl::transformPointCloud
    if (result_cloud_ptr) {
        CloudData::CLOUD& out = *result_cloud_ptr;
        const bool alias = (result_cloud_ptr.get() == input_source.get());
        if (!alias) {
            out.points.resize(n);
            out.width = input_source->width; out.height = input_source->height; out.is_dense = input_source->is_dense;
        }
        for (std::size_t i = 0; i < n; ++i) {
            const CloudData::POINT& s = input_source->points[i];
            const float x = s.x, y = s.y, z = s.z;
            CloudData::POINT o = s;
            o.x = ((pose[0] * x + pose[4] * y) + pose[8] * z) + pose[12];
            o.y = ((pose[1] * x + pose[5] * y) + pose[9] * z) + pose[13];
            o.z = ((pose[2] * x + pose[6] * y) + pose[10] * z) + pose[14];
            o.data[3] = 1.0f;
            out.points[i] = o;
        }
    }
    return ok;
}

bool NDTRegistration::SetInputTargetDevice(b2cloud* input_target) {
    if (!ndt_ || b2ndt_set_target_cloud(ndt_, input_target) != B2_OK) {
        std::cerr << "[NDTRegistration::SetInputTargetDevice] " << b2_last_error() << std::endl;
        return false;
    }
    return true;
}

bool NDTRegistration::UpdateInputTarget(const CloudData::CLOUD_PTR& new_cloud) {
    if (!ndt_ || b2ndt_update_target(ndt_, new_cloud->points.data(), new_cloud->points.size(), kStride, kIntensityOffset) != B2_OK) {
        std::cerr << "[NDTRegistration::UpdateInputTarget] " << b2_last_error() << std::endl;
        return false;
    }
    return true;
}

bool NDTRegistration::UpdateInputTargetDevice(b2cloud* new_cloud) {
    if (!ndt_ || b2ndt_update_target_cloud(ndt_, new_cloud) != B2_OK) {
        std::cerr << "[NDTRegistration::UpdateInputTargetDevice] " << b2_last_error() << std::endl;
        return false;
    }
    return true;
}

bool NDTRegistration::ScanMatchDevice(b2cloud* input_source, const Eigen::Matrix4f& predict_pose, b2cloud* result_cloud,
                                      Eigen::Matrix4f& result_pose) {
    float pose[16];
    if (!ndt_ || b2ndt_align_cloud(ndt_, input_source, predict_pose.data(), pose, &last_, result_cloud) != B2_OK) {
        std::cerr << "[NDTRegistration::ScanMatchDevice] " << b2_last_error() << std::endl;
        std::memcpy(result_pose.data(), predict_pose.data(), sizeof(pose));
        std::memset(&last_, 0, sizeof(last_));
        return false;
    }
    std::memcpy(result_pose.data(), pose, sizeof(pose));
    return true;
}

float NDTRegistration::GetFitnessScore() {
    double v = 0.0;
    if (!ndt_ || b2ndt_fitness(ndt_, 1.7976931348623157e308, &v) != B2_OK) {
        std::cerr << "[NDTRegistration::GetFitnessScore] " << b2_last_error() << std::endl;
        return 3.402823466e+38f;
    }
    return static_cast<float>(v);
}

bool NDTRegistration::ScanMatchBatch(const std::vector<CloudData::CLOUD_PTR>& sources,
                                     const std::vector<Eigen::Matrix4f>& predict_poses,
                                     std::vector<Eigen::Matrix4f>& result_poses, std::vector<b2ndt_result>* details) {
    const std::size_t B = sources.size();
    result_poses.assign(B, Eigen::Matrix4f::Identity());
    if (!ndt_ || predict_poses.size() != B || B == 0) return B == 0;
    std::vector<uint32_t> off(B + 1, 0);
    for (std::size_t b = 0; b < B; ++b) off[b + 1] = off[b] + (uint32_t)sources[b]->points.size();
    std::vector<CloudData::POINT> all(off[B]);
    for (std::size_t b = 0; b < B; ++b)
        if (!sources[b]->points.empty())
            std::memcpy(&all[off[b]], sources[b]->points.data(), sources[b]->points.size() * kStride);
    std::vector<float> g(B * 16), out(B * 16);
    for (std::size_t b = 0; b < B; ++b) std::memcpy(&g[b * 16], predict_poses[b].data(), 64);
    if (details) details->resize(B);
    if (b2ndt_align_batch(ndt_, all.data(), all.size(), kStride, kIntensityOffset, off.data(), B, g.data(), out.data(),
                          details ? details->data() : nullptr) != B2_OK) {
        std::cerr << "[NDTRegistration::ScanMatchBatch] " << b2_last_error() << std::endl;
        return false;
    }
    for (std::size_t b = 0; b < B; ++b) std::memcpy(result_poses[b].data(), &out[b * 16], 64);
    return true;
}
}  // namespace lidar_localization
