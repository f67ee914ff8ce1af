// test_dropin_reference_headers.cpp -- source-level proof of the drop-in: this file and csrc/host/*.cpp are compiled
// with -DB2_WITH_PCL -DB2_WITH_YAML against the REFERENCE'S OWN registration_interface.hpp,
// cloud_filter_interface.hpp and cloud_data.hpp (found first on the include path), with header stand-ins for PCL /
// boost (oracle/ref_stubs) and yaml-cpp (tests/cpp/stubs) and the reference's vendored Eigen.  The objects are made
// exactly as the reference's factories make them (front_end.cpp:52-53,76-77; matching.cpp:59-61,78-80,96-97):
// std::make_shared<NDTRegistration>(YAML node), std::make_shared<VoxelFilter>(YAML node),
// std::make_shared<BoxFilter>(YAML node), held through the interface pointers.
// Exit code: 0 = constructed and (on a GPU box) ran a ScanMatch; 3 = the engine refused to start without a CUDA
// device (what a CPU-only box must see: there is no CPU fallback behind these classes).
#include <cstdio>
#include <memory>
#include <stdexcept>

#include "lidar_localization/models/cloud_filter/box_filter.hpp"
#include "lidar_localization/models/cloud_filter/voxel_filter.hpp"
#include "lidar_localization/models/registration/ndt_registration.hpp"

using namespace lidar_localization;

int main() {
    // front_end.yaml:18-22 / 30-36, matching.yaml box_filter_size, as YAML::Node trees
    YAML::Node ndt;
    ndt.set("res", YAML::Node(1.0)).set("step_size", YAML::Node(0.1)).set("trans_eps", YAML::Node(0.01)).set("max_iter", YAML::Node(30));
    YAML::Node config;
    config.set("registration_method", YAML::Node("NDT")).set("NDT", ndt);
    YAML::Node leaf;
    leaf.push(YAML::Node(1.3)).push(YAML::Node(1.3)).push(YAML::Node(1.3));
    YAML::Node frame;
    frame.set("leaf_size", leaf);
    YAML::Node vfn;
    vfn.set("frame", frame);
    config.set("voxel_filter", vfn);
    YAML::Node box;
    for (double v : {-150.0, 150.0, -150.0, 150.0, -150.0, 150.0}) box.push(YAML::Node(v));
    config.set("box_filter_size", box);

    static_assert(sizeof(CloudData::POINT) == 32, "pcl::PointXYZI layout");
    std::shared_ptr<RegistrationInterface> registration_ptr;
    std::shared_ptr<CloudFilterInterface> filter_ptr;
    std::shared_ptr<BoxFilter> box_filter_ptr;
    try {
        const std::string registration_method = config["registration_method"].as<std::string>();
        registration_ptr = std::make_shared<NDTRegistration>(config[registration_method]);       // front_end.cpp:52-53
        filter_ptr = std::make_shared<VoxelFilter>(config["voxel_filter"]["frame"]);              // front_end.cpp:76-77
        box_filter_ptr = std::make_shared<BoxFilter>(config);                                     // matching.cpp:96-97
    } catch (const std::runtime_error& e) {
        std::printf("ENGINE_REFUSED %s\n", e.what());
        return 3;
    }
    // a GPU is present: one tiny call through every virtual
    CloudData::CLOUD_PTR cloud(new CloudData::CLOUD());
    for (int i = 0; i < 4000; ++i) {
        CloudData::POINT p;
        p.x = 0.05f * (i % 200); p.y = 0.07f * (i / 200); p.z = 0.01f * (i % 7); p.intensity = 1.f;
        cloud->push_back(p);
    }
    CloudData::CLOUD_PTR filtered(new CloudData::CLOUD());
    if (!filter_ptr->Filter(cloud, filtered) || filtered->points.empty()) return 1;
    if (!registration_ptr->SetInputTarget(cloud)) return 1;
    CloudData::CLOUD_PTR result(new CloudData::CLOUD());
    Eigen::Matrix4f pose = Eigen::Matrix4f::Identity(), out = Eigen::Matrix4f::Identity();
    if (!registration_ptr->ScanMatch(filtered, pose, result, out)) return 1;
    if (result->points.size() != filtered->points.size()) return 1;
    std::printf("OK %zu -> %zu points, fitness %g\n", cloud->points.size(), filtered->points.size(), (double)registration_ptr->GetFitnessScore());
    return 0;
}
