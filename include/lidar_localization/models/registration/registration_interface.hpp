// registration_interface.hpp -- the scan-matcher plug-in boundary of the drop-in.
//
// This header only restates the CONTRACT of the reference's abstract class (lidar_localization/include/
// lidar_localization/models/registration/registration_interface.hpp:14-24): three pure virtuals whose names,
// argument order and types a caller (front_end.cpp:52-53,221; matching.cpp:59-61,245; loop_closing.cpp:78-80,253)
// compiles against.  Inside the reference's catkin workspace keep the reference's own file; this copy exists so
// that the replacement classes build where PCL / yaml-cpp are not installed.
//
// Semantics the B200 implementation (ndt_registration.hpp) keeps:
//   * SetInputTarget  copies the target to the device and builds the NDT voxel grid there; the caller's cloud is
//                     not referenced afterwards.  Returns true (the reference never reports failure here).
//   * ScanMatch       aligns `source` to the target starting from `predict_pose`; fills `result_cloud` with the
//                     source under the final pose and `result_pose` with that pose (Eigen column-major 4x4 float).
//   * GetFitnessScore mean squared distance from the last transformed source to its nearest target points
//                     (pcl::Registration::getFitnessScore); valid after a ScanMatch on the same object.
// One object = one device handle = one CUDA stream; not thread-safe (the reference's nodes are single-threaded).
#ifndef LIDAR_LOCALIZATION_MODELS_REGISTRATION_INTERFACE_HPP_
#define LIDAR_LOCALIZATION_MODELS_REGISTRATION_INTERFACE_HPP_

#ifdef B2_WITH_YAML
#include <yaml-cpp/yaml.h>      // the YAML::Node constructors of the concrete classes
#endif
#include "lidar_localization/sensor_data/cloud_data.hpp"

namespace lidar_localization {

class RegistrationInterface {
  public:
    virtual ~RegistrationInterface() = default;

    // target map / local map the following ScanMatch calls register against
    virtual bool SetInputTarget(const CloudData::CLOUD_PTR& target) = 0;

    // source -> target alignment from an initial guess; outputs: transformed source, final pose
    virtual bool ScanMatch(const CloudData::CLOUD_PTR& source, const Eigen::Matrix4f& predict_pose,
                           CloudData::CLOUD_PTR& result_cloud, Eigen::Matrix4f& result_pose) = 0;

    // quality of the last alignment (lower is better)
    virtual float GetFitnessScore() = 0;
};

}  // namespace lidar_localization
#endif  // LIDAR_LOCALIZATION_MODELS_REGISTRATION_INTERFACE_HPP_
