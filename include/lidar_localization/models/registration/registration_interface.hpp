// registration_interface.hpp -- abstract scan matcher, same surface as the reference's
// lidar_localization/include/lidar_localization/models/registration/registration_interface.hpp:14-24.
#ifndef LIDAR_LOCALIZATION_MODELS_REGISTRATION_INTERFACE_HPP_
#define LIDAR_LOCALIZATION_MODELS_REGISTRATION_INTERFACE_HPP_

#ifdef B2_WITH_YAML
#include <yaml-cpp/yaml.h>
#endif
#include "lidar_localization/sensor_data/cloud_data.hpp"

namespace lidar_localization {
class RegistrationInterface {
  public:
    virtual ~RegistrationInterface() = default;

    virtual bool SetInputTarget(const CloudData::CLOUD_PTR& input_target) = 0;
    virtual bool ScanMatch(const CloudData::CLOUD_PTR& input_source,
                           const Eigen::Matrix4f& predict_pose,
                           CloudData::CLOUD_PTR& result_cloud_ptr,
                           Eigen::Matrix4f& result_pose) = 0;
    virtual float GetFitnessScore() = 0;
};
}  // namespace lidar_localization
#endif
