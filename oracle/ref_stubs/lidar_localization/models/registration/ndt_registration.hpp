#pragma once
// Stand-in (TEST INFRASTRUCTURE ONLY) used when the reference's matching.cpp is compiled for the height-map pin
// (oracle/ref_src/matching_ref.cpp): Matching's constructor makes one of these through the factory; the pinned
// functions (generateGauss2DMapCells / getInitialYawAngle) never call it.  The real class needs PCL's NDT.
#include <yaml-cpp/yaml.h>
#include "lidar_localization/models/registration/registration_interface.hpp"
namespace lidar_localization {
class NDTRegistration : public RegistrationInterface {
  public:
    NDTRegistration(const YAML::Node &) {}
    NDTRegistration(float, float, float, int) {}
    bool SetInputTarget(const CloudData::CLOUD_PTR &) override { return true; }
    bool ScanMatch(const CloudData::CLOUD_PTR &, const Eigen::Matrix4f &predict, CloudData::CLOUD_PTR &, Eigen::Matrix4f &pose) override { pose = predict; return true; }
    float GetFitnessScore() override { return 0.f; }
};
}
