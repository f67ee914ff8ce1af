// ndt_registration.hpp -- B200 drop-in for the reference's NDTRegistration
// (lidar_localization/include/lidar_localization/models/registration/ndt_registration.hpp:13-30,
//  src/models/registration/ndt_registration.cpp:12-66).  Same constructors, same three virtuals, same
// "always returns true" error behaviour; the pcl::NormalDistributionsTransform member is replaced by a
// b2ndt handle (include/b2ndt.h).  The factories that select it by the YAML string "NDT" are
// front_end.cpp:52-53, matching.cpp:59-61, loop_closing.cpp:78-80.
#ifndef LIDAR_LOCALIZATION_MODELS_REGISTRATION_NDT_REGISTRATION_HPP_
#define LIDAR_LOCALIZATION_MODELS_REGISTRATION_NDT_REGISTRATION_HPP_

#include <vector>

#include "b2ndt.h"
#include "lidar_localization/models/registration/registration_interface.hpp"

namespace lidar_localization {
class NDTRegistration : public RegistrationInterface {
  public:
#ifdef B2_WITH_YAML
    NDTRegistration(const YAML::Node& node);
#endif
    NDTRegistration(float res, float step_size, float trans_eps, int max_iter);
    ~NDTRegistration() override;
    NDTRegistration(const NDTRegistration&) = delete;
    NDTRegistration& operator=(const NDTRegistration&) = delete;

    bool SetInputTarget(const CloudData::CLOUD_PTR& input_target) override;
    bool ScanMatch(const CloudData::CLOUD_PTR& input_source,
                   const Eigen::Matrix4f& predict_pose,
                   CloudData::CLOUD_PTR& result_cloud_ptr,
                   Eigen::Matrix4f& result_pose) override;
    float GetFitnessScore() override;

    // Extension (not in the reference): many independent ScanMatch calls against the current target
    // in one launch (BASELINE.json configs 4/5).  result_poses is resized to sources.size().
    bool ScanMatchBatch(const std::vector<CloudData::CLOUD_PTR>& sources,
                        const std::vector<Eigen::Matrix4f>& predict_poses,
                        std::vector<Eigen::Matrix4f>& result_poses,
                        std::vector<b2ndt_result>* details = nullptr);
    // Extension: clouds that already live in HBM (b2cloud, include/b2ndt.h): the front end's local map assembled on the
    // device, the matching node's cropped map, a frame filtered on the device -- no host copies.  result_cloud may be null.
    bool SetInputTargetDevice(b2cloud* input_target);
    // Extension: the in-tree NDT's NormalDistributionsTransform::updateVoxelGrid(new_cloud)
    // (ndt_registration_manual/NormalDistributionsTransform.cpp:968-972): add a cloud to the current target without a
    // full rebuild; same target as SetInputTarget(old + new), bit for bit (include/b2ndt.h: b2ndt_update_target).
    bool UpdateInputTarget(const CloudData::CLOUD_PTR& new_cloud);
    bool UpdateInputTargetDevice(b2cloud* new_cloud);
    bool ScanMatchDevice(b2cloud* input_source, const Eigen::Matrix4f& predict_pose, b2cloud* result_cloud,
                         Eigen::Matrix4f& result_pose);
    // details of the last ScanMatch (iterations, converged, score ...)
    const b2ndt_result& LastResult() const { return last_; }
    // device ordinal used by objects constructed afterwards (default 0 or $B2NDT_DEVICE)
    static void SetDefaultDevice(int device);

  private:
    bool SetRegistrationParam(float res, float step_size, float trans_eps, int max_iter);

  private:
    b2ndt* ndt_ = nullptr;
    b2ndt_result last_{};
};
}  // namespace lidar_localization
#endif
