// box_filter.hpp -- B200 drop-in for the reference's BoxFilter
// (lidar_localization/include/lidar_localization/models/cloud_filter/box_filter.hpp:15-38,
//  src/models/cloud_filter/box_filter.cpp:12-75).  The pcl::CropBox member is replaced by device clouds:
// Filter uploads the input, crops on the GPU (order kept, bounds inclusive as pcl::CropBox) and downloads.
// Callers that keep their map resident use FilterDevice (matching.cpp:166-183 re-crops the same global map).
#ifndef LIDAR_LOCALIZATION_MODELS_CLOUD_FILTER_BOX_FILTER_HPP_
#define LIDAR_LOCALIZATION_MODELS_CLOUD_FILTER_BOX_FILTER_HPP_

#include <vector>

#include "b2ndt.h"
#include "lidar_localization/models/cloud_filter/cloud_filter_interface.hpp"

namespace lidar_localization {
class BoxFilter : public CloudFilterInterface {
  public:
#ifdef B2_WITH_YAML
    BoxFilter(YAML::Node node);
#endif
    BoxFilter();
    explicit BoxFilter(const std::vector<float>& size);
    ~BoxFilter() override;
    BoxFilter(const BoxFilter&) = delete;
    BoxFilter& operator=(const BoxFilter&) = delete;

    bool Filter(const CloudData::CLOUD_PTR& input_cloud_ptr, CloudData::CLOUD_PTR& filtered_cloud_ptr) override;
    // device-resident variant: crop `input` (b2cloud) into `output` (b2cloud), no host copies
    bool FilterDevice(b2cloud* input, b2cloud* output);

    void SetSize(std::vector<float> size);
    void SetOrigin(std::vector<float> origin);
    std::vector<float> GetEdge();

  private:
    void CalculateEdge();

  private:
    b2cloud* in_ = nullptr;
    b2cloud* out_ = nullptr;
    std::vector<float> origin_;
    std::vector<float> size_;
    std::vector<float> edge_;
    decltype(CloudData::CLOUD().points) tmp_;      // output staging, reused across calls
};
}  // namespace lidar_localization
#endif
