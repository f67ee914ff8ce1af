// cloud_filter_interface.hpp -- abstract cloud filter, same surface as the reference's
// lidar_localization/include/lidar_localization/models/cloud_filter/cloud_filter_interface.hpp:13-18.
#ifndef LIDAR_LOCALIZATION_MODELS_CLOUD_FILTER_CLOUD_FILTER_INTERFACE_HPP_
#define LIDAR_LOCALIZATION_MODELS_CLOUD_FILTER_CLOUD_FILTER_INTERFACE_HPP_

#ifdef B2_WITH_YAML
#include <yaml-cpp/yaml.h>
#endif
#include "lidar_localization/sensor_data/cloud_data.hpp"

namespace lidar_localization {
class CloudFilterInterface {
  public:
    virtual ~CloudFilterInterface() = default;

    virtual bool Filter(const CloudData::CLOUD_PTR& input_cloud_ptr, CloudData::CLOUD_PTR& filtered_cloud_ptr) = 0;
};
}  // namespace lidar_localization
#endif
