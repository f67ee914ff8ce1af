// cloud_filter_interface.hpp -- the cloud-filter plug-in boundary of the drop-in.
//
// Restates the CONTRACT of the reference's abstract class (lidar_localization/include/lidar_localization/models/
// cloud_filter/cloud_filter_interface.hpp:13-18): one pure virtual, Filter(in, out), that the factories of the
// front end, the matching node, loop closing and the viewer call through a shared_ptr<CloudFilterInterface>
// (front_end.cpp:76-77,106-107; matching.cpp:78-80,158; loop_closing.cpp:97-99,293,305; viewer.cpp:67-69,207).
// Inside the reference's catkin workspace keep the reference's own file.
//
// Contract kept by the B200 filters (voxel_filter.hpp, box_filter.hpp):
//   * `out` may be the SAME pointer as `in` (the reference filters in place in several callers) or an empty cloud;
//   * the result replaces the contents of `out` (width = number of points, height = 1, is_dense = true);
//   * the return value is always true; problems are logged, never thrown.
#ifndef LIDAR_LOCALIZATION_MODELS_CLOUD_FILTER_CLOUD_FILTER_INTERFACE_HPP_
#define LIDAR_LOCALIZATION_MODELS_CLOUD_FILTER_CLOUD_FILTER_INTERFACE_HPP_

#ifdef B2_WITH_YAML
#include <yaml-cpp/yaml.h>      // the YAML::Node constructors of the concrete classes
#endif
#include "lidar_localization/sensor_data/cloud_data.hpp"

namespace lidar_localization {

class CloudFilterInterface {
  public:
    virtual ~CloudFilterInterface() = default;

    // in -> out (aliasing allowed)
    virtual bool Filter(const CloudData::CLOUD_PTR& in, CloudData::CLOUD_PTR& out) = 0;
};

}  // namespace lidar_localization
#endif  // LIDAR_LOCALIZATION_MODELS_CLOUD_FILTER_CLOUD_FILTER_INTERFACE_HPP_
