#pragma once
#include <std_msgs/Header.h>
namespace jsk_recognition_msgs {
struct V3 { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { V3 position; V3 orientation; };
struct BoundingBox { std_msgs::Header header; Pose pose; V3 dimensions; float value = 0; unsigned label = 0; };
}
