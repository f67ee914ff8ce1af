#pragma once
#include <vector>
#include <jsk_recognition_msgs/BoundingBox.h>
namespace jsk_recognition_msgs { struct BoundingBoxArray { std_msgs::Header header; std::vector<BoundingBox> boxes; }; }
