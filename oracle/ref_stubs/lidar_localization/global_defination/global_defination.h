#pragma once
// stand-in for the cmake-generated global_defination.h (from global_defination.h.in): TEST INFRASTRUCTURE ONLY
#include <string>
namespace lidar_localization { const std::string WORK_SPACE_PATH = "."; }
