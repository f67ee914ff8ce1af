#pragma once
#include <pcl/common/common.h>
