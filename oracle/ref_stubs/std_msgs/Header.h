#pragma once
#include <string>
namespace std_msgs { struct Header { unsigned seq = 0; double stamp = 0; std::string frame_id; }; }
