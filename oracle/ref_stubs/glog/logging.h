#pragma once
// stub of <glog/logging.h>: the reference's translation units only stream into LOG(...)
#include <iostream>
struct B2NullLog { template <typename T> B2NullLog &operator<<(const T &) { return *this; } B2NullLog &operator<<(std::ostream &(*)(std::ostream &)) { return *this; } };
#define LOG(x) B2NullLog()
