// host_util.h — error reporting shared by the host-side translation units.
#pragma once
#include <string>

namespace rtb {
// records the message for rtb_last_error() and returns `code`
int set_error(int code, const std::string &msg);
}  // namespace rtb
