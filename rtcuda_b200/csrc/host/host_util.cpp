// host_util.cpp — thread-local error string behind rtb_last_error()
// (replaces the reference's print-and-exit CHECK_CUDA, utility.cuh:4-13).
#include "host_util.h"

#include "rtb.h"

namespace {
thread_local std::string g_error;
}

namespace rtb {
int set_error(int code, const std::string &msg) {
    g_error = msg;
    return code;
}
}  // namespace rtb

extern "C" {
const char *rtb_last_error(void) { return g_error.c_str(); }
const char *rtb_version(void) { return "rtcuda_b200 0.1 (sm_100a)"; }
}
