// host_util.h — error reporting shared by the host-side translation units.
#pragma once
#include <string>

namespace rtb {
// records the message for rtb_last_error() and returns `code`
int set_error(int code, const std::string &msg);
}  // namespace rtb

#include <new>
#include <stdexcept>
namespace rtb {
// no exception may cross the extern "C" boundary (rtb.h): host entry points run their bodies through this
template <class F>
int host_guarded(int io_code, F f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return set_error(-5 /* RTB_ERR_OOM */, "out of host memory");
    } catch (const std::length_error &) {
        return set_error(-5, "out of host memory (size overflow)");
    } catch (const std::exception &e) {
        return set_error(io_code, e.what());
    }
}
}  // namespace rtb
