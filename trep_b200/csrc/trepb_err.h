// Thread-local error text behind trepb_last_error() (shared by every translation unit of the library).
#pragma once
#include <string>
namespace trepb {
inline std::string& last_error() {
    static thread_local std::string e;
    return e;
}
}  // namespace trepb
