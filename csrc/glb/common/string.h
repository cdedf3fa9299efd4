// Variadic string building helper used by logging / error macros.
// Parity: gloo/common/string.h (MakeString).
#pragma once

#include <sstream>
#include <string>

namespace glb {

template <typename... Args>
inline std::string strcat_all(const Args&... args) {
  if constexpr (sizeof...(Args) == 0) {
    return std::string();
  } else {
    std::ostringstream os;
    (os << ... << args);
    return os.str();
  }
}

inline std::string strcat_all(const std::string& s) { return s; }
inline std::string strcat_all(const char* s) { return std::string(s); }

}  // namespace glb
