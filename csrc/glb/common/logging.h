// Levelled stderr logger + ENFORCE macros.
//
// Level comes from env GLB_LOG_LEVEL (GLOO_LOG_LEVEL is honoured as an alias so
// existing launch scripts keep working): ERROR, WARN (default), INFO, DEBUG.
// Parity: gloo/common/logging.h:40-208, logging.cc:19-44.
#pragma once

#include <atomic>
#include <sstream>
#include <string>
#include <vector>

#include "glb/common/error.h"
#include "glb/common/string.h"

namespace glb {

enum class LogLevel : int { ERROR = 0, WARN = 1, INFO = 2, DEBUG = 3 };

LogLevel logLevel();
void setLogLevel(LogLevel level);
void logMessage(LogLevel level, const char* file, int line, const std::string& msg);

// Thrown by GLB_ENFORCE*. Keeps the failing expression and a message stack so
// callers can append context while unwinding.
class EnforceNotMet : public Exception {
 public:
  EnforceNotMet(const char* file, int line, const char* cond, const std::string& msg);
  void appendMessage(const std::string& msg);
  const std::vector<std::string>& messageStack() const { return stack_; }
  const char* what() const noexcept override { return full_.c_str(); }

 private:
  std::vector<std::string> stack_;
  std::string full_;
  void rebuild();
};

namespace detail {
template <typename A, typename B>
inline std::string cmpMessage(const char* op, const A& a, const B& b) {
  std::ostringstream os;
  os << a << " " << op << " " << b;
  return os.str();
}
}  // namespace detail

}  // namespace glb

#define GLB_LOG(level, ...)                                                  \
  do {                                                                       \
    if (static_cast<int>(::glb::logLevel()) >= static_cast<int>(level)) {    \
      ::glb::logMessage(level, __FILE__, __LINE__,                           \
                        ::glb::strcat_all(__VA_ARGS__));                     \
    }                                                                        \
  } while (0)

#define GLB_ERROR(...) GLB_LOG(::glb::LogLevel::ERROR, __VA_ARGS__)
#define GLB_WARN(...) GLB_LOG(::glb::LogLevel::WARN, __VA_ARGS__)
#define GLB_INFO(...) GLB_LOG(::glb::LogLevel::INFO, __VA_ARGS__)
#define GLB_DEBUG(...) GLB_LOG(::glb::LogLevel::DEBUG, __VA_ARGS__)

#define GLB_ENFORCE(cond, ...)                                              \
  do {                                                                      \
    if (!(cond)) {                                                          \
      throw ::glb::EnforceNotMet(__FILE__, __LINE__, #cond,                 \
                                 ::glb::strcat_all(__VA_ARGS__));           \
    }                                                                       \
  } while (0)

#define GLB_ENFORCE_CMP(op, a, b, ...)                                       \
  do {                                                                       \
    const auto& glb_a_ = (a);                                                \
    const auto& glb_b_ = (b);                                                \
    if (!(glb_a_ op glb_b_)) {                                               \
      throw ::glb::EnforceNotMet(                                            \
          __FILE__, __LINE__, #a " " #op " " #b,                             \
          ::glb::strcat_all(::glb::detail::cmpMessage(#op, glb_a_, glb_b_),  \
                            ". ", ::glb::strcat_all(__VA_ARGS__)));          \
    }                                                                        \
  } while (0)

#define GLB_ENFORCE_EQ(a, b, ...) GLB_ENFORCE_CMP(==, a, b, __VA_ARGS__)
#define GLB_ENFORCE_NE(a, b, ...) GLB_ENFORCE_CMP(!=, a, b, __VA_ARGS__)
#define GLB_ENFORCE_LE(a, b, ...) GLB_ENFORCE_CMP(<=, a, b, __VA_ARGS__)
#define GLB_ENFORCE_LT(a, b, ...) GLB_ENFORCE_CMP(<, a, b, __VA_ARGS__)
#define GLB_ENFORCE_GE(a, b, ...) GLB_ENFORCE_CMP(>=, a, b, __VA_ARGS__)
#define GLB_ENFORCE_GT(a, b, ...) GLB_ENFORCE_CMP(>, a, b, __VA_ARGS__)
