// Transport-level error descriptors carried through to IoException messages.
// Parity: gloo/transport/tcp/error.{h,cc}.
#pragma once

#include <cstring>
#include <string>

#include "glb/common/string.h"

namespace glb {
namespace transport {
namespace tcp {

class Error {
 public:
  Error() : valid_(false) {}
  explicit Error(std::string what) : valid_(true), what_(std::move(what)) {}
  virtual ~Error() = default;
  explicit operator bool() const { return valid_; }
  virtual std::string what() const { return what_; }

 private:
  bool valid_;
  std::string what_;
};

class SystemError : public Error {
 public:
  SystemError(const char* syscall, int err, const std::string& remote = "")
      : Error(strcat_all(syscall, ": ", std::strerror(err), remote.empty() ? "" : " (peer " + remote + ")")),
        error_(err) {}
  int error() const { return error_; }

 private:
  int error_;
};

class ShortReadError : public Error {
 public:
  ShortReadError(size_t expected, size_t actual)
      : Error(strcat_all("short read (got ", actual, " of ", expected, " bytes)")) {}
};

class ShortWriteError : public Error {
 public:
  ShortWriteError(size_t expected, size_t actual)
      : Error(strcat_all("short write (wrote ", actual, " of ", expected, " bytes)")) {}
};

class TimeoutError : public Error {
 public:
  explicit TimeoutError(const std::string& msg) : Error(msg) {}
};

class LoopError : public Error {
 public:
  explicit LoopError(const std::string& msg) : Error(msg) {}
};

}  // namespace tcp
}  // namespace transport
}  // namespace glb
