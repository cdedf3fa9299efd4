#include "glb/common/logging.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace glb {

namespace {

LogLevel parseLevel(const char* s) {
  if (s == nullptr) return LogLevel::WARN;
  std::string v(s);
  for (auto& c : v) c = static_cast<char>(std::toupper(c));
  if (v == "ERROR") return LogLevel::ERROR;
  if (v == "WARN" || v == "WARNING") return LogLevel::WARN;
  if (v == "INFO") return LogLevel::INFO;
  if (v == "DEBUG" || v == "TRACE") return LogLevel::DEBUG;
  return LogLevel::WARN;
}

std::atomic<int>& levelStorage() {
  static std::atomic<int> lvl{[] {
    const char* e = std::getenv("GLB_LOG_LEVEL");
    if (e == nullptr) e = std::getenv("GLOO_LOG_LEVEL");
    return static_cast<int>(parseLevel(e));
  }()};
  return lvl;
}

const char* levelName(LogLevel l) {
  switch (l) {
    case LogLevel::ERROR: return "ERROR";
    case LogLevel::WARN: return "WARN";
    case LogLevel::INFO: return "INFO";
    case LogLevel::DEBUG: return "DEBUG";
  }
  return "?";
}

}  // namespace

LogLevel logLevel() { return static_cast<LogLevel>(levelStorage().load(std::memory_order_relaxed)); }
void setLogLevel(LogLevel level) { levelStorage().store(static_cast<int>(level)); }

void logMessage(LogLevel level, const char* file, int line, const std::string& msg) {
  static std::mutex mu;  // keep lines from different threads intact
  const char* base = std::strrchr(file, '/');
  base = base ? base + 1 : file;
  auto now = std::chrono::duration_cast<std::chrono::microseconds>(
                 std::chrono::system_clock::now().time_since_epoch()).count();
  std::lock_guard<std::mutex> g(mu);
  std::fprintf(stderr, "[glb %s %lld.%06lld %s:%d] %s\n", levelName(level),
               static_cast<long long>(now / 1000000), static_cast<long long>(now % 1000000),
               base, line, msg.c_str());
}

EnforceNotMet::EnforceNotMet(const char* file, int line, const char* cond, const std::string& msg)
    : Exception("") {
  stack_.push_back(strcat_all("[enforce fail at ", file, ":", line, "] ", cond, ". ", msg));
  rebuild();
}

void EnforceNotMet::appendMessage(const std::string& msg) {
  stack_.push_back(msg);
  rebuild();
}

void EnforceNotMet::rebuild() {
  full_.clear();
  for (size_t i = 0; i < stack_.size(); i++) {
    if (i) full_ += "\n  ";
    full_ += stack_[i];
  }
}

}  // namespace glb
