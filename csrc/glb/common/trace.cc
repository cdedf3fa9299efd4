#include "glb/common/trace.h"

#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "glb/common/utils.h"

namespace glb {
namespace trace {

namespace {
struct Event {
  const char* name;
  uint64_t start, end;
  long tid;
};

struct Sink {
  std::mutex mu;
  std::vector<Event> events;
  std::string path;
  bool wroteHeader = false;
  std::atomic<int> state{0};  // 0 = not looked at the environment yet, 1 = off, 2 = on

  ~Sink() { flushLocked(true); }

  size_t flushLocked(bool final) {
    std::lock_guard<std::mutex> g(mu);
    if (path.empty() || (events.empty() && !(final && wroteHeader))) return 0;
    FILE* f = std::fopen(path.c_str(), wroteHeader ? "a" : "w");
    if (f == nullptr) return 0;
    if (!wroteHeader) {
      std::fputs("[\n", f);  // the trailing "]" is optional in the trace-event format: appending stays valid
      wroteHeader = true;
    }
    const long pid = static_cast<long>(::getpid());
    for (const Event& e : events) {
      std::fprintf(f, "{\"name\":\"%s\",\"cat\":\"glb\",\"ph\":\"X\",\"ts\":%.3f,\"dur\":%.3f,\"pid\":%ld,\"tid\":%ld},\n", e.name,
                   static_cast<double>(e.start) / 1e3, static_cast<double>(e.end - e.start) / 1e3, pid, e.tid);
    }
    const size_t n = events.size();
    events.clear();
    std::fclose(f);
    return n;
  }
};

Sink& sink() {
  static Sink s;
  return s;
}

int init() {
  Sink& s = sink();
  std::lock_guard<std::mutex> g(s.mu);
  int st = s.state.load(std::memory_order_relaxed);
  if (st != 0) return st;
  std::string p = envStr("TRACE_FILE", "");
  if (p.empty()) {
    s.state.store(1, std::memory_order_release);
    return 1;
  }
  const size_t at = p.find("%r");
  if (at != std::string::npos) p.replace(at, 2, std::to_string(static_cast<long>(::getpid())));
  s.path = p;
  s.events.reserve(4096);
  s.state.store(2, std::memory_order_release);
  return 2;
}
}  // namespace

bool enabled() {
  const int st = sink().state.load(std::memory_order_acquire);
  return (st != 0 ? st : init()) == 2;
}

uint64_t nowNs() {
  return static_cast<uint64_t>(
      std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now().time_since_epoch()).count());
}

void record(const char* name, uint64_t startNs, uint64_t endNs) {
  Sink& s = sink();
  bool full = false;
  {
    std::lock_guard<std::mutex> g(s.mu);
    s.events.push_back(Event{name, startNs, endNs, static_cast<long>(::syscall(SYS_gettid))});
    full = s.events.size() >= 65536;
  }
  if (full) s.flushLocked(false);
}

size_t flush() { return sink().flushLocked(false); }

}  // namespace trace
}  // namespace glb
