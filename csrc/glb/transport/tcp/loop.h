// Event loop running on its own thread ("glb_tcp_loop"). Two readiness backends behind the
// same interface: epoll (default) and a portable poll(2) reactor (pipe wake-up, no
// Linux-only calls) that the "uv" device selects. Handlers are raw
// pointers; unregisterDescriptor() guarantees that, once it returns, the handler
// is not executing and will not be invoked again, so the caller may destroy it.
// `defer` runs a closure on the loop thread (woken through an eventfd).
// Parity: gloo/transport/tcp/loop.{h,cc} (epoll thread, Deferrables, the
// unregister-waits-for-tick rule).
#pragma once

#include <poll.h>
#include <sys/epoll.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <list>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace glb {
namespace transport {
namespace tcp {

class Handler {
 public:
  virtual ~Handler() = default;
  virtual void handleEvents(int events) = 0;
};

class Loop final {
 public:
  enum class Backend { EPOLL, POLL };
  explicit Loop(Backend backend = Backend::EPOLL);
  ~Loop();
  Backend backend() const { return backend_; }
  Loop(const Loop&) = delete;
  Loop& operator=(const Loop&) = delete;

  void registerDescriptor(int fd, int events, Handler* h);
  // Change the interest set of an already registered descriptor (one syscall).
  void modifyDescriptor(int fd, int events, Handler* h);
  void unregisterDescriptor(int fd, Handler* h);
  // Non-waiting variant: stops future events but a handler call that is already
  // queued in the current batch may still happen. Pair it with barrier() before
  // destroying the handler.
  void removeDescriptor(int fd);
  // Returns once the loop has completed the batch it was in when called.
  void barrier();
  void defer(std::function<void()> fn);
  bool inLoopThread() const { return std::this_thread::get_id() == threadId_; }

 private:
  static constexpr int kBatch = 64;
  void run();
  void wake();

  int waitBatch();  // fills batch_, returns the number of entries or -1

  const Backend backend_;
  int epfd_ = -1;
  int wakefd_ = -1;   // eventfd (epoll) or the read end of the wake pipe (poll)
  int wakeWr_ = -1;   // write end of the wake pipe (poll backend)
  // poll backend: the interest set, edited under fdsMu_ from any thread; the loop thread
  // polls a snapshot and is woken after every edit.
  std::mutex fdsMu_;
  std::vector<struct pollfd> fds_;
  std::vector<Handler*> handlers_;
  std::vector<struct pollfd> snapFds_;
  std::vector<Handler*> snapHandlers_;
  size_t pollCursor_ = 0;  // rotate the start so that a busy descriptor cannot starve the others
  std::atomic<bool> done_{false};
  std::thread thread_;
  std::thread::id threadId_;

  std::mutex mu_;
  std::condition_variable cv_;
  uint64_t tick_ = 0;
  std::list<std::function<void()>> deferred_;

  // Current batch, visible to unregisterDescriptor when it runs on the loop thread.
  struct epoll_event batch_[kBatch];
  int batchSize_ = 0;
  int batchPos_ = 0;
};

}  // namespace tcp
}  // namespace transport
}  // namespace glb
