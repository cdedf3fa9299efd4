// epoll event loop running on its own thread ("glb_tcp_loop"). Handlers are raw
// pointers; unregisterDescriptor() guarantees that, once it returns, the handler
// is not executing and will not be invoked again, so the caller may destroy it.
// `defer` runs a closure on the loop thread (woken through an eventfd).
// Parity: gloo/transport/tcp/loop.{h,cc} (epoll thread, Deferrables, the
// unregister-waits-for-tick rule).
#pragma once

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
  Loop();
  ~Loop();
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

  int epfd_ = -1;
  int wakefd_ = -1;
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
