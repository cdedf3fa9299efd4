#include "glb/transport/tcp/loop.h"

#include <sys/eventfd.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>

#include "glb/common/logging.h"
#include "glb/common/utils.h"

namespace glb {
namespace transport {
namespace tcp {

Loop::Loop() {
  epfd_ = epoll_create1(EPOLL_CLOEXEC);
  GLB_ENFORCE_NE(epfd_, -1, "epoll_create1: ", std::strerror(errno));
  wakefd_ = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
  GLB_ENFORCE_NE(wakefd_, -1, "eventfd: ", std::strerror(errno));
  struct epoll_event ev;
  std::memset(&ev, 0, sizeof(ev));
  ev.events = EPOLLIN;
  ev.data.ptr = nullptr;  // nullptr marks the wake fd
  GLB_ENFORCE_NE(epoll_ctl(epfd_, EPOLL_CTL_ADD, wakefd_, &ev), -1, "epoll_ctl: ", std::strerror(errno));
  thread_ = std::thread(&Loop::run, this);
  threadId_ = thread_.get_id();
}

Loop::~Loop() {
  done_.store(true);
  wake();
  if (thread_.joinable()) thread_.join();
  ::close(wakefd_);
  ::close(epfd_);
}

void Loop::wake() {
  uint64_t one = 1;
  ssize_t rv = ::write(wakefd_, &one, sizeof(one));
  (void)rv;
}

void Loop::registerDescriptor(int fd, int events, Handler* h) {
  struct epoll_event ev;
  std::memset(&ev, 0, sizeof(ev));
  ev.events = static_cast<uint32_t>(events);
  ev.data.ptr = h;
  int rv = epoll_ctl(epfd_, EPOLL_CTL_ADD, fd, &ev);
  if (rv == -1 && errno == EEXIST) rv = epoll_ctl(epfd_, EPOLL_CTL_MOD, fd, &ev);
  GLB_ENFORCE_NE(rv, -1, "epoll_ctl: ", std::strerror(errno));
}

void Loop::modifyDescriptor(int fd, int events, Handler* h) {
  struct epoll_event ev;
  std::memset(&ev, 0, sizeof(ev));
  ev.events = static_cast<uint32_t>(events);
  ev.data.ptr = h;
  if (epoll_ctl(epfd_, EPOLL_CTL_MOD, fd, &ev) == -1 && errno == ENOENT) registerDescriptor(fd, events, h);
}

void Loop::unregisterDescriptor(int fd, Handler* h) {
  int rv = epoll_ctl(epfd_, EPOLL_CTL_DEL, fd, nullptr);
  if (rv == -1 && errno != ENOENT && errno != EBADF) {
    GLB_WARN("epoll_ctl(DEL): ", std::strerror(errno));
  }
  if (inLoopThread()) {
    // Drop events for this handler that are still queued in the current batch.
    for (int i = batchPos_ + 1; i < batchSize_; i++) {
      if (batch_[i].data.ptr == h) batch_[i].data.ptr = reinterpret_cast<void*>(1);
    }
    return;
  }
  // Wait for the loop to finish the batch it may be in the middle of.
  std::unique_lock<std::mutex> g(mu_);
  uint64_t t = tick_;
  wake();
  cv_.wait(g, [&] { return tick_ != t || done_.load(); });
}

void Loop::removeDescriptor(int fd) {
  int rv = epoll_ctl(epfd_, EPOLL_CTL_DEL, fd, nullptr);
  if (rv == -1 && errno != ENOENT && errno != EBADF) {
    GLB_WARN("epoll_ctl(DEL): ", std::strerror(errno));
  }
}

void Loop::barrier() {
  if (inLoopThread()) return;
  std::unique_lock<std::mutex> g(mu_);
  uint64_t t = tick_;
  wake();
  cv_.wait(g, [&] { return tick_ != t || done_.load(); });
}

void Loop::defer(std::function<void()> fn) {
  {
    std::lock_guard<std::mutex> g(mu_);
    deferred_.push_back(std::move(fn));
  }
  wake();
}

void Loop::run() {
  setThreadName("glb_tcp_loop");
  while (!done_.load()) {
    int n = epoll_wait(epfd_, batch_, kBatch, 50);
    if (n == -1) {
      if (errno == EINTR) continue;
      GLB_ERROR("epoll_wait: ", std::strerror(errno));
      break;
    }
    batchSize_ = n;
    for (batchPos_ = 0; batchPos_ < batchSize_; batchPos_++) {
      void* p = batch_[batchPos_].data.ptr;
      if (p == nullptr) {
        uint64_t v;
        while (::read(wakefd_, &v, sizeof(v)) > 0) {
        }
        continue;
      }
      if (p == reinterpret_cast<void*>(1)) continue;  // cancelled within this batch
      static_cast<Handler*>(p)->handleEvents(static_cast<int>(batch_[batchPos_].events));
    }
    batchSize_ = 0;
    batchPos_ = 0;

    std::list<std::function<void()>> fns;
    {
      std::lock_guard<std::mutex> g(mu_);
      fns.swap(deferred_);
      tick_++;
    }
    cv_.notify_all();
    for (auto& fn : fns) fn();
  }
  std::lock_guard<std::mutex> g(mu_);
  tick_++;
  cv_.notify_all();
}

}  // namespace tcp
}  // namespace transport
}  // namespace glb
