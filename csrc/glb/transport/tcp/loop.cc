#include "glb/transport/tcp/loop.h"

#include <fcntl.h>
#include <sys/eventfd.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>

#include "glb/common/logging.h"
#include "glb/common/utils.h"

namespace glb {
namespace transport {
namespace tcp {

// The handlers speak epoll's event bits; poll(2) uses the same values on Linux.
static_assert(EPOLLIN == POLLIN && EPOLLOUT == POLLOUT && EPOLLERR == POLLERR && EPOLLHUP == POLLHUP,
              "event bits of epoll and poll differ");

Loop::Loop(Backend backend) : backend_(backend) {
  if (backend_ == Backend::POLL) {
    int p[2];
    GLB_ENFORCE_NE(::pipe(p), -1, "pipe: ", std::strerror(errno));
    for (int fd : p) {
      ::fcntl(fd, F_SETFL, ::fcntl(fd, F_GETFL) | O_NONBLOCK);
      ::fcntl(fd, F_SETFD, FD_CLOEXEC);
    }
    wakefd_ = p[0];
    wakeWr_ = p[1];
    fds_.push_back({wakefd_, POLLIN, 0});
    handlers_.push_back(nullptr);  // nullptr marks the wake descriptor
    thread_ = std::thread(&Loop::run, this);
    threadId_ = thread_.get_id();
    return;
  }
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
  if (wakeWr_ != -1) ::close(wakeWr_);
  if (epfd_ != -1) ::close(epfd_);
}

void Loop::wake() {
  if (backend_ == Backend::POLL) {
    char c = 1;
    ssize_t rv = ::write(wakeWr_, &c, 1);  // a full pipe already guarantees a wake-up
    (void)rv;
    return;
  }
  uint64_t one = 1;
  ssize_t rv = ::write(wakefd_, &one, sizeof(one));
  (void)rv;
}

void Loop::registerDescriptor(int fd, int events, Handler* h) {
  if (backend_ == Backend::POLL) {
    {
      std::lock_guard<std::mutex> g(fdsMu_);
      size_t i = 0;
      while (i < fds_.size() && fds_[i].fd != fd) i++;
      if (i == fds_.size()) {
        fds_.push_back({fd, static_cast<short>(events), 0});
        handlers_.push_back(h);
      } else {
        fds_[i].events = static_cast<short>(events);
        handlers_[i] = h;
      }
    }
    if (!inLoopThread()) wake();
    return;
  }
  struct epoll_event ev;
  std::memset(&ev, 0, sizeof(ev));
  ev.events = static_cast<uint32_t>(events);
  ev.data.ptr = h;
  int rv = epoll_ctl(epfd_, EPOLL_CTL_ADD, fd, &ev);
  if (rv == -1 && errno == EEXIST) rv = epoll_ctl(epfd_, EPOLL_CTL_MOD, fd, &ev);
  GLB_ENFORCE_NE(rv, -1, "epoll_ctl: ", std::strerror(errno));
}

void Loop::modifyDescriptor(int fd, int events, Handler* h) {
  if (backend_ == Backend::POLL) {
    registerDescriptor(fd, events, h);
    return;
  }
  struct epoll_event ev;
  std::memset(&ev, 0, sizeof(ev));
  ev.events = static_cast<uint32_t>(events);
  ev.data.ptr = h;
  if (epoll_ctl(epfd_, EPOLL_CTL_MOD, fd, &ev) == -1 && errno == ENOENT) registerDescriptor(fd, events, h);
}

void Loop::unregisterDescriptor(int fd, Handler* h) {
  removeDescriptor(fd);
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
  if (backend_ == Backend::POLL) {
    std::lock_guard<std::mutex> g(fdsMu_);
    for (size_t i = 1; i < fds_.size(); i++) {
      if (fds_[i].fd == fd) {
        fds_.erase(fds_.begin() + static_cast<long>(i));
        handlers_.erase(handlers_.begin() + static_cast<long>(i));
        break;
      }
    }
    return;  // the loop thread works on a snapshot; callers that destroy the handler wait for a tick
  }
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

int Loop::waitBatch() {
  if (backend_ == Backend::EPOLL) return epoll_wait(epfd_, batch_, kBatch, 50);
  {
    std::lock_guard<std::mutex> g(fdsMu_);
    snapFds_ = fds_;
    snapHandlers_ = handlers_;
  }
  for (auto& p : snapFds_) p.revents = 0;
  const int ready = ::poll(snapFds_.data(), static_cast<nfds_t>(snapFds_.size()), 50);
  if (ready <= 0) return ready;
  // Same batch format as epoll_wait. A handler whose descriptor was removed after the
  // snapshot must not run: check it is still registered (the set is small).
  int n = 0;
  const size_t total = snapFds_.size();
  const size_t start = pollCursor_++ % total;
  for (size_t k = 0; k < total && n < kBatch; k++) {
    const size_t i = (start + k) % total;
    if (snapFds_[i].revents == 0) continue;
    if (snapFds_[i].revents & POLLNVAL) {  // closed without being removed: epoll forgets such descriptors by itself
      removeDescriptor(snapFds_[i].fd);
      continue;
    }
    if (snapHandlers_[i] != nullptr) {
      std::lock_guard<std::mutex> g(fdsMu_);
      bool live = false;
      for (size_t j = 1; j < fds_.size(); j++) live = live || (fds_[j].fd == snapFds_[i].fd && handlers_[j] == snapHandlers_[i]);
      if (!live) continue;
    }
    batch_[n].events = static_cast<uint32_t>(snapFds_[i].revents);
    batch_[n].data.ptr = snapHandlers_[i];
    n++;
  }
  return n;
}

void Loop::run() {
  setThreadName("glb_tcp_loop");
  while (!done_.load()) {
    int n = waitBatch();
    if (n == -1) {
      if (errno == EINTR) continue;
      GLB_ERROR(backend_ == Backend::EPOLL ? "epoll_wait: " : "poll: ", std::strerror(errno));
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
