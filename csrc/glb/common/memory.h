// Lifetime helpers for objects that are owned by user code (often on the stack)
// but touched by I/O threads: the owner holds an `Anchor<T>`, worker threads
// obtain short-lived `Lease<T>`s from a `WeakAnchor<T>`. Destroying the anchor
// blocks until every outstanding lease is gone, after which leases can no longer
// be obtained — so a worker never dereferences a dead object.
//
// Same problem the reference solves with ShareableNonOwningPtr / WeakNonOwningPtr
// / NonOwningPtr (gloo/common/memory.h:75-154); this design uses a condvar rather
// than a spin so a destructor racing a slow completion handler sleeps.
#pragma once

#include <condition_variable>
#include <memory>
#include <mutex>

namespace glb {

namespace detail {
template <typename T>
struct AnchorState {
  std::mutex mu;
  std::condition_variable cv;
  T* ptr = nullptr;
  int leases = 0;
};
}  // namespace detail

template <typename T>
class Lease {
 public:
  Lease() = default;
  explicit Lease(std::shared_ptr<detail::AnchorState<T>> s) : s_(std::move(s)) {}
  Lease(Lease&& o) noexcept : s_(std::move(o.s_)), p_(o.p_) { o.p_ = nullptr; }
  Lease& operator=(Lease&& o) noexcept {
    release();
    s_ = std::move(o.s_);
    p_ = o.p_;
    o.p_ = nullptr;
    return *this;
  }
  Lease(const Lease&) = delete;
  Lease& operator=(const Lease&) = delete;
  ~Lease() { release(); }

  explicit operator bool() const { return p_ != nullptr; }
  T* get() const { return p_; }
  T* operator->() const { return p_; }
  T& operator*() const { return *p_; }

  void release() {
    if (p_ != nullptr) {
      std::lock_guard<std::mutex> g(s_->mu);
      if (--s_->leases == 0) s_->cv.notify_all();
      p_ = nullptr;
    }
    s_.reset();
  }

 private:
  template <typename U>
  friend class WeakAnchor;
  std::shared_ptr<detail::AnchorState<T>> s_;
  T* p_ = nullptr;
};

template <typename T>
class WeakAnchor {
 public:
  WeakAnchor() = default;
  explicit WeakAnchor(std::shared_ptr<detail::AnchorState<T>> s) : s_(std::move(s)) {}

  // Empty lease if the anchor has been destroyed.
  Lease<T> lock() const {
    Lease<T> l;
    if (!s_) return l;
    std::lock_guard<std::mutex> g(s_->mu);
    if (s_->ptr == nullptr) return l;
    s_->leases++;
    l.s_ = s_;
    l.p_ = s_->ptr;
    return l;
  }
  bool expired() const {
    if (!s_) return true;
    std::lock_guard<std::mutex> g(s_->mu);
    return s_->ptr == nullptr;
  }

 private:
  std::shared_ptr<detail::AnchorState<T>> s_;
};

template <typename T>
class Anchor {
 public:
  explicit Anchor(T* self) : s_(std::make_shared<detail::AnchorState<T>>()) { s_->ptr = self; }
  Anchor(const Anchor&) = delete;
  Anchor& operator=(const Anchor&) = delete;
  ~Anchor() { retire(); }

  WeakAnchor<T> weak() const { return WeakAnchor<T>(s_); }

  // Stop handing out leases and wait until the outstanding ones are released.
  void retire() {
    std::unique_lock<std::mutex> g(s_->mu);
    s_->ptr = nullptr;
    s_->cv.wait(g, [&] { return s_->leases == 0; });
  }

 private:
  std::shared_ptr<detail::AnchorState<T>> s_;
};

}  // namespace glb
