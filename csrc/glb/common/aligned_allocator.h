// STL allocator handing out 64-byte (cache line / AVX-512) aligned storage.
// Parity: gloo/common/aligned_allocator.h (32-byte posix_memalign allocator).
#pragma once

#include <cstddef>
#include <cstdlib>
#include <new>

namespace glb {

constexpr size_t kBufferAlignment = 64;

template <typename T, size_t Alignment = kBufferAlignment>
class aligned_allocator {
 public:
  using value_type = T;
  template <typename U>
  struct rebind {
    using other = aligned_allocator<U, Alignment>;
  };

  aligned_allocator() noexcept = default;
  template <typename U>
  aligned_allocator(const aligned_allocator<U, Alignment>&) noexcept {}

  T* allocate(size_t n) {
    void* p = nullptr;
    size_t bytes = n * sizeof(T);
    if (bytes == 0) bytes = Alignment;
    if (posix_memalign(&p, Alignment, bytes) != 0) throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate(T* p, size_t) noexcept { std::free(p); }

  template <typename U>
  bool operator==(const aligned_allocator<U, Alignment>&) const noexcept { return true; }
  template <typename U>
  bool operator!=(const aligned_allocator<U, Alignment>&) const noexcept { return false; }
};

}  // namespace glb
