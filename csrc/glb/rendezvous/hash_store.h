// In-process store (map + condvar). Threads-as-ranks tests rendezvous through it.
// Write-once keys, like the reference (hash_store.cc:17-22); implements the
// extended API natively. Parity: gloo/rendezvous/hash_store.{h,cc}.
#pragma once

#include <condition_variable>
#include <mutex>
#include <unordered_map>

#include "glb/rendezvous/store.h"

namespace glb {
namespace rendezvous {

class HashStore : public Store {
 public:
  void set(const std::string& key, const Bytes& data) override;
  Bytes get(const std::string& key) override;
  void wait(const std::vector<std::string>& keys, std::chrono::milliseconds timeout) override;
  using Store::wait;

  bool has_extended_api() const override { return true; }
  std::vector<Bytes> multi_get(const std::vector<std::string>& keys) override;
  void multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) override;
  void append(const std::string& key, const Bytes& data) override;
  int64_t add(const std::string& key, int64_t value) override;

  size_t size() const;

 private:
  mutable std::mutex mu_;
  std::condition_variable cv_;
  std::unordered_map<std::string, Bytes> map_;
  std::unordered_map<std::string, bool> mutable_;  // keys created by append/add
};

}  // namespace rendezvous
}  // namespace glb
