// Namespacing decorator: every key becomes "<prefix>/<key>".
// Parity: gloo/rendezvous/prefix_store.{h,cc}.
#pragma once

#include <memory>

#include "glb/rendezvous/store.h"

namespace glb {
namespace rendezvous {

class PrefixStore : public Store {
 public:
  PrefixStore(const std::string& prefix, std::shared_ptr<Store> store)
      : prefix_(prefix), store_(std::move(store)) {}

  void set(const std::string& key, const Bytes& data) override { store_->set(join(key), data); }
  Bytes get(const std::string& key) override { return store_->get(join(key)); }
  Bytes wait_get(const std::string& key, std::chrono::milliseconds timeout) override {
    return store_->wait_get(join(key), timeout);
  }
  void wait(const std::vector<std::string>& keys, std::chrono::milliseconds timeout) override {
    store_->wait(joinAll(keys), timeout);
  }
  using Store::wait;

  bool has_extended_api() const override { return store_->has_extended_api(); }
  std::vector<Bytes> multi_get(const std::vector<std::string>& keys) override {
    return store_->multi_get(joinAll(keys));
  }
  void multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) override {
    store_->multi_set(joinAll(keys), values);
  }
  void append(const std::string& key, const Bytes& data) override { store_->append(join(key), data); }
  int64_t add(const std::string& key, int64_t value) override { return store_->add(join(key), value); }

  const std::string& prefix() const { return prefix_; }

 private:
  std::string join(const std::string& key) const { return prefix_ + "/" + key; }
  std::vector<std::string> joinAll(const std::vector<std::string>& keys) const {
    std::vector<std::string> out;
    out.reserve(keys.size());
    for (const auto& k : keys) out.push_back(join(k));
    return out;
  }
  const std::string prefix_;
  std::shared_ptr<Store> store_;
};

}  // namespace rendezvous
}  // namespace glb
