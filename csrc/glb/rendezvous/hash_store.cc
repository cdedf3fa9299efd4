#include "glb/rendezvous/hash_store.h"

#include <cstring>

#include "glb/common/logging.h"

namespace glb {
namespace rendezvous {

void HashStore::set(const std::string& key, const Bytes& data) {
  std::lock_guard<std::mutex> g(mu_);
  GLB_ENFORCE(map_.find(key) == map_.end(), "Key '", key, "' already set (keys are write-once)");
  map_[key] = data;
  cv_.notify_all();
}

IStore::Bytes HashStore::get(const std::string& key) {
  std::unique_lock<std::mutex> g(mu_);
  auto pred = [&] { return map_.find(key) != map_.end(); };
  if (!cv_.wait_for(g, kDefaultTimeout, pred)) {
    GLB_THROW_IO_EXCEPTION("Wait timeout for key: ", key);
  }
  return map_[key];
}

void HashStore::wait(const std::vector<std::string>& keys, std::chrono::milliseconds timeout) {
  std::unique_lock<std::mutex> g(mu_);
  auto pred = [&] {
    for (const auto& k : keys) {
      if (map_.find(k) == map_.end()) return false;
    }
    return true;
  };
  if (timeout == kNoTimeout) {
    cv_.wait(g, pred);
    return;
  }
  if (!cv_.wait_for(g, timeout, pred)) {
    std::string missing;
    for (const auto& k : keys) {
      if (map_.find(k) == map_.end()) missing += (missing.empty() ? "" : ", ") + k;
    }
    GLB_THROW_IO_EXCEPTION("Wait timeout for key(s): [", missing, "]");
  }
}

std::vector<IStore::Bytes> HashStore::multi_get(const std::vector<std::string>& keys) {
  wait(keys, kDefaultTimeout);
  std::lock_guard<std::mutex> g(mu_);
  std::vector<Bytes> out;
  out.reserve(keys.size());
  for (const auto& k : keys) out.push_back(map_[k]);
  return out;
}

void HashStore::multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) {
  GLB_ENFORCE_EQ(keys.size(), values.size());
  std::lock_guard<std::mutex> g(mu_);
  for (size_t i = 0; i < keys.size(); i++) {
    GLB_ENFORCE(map_.find(keys[i]) == map_.end(), "Key '", keys[i], "' already set");
    map_[keys[i]] = values[i];
  }
  cv_.notify_all();
}

void HashStore::append(const std::string& key, const Bytes& data) {
  std::lock_guard<std::mutex> g(mu_);
  auto& v = map_[key];
  v.insert(v.end(), data.begin(), data.end());
  mutable_[key] = true;
  cv_.notify_all();
}

int64_t HashStore::add(const std::string& key, int64_t value) {
  std::lock_guard<std::mutex> g(mu_);
  auto& v = map_[key];
  int64_t cur = 0;
  if (v.size() == sizeof(int64_t)) std::memcpy(&cur, v.data(), sizeof(cur));
  cur += value;
  v.resize(sizeof(int64_t));
  std::memcpy(v.data(), &cur, sizeof(cur));
  mutable_[key] = true;
  cv_.notify_all();
  return cur;
}

size_t HashStore::size() const {
  std::lock_guard<std::mutex> g(mu_);
  return map_.size();
}

}  // namespace rendezvous
}  // namespace glb
