// Key/value store interface used for rendezvous.
//
// Basic contract: `set` publishes bytes under a key, `get` blocks until the key
// exists (or the store's timeout expires -> IoException), `wait` blocks until all
// keys exist. The "extended" operations (multi_get / multi_set / append / add)
// let the TCP context batch its O(P) lookups; stores that lack them report
// has_extended_api() == false and the callers fall back to the basic ones.
// Parity: gloo/common/store.h:20-53, gloo/rendezvous/store.h:25-74.
#pragma once

#include <chrono>
#include <cstdint>
#include <string>
#include <vector>

#include "glb/common/error.h"

namespace glb {

class IStore {
 public:
  using Bytes = std::vector<char>;
  static constexpr std::chrono::milliseconds kDefaultTimeout = std::chrono::seconds(30);

  virtual ~IStore() = default;

  virtual void set(const std::string& key, const Bytes& data) = 0;
  virtual Bytes get(const std::string& key) = 0;
  virtual Bytes wait_get(const std::string& key, std::chrono::milliseconds timeout) {
    wait({key}, timeout);
    return get(key);
  }
  virtual void wait(const std::vector<std::string>& keys) { wait(keys, kDefaultTimeout); }
  virtual void wait(const std::vector<std::string>& keys, std::chrono::milliseconds timeout) = 0;

  // --- extended ("v2") API -------------------------------------------------
  virtual bool has_extended_api() const { return false; }
  bool has_v2_support() const { return has_extended_api(); }  // reference spelling
  virtual std::vector<Bytes> multi_get(const std::vector<std::string>& keys) {
    (void)keys;
    GLB_THROW_INVALID_OPERATION_EXCEPTION("multi_get not supported by this store");
  }
  virtual void multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) {
    (void)keys;
    (void)values;
    GLB_THROW_INVALID_OPERATION_EXCEPTION("multi_set not supported by this store");
  }
  virtual void append(const std::string& key, const Bytes& data) {
    (void)key;
    (void)data;
    GLB_THROW_INVALID_OPERATION_EXCEPTION("append not supported by this store");
  }
  virtual int64_t add(const std::string& key, int64_t value) {
    (void)key;
    (void)value;
    GLB_THROW_INVALID_OPERATION_EXCEPTION("add not supported by this store");
  }
};

inline IStore::Bytes toBytes(const std::string& s) { return IStore::Bytes(s.begin(), s.end()); }
inline std::string toString(const IStore::Bytes& b) { return std::string(b.begin(), b.end()); }

}  // namespace glb
