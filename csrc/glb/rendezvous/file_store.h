// Store backed by a shared directory (works across processes and, on a shared
// filesystem, across hosts). One file per key, named by a 64-bit FNV-1a hash of
// the key followed by the escaped key tail for debuggability; written to a
// temporary name and atomically renamed so readers never see partial values.
// `add`/`append` take an flock on a per-key lock file.
// Parity: gloo/rendezvous/file_store.{h,cc}.
#pragma once

#include <mutex>

#include "glb/rendezvous/store.h"

namespace glb {
namespace rendezvous {

class FileStore : public Store {
 public:
  explicit FileStore(const std::string& path);

  void set(const std::string& key, const Bytes& data) override;
  Bytes get(const std::string& key) override;
  void wait(const std::vector<std::string>& keys, std::chrono::milliseconds timeout) override;
  using Store::wait;

  bool has_extended_api() const override { return true; }
  std::vector<Bytes> multi_get(const std::vector<std::string>& keys) override;
  void multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) override;
  void append(const std::string& key, const Bytes& data) override;
  int64_t add(const std::string& key, int64_t value) override;

  // Paths of every key file this instance created (for cleanup by tests / benchmarks).
  std::vector<std::string> getAllKeyFilePaths() const;
  const std::string& basePath() const { return base_; }

  static std::string keyFileName(const std::string& key);

 private:
  std::string objectPath(const std::string& key) const;
  bool exists(const std::string& path) const;
  void writeAtomic(const std::string& path, const Bytes& data, bool exclusive);

  std::string base_;
  mutable std::mutex mu_;
  std::vector<std::string> created_;
};

}  // namespace rendezvous
}  // namespace glb
