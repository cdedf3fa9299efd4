// Store backed by a Redis server. Speaks RESP directly over a TCP socket (SETNX /
// GET / EXISTS / APPEND / INCRBY / MGET), so there is no hiredis dependency.
// Write-once keys via SETNX; wait() polls EXISTS like the reference.
// Parity: gloo/rendezvous/redis_store.{h,cc}.
#pragma once

#include <mutex>

#include "glb/rendezvous/store.h"

namespace glb {
namespace rendezvous {

class RedisStore : public Store {
 public:
  explicit RedisStore(const std::string& host, int port = 6379);
  ~RedisStore() override;

  void set(const std::string& key, const Bytes& data) override;
  Bytes get(const std::string& key) override;
  void wait(const std::vector<std::string>& keys, std::chrono::milliseconds timeout) override;
  using Store::wait;
  bool check(const std::vector<std::string>& keys);

  bool has_extended_api() const override { return true; }
  std::vector<Bytes> multi_get(const std::vector<std::string>& keys) override;
  void multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) override;
  void append(const std::string& key, const Bytes& data) override;
  int64_t add(const std::string& key, int64_t value) override;

 private:
  struct Reply {
    char type = 0;  // '+', '-', ':', '$', '*'
    std::string str;
    int64_t integer = 0;
    bool nil = false;
    std::vector<Reply> elems;
  };
  Reply command(const std::vector<std::string>& args);
  Reply readReply();
  std::string readLine();
  void readExact(char* dst, size_t n);

  std::string host_;
  int port_;
  int fd_ = -1;
  std::mutex mu_;
  std::string rbuf_;
};

}  // namespace rendezvous
}  // namespace glb
