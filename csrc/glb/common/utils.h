// Small process-level helpers and environment flags.
// Parity: gloo/common/utils.{h,cc} (getHostname, setThreadName, env flags).
#pragma once

#include <chrono>
#include <cstdint>
#include <string>

namespace glb {

std::string getHostname();
void setThreadName(const std::string& name);  // truncated to 15 chars (Linux limit)

// Generic env accessors. `GLB_<X>` wins; `GLOO_<X>` is accepted as an alias.
bool envFlag(const char* name, bool dflt);
long envInt(const char* name, long dflt);
std::string envStr(const char* name, const std::string& dflt);

// Flags with the same meaning as the reference's (utils.cc:40-60).
bool useRankAsSeqNumber();      // *_ENABLE_RANK_AS_SEQUENCE_NUMBER
bool isStoreExtendedApiEnabled();  // *_ENABLE_STORE_V2_API
bool disableConnectionRetries();   // *_DISABLE_CONNECTION_RETRIES

inline uint64_t nowNs() {
  return static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(
                                   std::chrono::steady_clock::now().time_since_epoch())
                                   .count());
}

inline size_t roundUp(size_t v, size_t m) { return (v + m - 1) / m * m; }
inline size_t ceilDiv(size_t a, size_t b) { return (a + b - 1) / b; }
inline uint32_t log2ceil(uint32_t v) {
  uint32_t r = 0;
  while ((1u << r) < v) r++;
  return r;
}
inline bool isPow2(uint64_t v) { return v && !(v & (v - 1)); }

}  // namespace glb
