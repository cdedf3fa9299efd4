#include "glb/common/utils.h"

#include <pthread.h>
#include <unistd.h>

#include <climits>
#include <cstdlib>
#include <cstring>

namespace glb {

std::string getHostname() {
  char buf[HOST_NAME_MAX + 1];
  if (gethostname(buf, sizeof(buf)) != 0) return "localhost";
  buf[HOST_NAME_MAX] = '\0';
  return std::string(buf);
}

void setThreadName(const std::string& name) {
  std::string n = name.substr(0, 15);
  pthread_setname_np(pthread_self(), n.c_str());
}

namespace {
const char* lookup(const char* name) {
  std::string a = std::string("GLB_") + name;
  if (const char* v = std::getenv(a.c_str())) return v;
  std::string b = std::string("GLOO_") + name;
  return std::getenv(b.c_str());
}
}  // namespace

bool envFlag(const char* name, bool dflt) {
  const char* v = lookup(name);
  if (v == nullptr || *v == '\0') return dflt;
  return !(std::strcmp(v, "0") == 0 || strcasecmp(v, "false") == 0 || strcasecmp(v, "off") == 0 ||
           strcasecmp(v, "no") == 0);
}

long envInt(const char* name, long dflt) {
  const char* v = lookup(name);
  if (v == nullptr || *v == '\0') return dflt;
  char* end = nullptr;
  long r = std::strtol(v, &end, 0);
  return end == v ? dflt : r;
}

std::string envStr(const char* name, const std::string& dflt) {
  const char* v = lookup(name);
  return v ? std::string(v) : dflt;
}

bool useRankAsSeqNumber() { return envFlag("ENABLE_RANK_AS_SEQUENCE_NUMBER", false); }
bool isStoreExtendedApiEnabled() { return envFlag("ENABLE_STORE_V2_API", false); }
bool disableConnectionRetries() { return envFlag("DISABLE_CONNECTION_RETRIES", false); }

}  // namespace glb
