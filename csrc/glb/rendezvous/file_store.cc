#include "glb/rendezvous/file_store.h"

#include <fcntl.h>
#include <sys/file.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cstring>
#include <fstream>
#include <thread>

#include "glb/common/logging.h"
#include "glb/common/utils.h"

namespace glb {
namespace rendezvous {

namespace {
constexpr auto kPollInterval = std::chrono::milliseconds(2);

void mkdirs(const std::string& path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); i++) {
    if (i == path.size() || path[i] == '/') {
      if (!cur.empty() && ::mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) {
        GLB_THROW_IO_EXCEPTION("mkdir ", cur, ": ", std::strerror(errno));
      }
    }
    if (i < path.size()) cur.push_back(path[i]);
  }
}
}  // namespace

FileStore::FileStore(const std::string& path) : base_(path) {
  GLB_ENFORCE(!path.empty(), "FileStore needs a directory");
  mkdirs(base_);
  struct stat st;
  GLB_ENFORCE(::stat(base_.c_str(), &st) == 0 && S_ISDIR(st.st_mode), "Not a directory: ", base_);
}

std::string FileStore::keyFileName(const std::string& key) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned char c : key) {
    h ^= c;
    h *= 1099511628211ull;
  }
  char hex[17];
  std::snprintf(hex, sizeof(hex), "%016llx", static_cast<unsigned long long>(h));
  std::string tail;
  for (size_t i = key.size() > 40 ? key.size() - 40 : 0; i < key.size(); i++) {
    char c = key[i];
    tail.push_back((std::isalnum(static_cast<unsigned char>(c)) || c == '-' || c == '.') ? c : '_');
  }
  return std::string(hex) + "_" + tail;
}

std::string FileStore::objectPath(const std::string& key) const { return base_ + "/" + keyFileName(key); }

bool FileStore::exists(const std::string& path) const { return ::access(path.c_str(), F_OK) == 0; }

void FileStore::writeAtomic(const std::string& path, const Bytes& data, bool exclusive) {
  static std::atomic<uint64_t> counter{0};
  std::string tmp = strcat_all(path, ".tmp.", ::getpid(), ".", counter.fetch_add(1));
  int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL, 0644);
  if (fd < 0) GLB_THROW_IO_EXCEPTION("open ", tmp, ": ", std::strerror(errno));
  size_t off = 0;
  while (off < data.size()) {
    ssize_t n = ::write(fd, data.data() + off, data.size() - off);
    if (n < 0) {
      if (errno == EINTR) continue;
      ::close(fd);
      ::unlink(tmp.c_str());
      GLB_THROW_IO_EXCEPTION("write ", tmp, ": ", std::strerror(errno));
    }
    off += static_cast<size_t>(n);
  }
  ::close(fd);
  if (exclusive) {
    // link() fails with EEXIST if the key was already published: write-once semantics.
    int rv = ::link(tmp.c_str(), path.c_str());
    int err = errno;
    ::unlink(tmp.c_str());
    if (rv != 0) {
      if (err == EEXIST) {
        throw EnforceNotMet(__FILE__, __LINE__, "key not set",
                            strcat_all("Key already set (write-once): ", path));
      }
      GLB_THROW_IO_EXCEPTION("link ", path, ": ", std::strerror(err));
    }
  } else if (::rename(tmp.c_str(), path.c_str()) != 0) {
    int err = errno;
    ::unlink(tmp.c_str());
    GLB_THROW_IO_EXCEPTION("rename ", path, ": ", std::strerror(err));
  }
  std::lock_guard<std::mutex> g(mu_);
  created_.push_back(path);
}

void FileStore::set(const std::string& key, const Bytes& data) {
  writeAtomic(objectPath(key), data, /*exclusive=*/true);
}

IStore::Bytes FileStore::get(const std::string& key) {
  auto path = objectPath(key);
  wait({key}, kDefaultTimeout);
  std::ifstream in(path, std::ios::binary);
  GLB_ENFORCE(in.good(), "Cannot open ", path);
  return Bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}

void FileStore::wait(const std::vector<std::string>& keys, std::chrono::milliseconds timeout) {
  const auto start = std::chrono::steady_clock::now();
  std::vector<std::string> pending;
  for (const auto& k : keys) pending.push_back(objectPath(k));
  auto sleep = std::chrono::microseconds(200);
  while (true) {
    while (!pending.empty() && exists(pending.back())) pending.pop_back();
    if (pending.empty()) return;
    if (timeout != kNoTimeout && std::chrono::steady_clock::now() - start > timeout) {
      GLB_THROW_IO_EXCEPTION("Wait timeout for key(s): ", keys.size() == 1 ? keys[0] : "[multiple]",
                             " in ", base_);
    }
    std::this_thread::sleep_for(sleep);
    if (sleep < kPollInterval) sleep *= 2;
  }
}

std::vector<IStore::Bytes> FileStore::multi_get(const std::vector<std::string>& keys) {
  wait(keys, kDefaultTimeout);
  std::vector<Bytes> out;
  for (const auto& k : keys) out.push_back(get(k));
  return out;
}

void FileStore::multi_set(const std::vector<std::string>& keys, const std::vector<Bytes>& values) {
  GLB_ENFORCE_EQ(keys.size(), values.size());
  for (size_t i = 0; i < keys.size(); i++) set(keys[i], values[i]);
}

namespace {
struct FileLock {
  int fd;
  explicit FileLock(const std::string& path) {
    fd = ::open(path.c_str(), O_RDWR | O_CREAT, 0644);
    if (fd < 0) GLB_THROW_IO_EXCEPTION("open ", path, ": ", std::strerror(errno));
    while (::flock(fd, LOCK_EX) != 0) {
      if (errno != EINTR) {
        ::close(fd);
        GLB_THROW_IO_EXCEPTION("flock ", path, ": ", std::strerror(errno));
      }
    }
  }
  ~FileLock() {
    ::flock(fd, LOCK_UN);
    ::close(fd);
  }
};

IStore::Bytes slurp(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in.good()) return {};
  return IStore::Bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}
}  // namespace

void FileStore::append(const std::string& key, const Bytes& data) {
  auto path = objectPath(key);
  FileLock lock(path + ".lock");
  Bytes cur = slurp(path);
  cur.insert(cur.end(), data.begin(), data.end());
  writeAtomic(path, cur, /*exclusive=*/false);
}

int64_t FileStore::add(const std::string& key, int64_t value) {
  auto path = objectPath(key);
  FileLock lock(path + ".lock");
  Bytes cur = slurp(path);
  int64_t v = 0;
  if (cur.size() == sizeof(v)) std::memcpy(&v, cur.data(), sizeof(v));
  v += value;
  cur.resize(sizeof(v));
  std::memcpy(cur.data(), &v, sizeof(v));
  writeAtomic(path, cur, /*exclusive=*/false);
  return v;
}

std::vector<std::string> FileStore::getAllKeyFilePaths() const {
  std::lock_guard<std::mutex> g(mu_);
  return created_;
}

}  // namespace rendezvous
}  // namespace glb
