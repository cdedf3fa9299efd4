#include "glb/cuda/nccl_wrapper.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "glb/broadcast.h"
#include "glb/common/logging.h"
#include "glb/cuda/cuda_util.h"
#include "glb/cuda/stream.h"

namespace glb {
namespace cuda {

namespace {

// Minimal mirror of the NCCL ABI we use (stable across 2.x).
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId {
  char internal[128];
};
enum { ncclSuccess = 0 };
enum ncclDataType { nInt8 = 0, nUint8 = 1, nInt32 = 2, nUint32 = 3, nInt64 = 4, nUint64 = 5, nFloat16 = 6, nFloat32 = 7, nFloat64 = 8, nBfloat16 = 9 };
enum ncclRedOp { nSum = 0, nProd = 1, nMax = 2, nMin = 3 };

struct Api {
  int (*ncclGetVersion)(int*);
  int (*ncclGetUniqueId)(ncclUniqueId*);
  int (*ncclCommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  int (*ncclCommInitAll)(ncclComm_t*, int, const int*);
  int (*ncclCommDestroy)(ncclComm_t);
  const char* (*ncclGetErrorString)(int);
  int (*ncclAllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*ncclReduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t);
  int (*ncclReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*ncclBroadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*ncclAllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  int (*ncclSend)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*ncclRecv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*ncclGroupStart)();
  int (*ncclGroupEnd)();
  // Optional (NCCL >= 2.19): NCCL-allocated, registered user buffers — what lets NCCL run
  // NVLS / zero-copy on the caller's buffer. The comparator uses them so that "ours on
  // symmetric memory vs NCCL" compares like with like.
  int (*ncclMemAlloc)(void**, size_t);
  int (*ncclMemFree)(void*);
  int (*ncclCommRegister)(ncclComm_t, void*, size_t, void**);
  int (*ncclCommDeregister)(ncclComm_t, void*);
};

Api gApi;
bool gLoaded = false;
std::string gError;
std::once_flag gOnce;

void load() {
  void* lib = nullptr;
  for (const char* n : {"libnccl.so.2", "libnccl.so"}) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib != nullptr) break;
  }
  if (lib == nullptr) {
    gError = "libnccl.so.2 not found";
    return;
  }
  bool ok = true;
#define GLB_NCCL(name)                                                   \
  gApi.name = reinterpret_cast<decltype(gApi.name)>(dlsym(lib, #name)); \
  if (gApi.name == nullptr) {                                            \
    ok = false;                                                          \
    gError = "missing NCCL symbol " #name;                               \
  }
  GLB_NCCL(ncclGetVersion)
  GLB_NCCL(ncclGetUniqueId)
  GLB_NCCL(ncclCommInitRank)
  GLB_NCCL(ncclCommInitAll)
  GLB_NCCL(ncclCommDestroy)
  GLB_NCCL(ncclGetErrorString)
  GLB_NCCL(ncclAllReduce)
  GLB_NCCL(ncclReduce)
  GLB_NCCL(ncclReduceScatter)
  GLB_NCCL(ncclBroadcast)
  GLB_NCCL(ncclAllGather)
  GLB_NCCL(ncclSend)
  GLB_NCCL(ncclRecv)
  GLB_NCCL(ncclGroupStart)
  GLB_NCCL(ncclGroupEnd)
#undef GLB_NCCL
  gApi.ncclMemAlloc = reinterpret_cast<decltype(gApi.ncclMemAlloc)>(dlsym(lib, "ncclMemAlloc"));
  gApi.ncclMemFree = reinterpret_cast<decltype(gApi.ncclMemFree)>(dlsym(lib, "ncclMemFree"));
  gApi.ncclCommRegister = reinterpret_cast<decltype(gApi.ncclCommRegister)>(dlsym(lib, "ncclCommRegister"));
  gApi.ncclCommDeregister = reinterpret_cast<decltype(gApi.ncclCommDeregister)>(dlsym(lib, "ncclCommDeregister"));
  gLoaded = ok;
}

const Api& api() {
  std::call_once(gOnce, load);
  if (!gLoaded) GLB_THROW_INVALID_OPERATION_EXCEPTION("NCCL unavailable: ", gError);
  return gApi;
}

void check(int rc, const char* what) {
  if (rc != ncclSuccess) GLB_THROW(Exception, what, ": NCCL error: ", api().ncclGetErrorString(rc));
}

int toType(DataType dt) {
  switch (dt) {
    case DataType::INT8: return nInt8;
    case DataType::UINT8: return nUint8;
    case DataType::INT32: return nInt32;
    case DataType::UINT32: return nUint32;
    case DataType::INT64: return nInt64;
    case DataType::UINT64: return nUint64;
    case DataType::FLOAT16: return nFloat16;
    case DataType::FLOAT32: return nFloat32;
    case DataType::FLOAT64: return nFloat64;
    case DataType::BFLOAT16: return nBfloat16;
    default: break;
  }
  GLB_THROW_INVALID_OPERATION_EXCEPTION("dtype not supported by NCCL: ", dataTypeName(dt));
}

int toOp(ReduceOp op) {
  switch (op) {
    case ReduceOp::SUM: return nSum;
    case ReduceOp::PRODUCT: return nProd;
    case ReduceOp::MAX: return nMax;
    case ReduceOp::MIN: return nMin;
    default: break;
  }
  GLB_THROW_INVALID_OPERATION_EXCEPTION("reduce op not supported by NCCL");
}

}  // namespace

bool ncclAvailable() {
  std::call_once(gOnce, load);
  return gLoaded;
}

std::string ncclVersionString() {
  int v = 0;
  if (!ncclAvailable() || gApi.ncclGetVersion(&v) != ncclSuccess) return "unavailable";
  return strcat_all(v / 10000, ".", (v / 100) % 100, ".", v % 100);
}

std::shared_ptr<NcclComm> NcclComm::initRank(const std::shared_ptr<Context>& ctx, int device) {
  const auto& a = api();
  ncclUniqueId id;
  std::memset(&id, 0, sizeof(id));
  if (ctx->rank == 0) check(a.ncclGetUniqueId(&id), "ncclGetUniqueId");
  if (ctx->size > 1) {
    BroadcastOptions o(ctx);
    o.setOutputRaw(&id, sizeof(id));
    o.setRoot(0);
    o.setTag(0x7CC10000u);
    ::glb::broadcast(o);
  }
  DeviceGuard g(device);
  std::shared_ptr<NcclComm> c(new NcclComm());
  c->rank_ = ctx->rank;
  c->size_ = ctx->size;
  c->device_ = device;
  ncclComm_t comm = nullptr;
  {
    // NCCL init allocates device memory: keep it from interleaving with other allocators
    // (the deadlock described in the reference's docs/cuda.md:40-59).
    std::lock_guard<std::mutex> lk(CudaShared::getMutex());
    check(a.ncclCommInitRank(&comm, ctx->size, id, ctx->rank), "ncclCommInitRank");
  }
  c->comm_ = comm;
  return c;
}

std::vector<std::shared_ptr<NcclComm>> NcclComm::initAll(const std::vector<int>& devices) {
  const auto& a = api();
  std::vector<ncclComm_t> comms(devices.size());
  {
    std::lock_guard<std::mutex> lk(CudaShared::getMutex());
    check(a.ncclCommInitAll(comms.data(), static_cast<int>(devices.size()), devices.data()), "ncclCommInitAll");
  }
  std::vector<std::shared_ptr<NcclComm>> out;
  for (size_t i = 0; i < devices.size(); i++) {
    std::shared_ptr<NcclComm> c(new NcclComm());
    c->rank_ = static_cast<int>(i);
    c->size_ = static_cast<int>(devices.size());
    c->device_ = devices[i];
    c->comm_ = comms[i];
    out.push_back(std::move(c));
  }
  return out;
}

NcclComm::~NcclComm() {
  if (comm_ != nullptr && gLoaded) {
    DeviceGuard g(device_);
    gApi.ncclCommDestroy(static_cast<ncclComm_t>(comm_));
  }
}

void* NcclComm::memAlloc(size_t bytes) {
  const Api& a = api();
  if (a.ncclMemAlloc == nullptr) GLB_THROW_INVALID_OPERATION_EXCEPTION("this NCCL has no ncclMemAlloc");
  DeviceGuard g(device_);
  void* p = nullptr;
  check(a.ncclMemAlloc(&p, bytes), "ncclMemAlloc");
  return p;
}

void NcclComm::memFree(void* p) {
  const Api& a = api();
  if (a.ncclMemFree != nullptr && p != nullptr) {
    DeviceGuard g(device_);
    a.ncclMemFree(p);
  }
}

void* NcclComm::registerBuffer(void* p, size_t bytes) {
  const Api& a = api();
  if (a.ncclCommRegister == nullptr) GLB_THROW_INVALID_OPERATION_EXCEPTION("this NCCL has no ncclCommRegister");
  DeviceGuard g(device_);
  void* handle = nullptr;
  check(a.ncclCommRegister(static_cast<ncclComm_t>(comm_), p, bytes, &handle), "ncclCommRegister");
  return handle;
}

void NcclComm::deregisterBuffer(void* handle) {
  const Api& a = api();
  if (a.ncclCommDeregister != nullptr && handle != nullptr) a.ncclCommDeregister(static_cast<ncclComm_t>(comm_), handle);
}

void NcclComm::groupStart() { check(api().ncclGroupStart(), "ncclGroupStart"); }
void NcclComm::groupEnd() { check(api().ncclGroupEnd(), "ncclGroupEnd"); }

void NcclComm::allreduce(const void* src, void* dst, size_t count, DataType dt, ReduceOp op, cudaStream_t stream) {
  DeviceGuard g(device_);
  check(api().ncclAllReduce(src, dst, count, toType(dt), toOp(op), static_cast<ncclComm_t>(comm_), stream), "ncclAllReduce");
}

void NcclComm::reduce(const void* src, void* dst, size_t count, DataType dt, ReduceOp op, int root, cudaStream_t stream) {
  DeviceGuard g(device_);
  check(api().ncclReduce(src, dst, count, toType(dt), toOp(op), root, static_cast<ncclComm_t>(comm_), stream), "ncclReduce");
}

void NcclComm::reduceScatter(const void* src, void* dst, size_t recvCount, DataType dt, ReduceOp op, cudaStream_t stream) {
  DeviceGuard g(device_);
  check(api().ncclReduceScatter(src, dst, recvCount, toType(dt), toOp(op), static_cast<ncclComm_t>(comm_), stream),
        "ncclReduceScatter");
}

void NcclComm::broadcast(const void* src, void* dst, size_t count, DataType dt, int root, cudaStream_t stream) {
  DeviceGuard g(device_);
  check(api().ncclBroadcast(src, dst, count, toType(dt), root, static_cast<ncclComm_t>(comm_), stream), "ncclBroadcast");
}

void NcclComm::allgather(const void* src, void* dst, size_t sendCount, DataType dt, cudaStream_t stream) {
  DeviceGuard g(device_);
  check(api().ncclAllGather(src, dst, sendCount, toType(dt), static_cast<ncclComm_t>(comm_), stream), "ncclAllGather");
}

void NcclComm::alltoall(const void* src, void* dst, size_t countPerRank, DataType dt, cudaStream_t stream) {
  DeviceGuard g(device_);
  const auto& a = api();
  const size_t es = elementSize(dt);
  check(a.ncclGroupStart(), "ncclGroupStart");
  for (int r = 0; r < size_; r++) {
    check(a.ncclSend(static_cast<const char*>(src) + r * countPerRank * es, countPerRank, toType(dt), r,
                     static_cast<ncclComm_t>(comm_), stream), "ncclSend");
    check(a.ncclRecv(static_cast<char*>(dst) + r * countPerRank * es, countPerRank, toType(dt), r,
                     static_cast<ncclComm_t>(comm_), stream), "ncclRecv");
  }
  check(a.ncclGroupEnd(), "ncclGroupEnd");
}

void NcclComm::sendrecv(const void* src, int dst, void* dstBuf, int srcRank, size_t count, DataType dt,
                        cudaStream_t stream) {
  DeviceGuard g(device_);
  const auto& a = api();
  check(a.ncclGroupStart(), "ncclGroupStart");
  check(a.ncclSend(src, count, toType(dt), dst, static_cast<ncclComm_t>(comm_), stream), "ncclSend");
  check(a.ncclRecv(dstBuf, count, toType(dt), srcRank, static_cast<ncclComm_t>(comm_), stream), "ncclRecv");
  check(a.ncclGroupEnd(), "ncclGroupEnd");
}

}  // namespace cuda
}  // namespace glb
