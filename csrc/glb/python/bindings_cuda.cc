// pybind11 bindings for the CUDA side: PeerContext (topology + symmetric memory),
// peer buffers, fused collectives, tuning knobs, test helpers.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "glb/cuda/collectives.h"
#include "glb/cuda/cuda_util.h"
#include "glb/cuda/kernels.h"
#include "glb/cuda/peer_context.h"
#include "glb/cuda/stream.h"

namespace py = pybind11;
using namespace glb;
using namespace glb::cuda;

namespace glb_py {

void registerCudaAlgorithms(py::module_& m);  // bindings_cuda_algorithms.cc

namespace {
inline void* P(uintptr_t p) { return reinterpret_cast<void*>(p); }
inline cudaStream_t S(uintptr_t s) { return reinterpret_cast<cudaStream_t>(s); }
}  // namespace

void registerCuda(py::module_& root) {
  auto m = root.def_submodule("cuda", "NVLink peer-memory collectives");

  m.def("device_count", &deviceCount);
  m.def("device_pci_bus_id", &devicePCIBusId);
  m.attr("MAX_RANKS") = kMaxRanks;
  m.attr("THREADS") = kThreads;

  py::class_<DeviceInfo>(m, "DeviceInfo")
      .def_property_readonly("hostname", [](const DeviceInfo& d) { return std::string(d.hostname); })
      .def_readonly("pid", &DeviceInfo::pid)
      .def_readonly("device", &DeviceInfo::device)
      .def_property_readonly("uuid", [](const DeviceInfo& d) {
        return py::bytes(reinterpret_cast<const char*>(d.uuid), 16);
      })
      .def_property_readonly("pci_bus_id", [](const DeviceInfo& d) { return std::string(d.pciBusId); })
      .def_readonly("sm_count", &DeviceInfo::smCount)
      .def_readonly("cc_major", &DeviceInfo::ccMajor)
      .def_readonly("cc_minor", &DeviceInfo::ccMinor)
      .def_readonly("vmm_supported", &DeviceInfo::vmmSupported)
      .def_readonly("multicast_supported", &DeviceInfo::multicastSupported)
      .def_readonly("total_mem", &DeviceInfo::totalMem);

  py::class_<PeerBuffer, std::shared_ptr<PeerBuffer>>(m, "PeerBuffer")
      .def_property_readonly("ptr", [](const PeerBuffer& b) { return reinterpret_cast<uintptr_t>(b.local); })
      .def_readonly("nbytes", &PeerBuffer::bytes)
      .def_readonly("vector_ok", &PeerBuffer::vectorOk)
      .def_property_readonly("has_multicast", [](const PeerBuffer& b) { return b.mc != nullptr; })
      .def_property_readonly("multicast_ptr", [](const PeerBuffer& b) { return reinterpret_cast<uintptr_t>(b.mc); })
      .def("peer_ptr", [](const PeerBuffer& b, int r) { return reinterpret_cast<uintptr_t>(b.peer[r]); });

  py::class_<PeerContext, std::shared_ptr<PeerContext>>(m, "PeerContext")
      .def(py::init([](std::shared_ptr<Context> ctx, int device, size_t stageBytes, bool useVmm, bool useNvls) {
        PeerOptions o;
        if (stageBytes > 0) o.stageBytes = stageBytes;
        o.useVmm = useVmm;
        o.useNvls = useNvls;
        py::gil_scoped_release nogil;
        return std::make_shared<PeerContext>(std::move(ctx), device, o);
      }), py::arg("ctx"), py::arg("device"), py::arg("stage_bytes") = 0, py::arg("use_vmm") = true,
          py::arg("use_nvls") = true)
      .def_readonly("rank", &PeerContext::rank)
      .def_readonly("size", &PeerContext::size)
      .def_readonly("device", &PeerContext::device)
      .def("topology", &PeerContext::topology)
      .def("peer_access_everywhere", &PeerContext::peerAccessEverywhere)
      .def("using_vmm", &PeerContext::usingVmm)
      .def("nvls_available", &PeerContext::nvlsAvailable)
      .def("ranks_on_my_device", &PeerContext::ranksOnMyDevice)
      .def("max_blocks", &PeerContext::maxBlocks)
      .def("stage_bytes", &PeerContext::stageBytes)
      .def("describe", &PeerContext::describe)
      .def("alloc_symmetric", [](PeerContext& pc, size_t bytes) {
        py::gil_scoped_release nogil;
        return pc.allocSymmetric(bytes);
      })
      .def("register_buffer", [](PeerContext& pc, uintptr_t ptr, size_t bytes) {
        py::gil_scoped_release nogil;
        return pc.registerBuffer(P(ptr), bytes);
      })
      .def("host_barrier", [](PeerContext& pc) {
        py::gil_scoped_release nogil;
        pc.hostBarrier();
      });

  m.def("barrier", [](PeerContext& pc, uintptr_t stream) {
    py::gil_scoped_release nogil;
    barrier(pc, S(stream));
  }, py::arg("pc"), py::arg("stream") = 0);

  m.def("allreduce_registered", [](PeerContext& pc, const PeerBuffer& buf, size_t byteOffset, size_t count, int dtype,
                                   int op, int algo, uintptr_t stream) {
    py::gil_scoped_release nogil;
    allreduce(pc, buf, byteOffset, count, static_cast<DataType>(dtype), static_cast<ReduceOp>(op),
              static_cast<AllreduceAlgo>(algo), S(stream));
  }, py::arg("pc"), py::arg("buf"), py::arg("byte_offset"), py::arg("count"), py::arg("dtype"), py::arg("op") = 1,
     py::arg("algo") = 0, py::arg("stream") = 0);

  m.def("allreduce", [](PeerContext& pc, uintptr_t in, uintptr_t out, size_t count, int dtype, int op, int algo,
                        uintptr_t stream) {
    py::gil_scoped_release nogil;
    allreduce(pc, P(in), P(out), count, static_cast<DataType>(dtype), static_cast<ReduceOp>(op),
              static_cast<AllreduceAlgo>(algo), S(stream));
  }, py::arg("pc"), py::arg("input"), py::arg("output"), py::arg("count"), py::arg("dtype"), py::arg("op") = 1,
     py::arg("algo") = 0, py::arg("stream") = 0);

  m.def("choose_allreduce", [](const PeerContext& pc, size_t bytes, int dtype, int op, bool registered, bool mc) {
    return static_cast<int>(chooseAllreduce(pc, bytes, static_cast<DataType>(dtype), static_cast<ReduceOp>(op),
                                            registered, mc));
  });
  m.def("allreduce_algo_name", [](int a) { return std::string(allreduceAlgoName(static_cast<AllreduceAlgo>(a))); });

  // Dedicated (never pooled / shared) streams: torch.cuda.Stream() hands out streams from
  // a 32-entry pool, so two ranks living in one process can end up on the same stream —
  // fatal for kernels that wait for each other on the device.
  m.def("create_stream", [](int device, bool highPriority) {
    DeviceGuard g(device);
    int lo = 0, hi = 0;
    GLB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s = nullptr;
    GLB_CUDA_CHECK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, highPriority ? hi : lo));
    return reinterpret_cast<uintptr_t>(s);
  }, py::arg("device"), py::arg("high_priority") = false);
  m.def("destroy_stream", [](uintptr_t s) { cudaStreamDestroy(S(s)); });
  m.def("launch_count", &launchCount);
  m.def("get_tuning", [] {
    const auto& t = tuning();
    py::dict d;
    d["one_shot_max_bytes"] = t.oneShotMaxBytes;
    d["nvls_min_bytes"] = t.nvlsMinBytes;
    d["max_blocks"] = t.maxBlocks;
    d["one_shot_blocks"] = t.oneShotBlocks;
    return d;
  });
  m.def("set_tuning", [](py::dict d) {
    auto& t = tuning();
    if (d.contains("one_shot_max_bytes")) t.oneShotMaxBytes = d["one_shot_max_bytes"].cast<size_t>();
    if (d.contains("nvls_min_bytes")) t.nvlsMinBytes = d["nvls_min_bytes"].cast<size_t>();
    if (d.contains("max_blocks")) t.maxBlocks = d["max_blocks"].cast<int>();
    if (d.contains("one_shot_blocks")) t.oneShotBlocks = d["one_shot_blocks"].cast<int>();
    if (d.contains("nvls_reduce_scatter")) t.nvlsReduceScatter = d["nvls_reduce_scatter"].cast<bool>();
    if (d.contains("copy_blocks")) t.copyBlocks = d["copy_blocks"].cast<int>();
    if (d.contains("one_shot_push")) setOneShotPush(d["one_shot_push"].cast<bool>());
  });

  // ---- local ops / helpers ------------------------------------------------------------
  m.def("local_reduce", [](uintptr_t dst, uintptr_t src, size_t count, int dtype, int op, uintptr_t stream) {
    launchLocalReduce(P(dst), P(src), count, static_cast<DataType>(dtype), static_cast<ReduceOp>(op), S(stream));
    GLB_CUDA_CHECK(cudaGetLastError());
  }, py::arg("dst"), py::arg("src"), py::arg("count"), py::arg("dtype"), py::arg("op") = 1, py::arg("stream") = 0);
  m.def("local_reduce_many", [](uintptr_t dst, std::vector<uintptr_t> srcs, size_t count, int dtype, int op,
                                uintptr_t stream) {
    std::vector<const void*> ps;
    for (auto s : srcs) ps.push_back(P(s));
    launchLocalReduceMany(P(dst), ps.data(), static_cast<int>(ps.size()), count, static_cast<DataType>(dtype),
                          static_cast<ReduceOp>(op), S(stream));
    GLB_CUDA_CHECK(cudaGetLastError());
  }, py::arg("dst"), py::arg("srcs"), py::arg("count"), py::arg("dtype"), py::arg("op") = 1, py::arg("stream") = 0);
  m.def("local_broadcast", [](std::vector<uintptr_t> dsts, uintptr_t src, size_t bytes, uintptr_t stream) {
    std::vector<void*> ps;
    for (auto d : dsts) ps.push_back(P(d));
    launchLocalBroadcast(ps.data(), static_cast<int>(ps.size()), P(src), bytes, S(stream));
    GLB_CUDA_CHECK(cudaGetLastError());
  }, py::arg("dsts"), py::arg("src"), py::arg("bytes"), py::arg("stream") = 0);
  m.def("fill", [](uintptr_t dst, size_t count, int dtype, double start, double stride, uintptr_t stream) {
    launchFill(P(dst), count, static_cast<DataType>(dtype), start, stride, S(stream));
    GLB_CUDA_CHECK(cudaGetLastError());
  }, py::arg("dst"), py::arg("count"), py::arg("dtype"), py::arg("start") = 0.0, py::arg("stride") = 1.0,
     py::arg("stream") = 0);
  m.def("spin", [](long long cycles, uintptr_t stream) {
    launchSpin(cycles, S(stream));
    GLB_CUDA_CHECK(cudaGetLastError());
  }, py::arg("cycles"), py::arg("stream") = 0);

  registerCudaAlgorithms(m);
}

}  // namespace glb_py
