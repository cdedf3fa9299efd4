// Diagnostics that do not belong to a class binding.
#include <pybind11/pybind11.h>

#include "glb/common/trace.h"
#include "glb/config.h"
#include "glb/transport/tcp/pair.h"

namespace py = pybind11;
namespace glb_py {
void registerExtras(py::module_& m) {
  m.def(
      "tcp_stats",
      [] {
        py::dict d;
        d["cma_messages"] = ::glb::transport::tcp::Pair::cmaMessages();
        d["cma_bytes"] = ::glb::transport::tcp::Pair::cmaBytes();
        d["spin_us"] = ::glb::transport::tcp::Pair::spinBudgetNanos() / 1000;
        return d;
      },
      "Process-wide tcp transport counters: payloads pulled through the same-host single-copy "
      "path (process_vm_readv) and the configured spin budget.");
  m.def(
      "build_config",
      [] {
        py::dict d;
        d["version"] = GLB_VERSION_STRING;
        d["cuda"] = static_cast<bool>(GLB_USE_CUDA);
        d["cuda_arch"] = GLB_CUDA_ARCH;
        d["mpi"] = static_cast<bool>(GLB_USE_MPI);
        d["transport_tcp"] = static_cast<bool>(GLB_HAVE_TRANSPORT_TCP);
        d["transport_tls"] = static_cast<bool>(GLB_HAVE_TRANSPORT_TCP_TLS);
        d["transport_uv"] = static_cast<bool>(GLB_HAVE_TRANSPORT_UV);
        d["transport_ibverbs"] = static_cast<bool>(GLB_HAVE_TRANSPORT_IBVERBS);
        d["transport_nvlink"] = static_cast<bool>(GLB_HAVE_TRANSPORT_NVLINK);
        return d;
      },
      "Compile-time configuration (glb/config.h).");
  m.def("trace_enabled", [] { return ::glb::trace::enabled(); }, "True when GLB_TRACE_FILE is set (chrome-trace event sink).");
  m.def("trace_flush", [] { return ::glb::trace::flush(); }, "Write the buffered trace events now; returns how many.");
  m.attr("__version__") = GLB_VERSION_STRING;
}
}  // namespace glb_py
