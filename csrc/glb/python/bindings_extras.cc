// Diagnostics that do not belong to a class binding.
#include <pybind11/pybind11.h>

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
}
}  // namespace glb_py
