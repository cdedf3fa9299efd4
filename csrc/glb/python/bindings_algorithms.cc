// Old-style Algorithm classes exposed to Python (filled in with algorithm.h).
#include <pybind11/pybind11.h>
namespace py = pybind11;
namespace glb_py {
void registerOldStyle(py::module_& m) {}
}  // namespace glb_py
