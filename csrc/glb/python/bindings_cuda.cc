#include <pybind11/pybind11.h>
namespace py = pybind11;
namespace glb_py {
void registerCuda(py::module_& m) {}
}  // namespace glb_py
