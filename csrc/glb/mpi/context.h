// MPI bootstrap: build the full mesh with MPI collectives instead of a Store; MPI is
// not used after setup. Compiled only when an MPI toolchain is present (build.py
// defines GLB_USE_MPI when mpi.h is found; this image has none), otherwise the
// factory functions throw. Parity: gloo/mpi/context.{h,cc}.
#pragma once

#include <memory>

#include "glb/context.h"
#include "glb/transport/device.h"

#if GLB_USE_MPI
#include <mpi.h>
#endif

namespace glb {
namespace mpi {

#if GLB_USE_MPI
// Ref-counted MPI_Init / MPI_Finalize for contexts that own the MPI session.
class MPIScope {
 public:
  MPIScope();
  ~MPIScope();
};

class Context : public ::glb::Context {
 public:
  // Initialises MPI if needed and finalises it when the last managed context dies.
  static std::shared_ptr<Context> createManaged();

  explicit Context(const MPI_Comm& comm);
  ~Context() override;

  void connectFullMesh(std::shared_ptr<transport::Device>& dev);

 protected:
  MPI_Comm comm_;
  std::shared_ptr<MPIScope> scope_;
};
#else
// Stub so callers can probe for the feature at run time.
inline bool available() { return false; }
#endif

}  // namespace mpi
}  // namespace glb
