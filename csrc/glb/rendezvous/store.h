// rendezvous::Store — the abstract KV store handed to connectFullMesh().
// Parity: gloo/rendezvous/store.h:25-74.
#pragma once

#include "glb/common/store.h"

namespace glb {
namespace rendezvous {

using Store = ::glb::IStore;

}  // namespace rendezvous
}  // namespace glb
