// Single-GPU self-test of the collective kernels (see selftest.cc).
#pragma once

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "glb/cuda/collectives.h"

namespace glb {
namespace cuda {

struct SelfTestResult {
  std::string name;    // kernel (instantiation) exercised
  bool ok = false;
  bool skipped = false;
  std::string detail;  // first mismatch / reason for skipping
};

// Runs every hot kernel with 2 / 4 / 8 virtual ranks on this GPU and checks rank 0's part
// of the results. `pc` may be a context of any size (only local memory is touched); no
// peer takes part, so this is safe under Nsight Compute.
std::vector<SelfTestResult> loopbackSelfTest(PeerContext& pc, cudaStream_t stream, size_t count = 1 << 18);

// Launches a barrier whose peer never arrives, with the given device timeout. Returns true
// when the kernel gave up, the stream drained and checkHealth() raised IoException.
// Poisons `pc` (use a throw-away context).
bool loopbackTimeoutTest(PeerContext& pc, cudaStream_t stream, int timeoutMs, double* elapsedMs = nullptr);

// The LocalOp classes (memcpy, native / host / NCCL reduce and broadcast, dispatchers) over
// buffers on the given devices (one entry = everything on that device).
std::vector<SelfTestResult> localOpsSelfTest(const std::vector<int>& devices, size_t count = 100003);

}  // namespace cuda
}  // namespace glb
