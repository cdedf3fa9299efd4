// glb_benchmark — the benchmark CLI (host + CUDA collectives in one binary).
//
//   glb_benchmark --size P --rank r --shared-path DIR [options] BENCHMARK
//
// Same flags, sweep, iteration-count logic and output table as the reference's
// `benchmark` / `benchmark_cuda` (gloo/benchmark/{options,runner,main,cuda_main}.cc),
// plus device-timed columns for the CUDA benchmarks (CUDA events, max over ranks)
// and a bus-bandwidth column. One process per rank (per GPU).
#include <getopt.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <csignal>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "glb/benchmark/harness.h"

using namespace glb;
using namespace glb::benchmark;

int main(int argc, char** argv) {
  std::signal(SIGPIPE, SIG_IGN);  // a vanished peer is an IoException, not a signal (reference: test/main.cc:38-41)
  try {
    Options opts = parseOptions(argc, argv);
    Runner runner(opts);
    runner.run();
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "glb_benchmark: " << e.what() << std::endl;
    return 1;
  }
}
