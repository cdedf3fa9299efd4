// Benchmark harness: options, timer/samples/distribution, Benchmark interface,
// Runner (rendezvous, sweep, warm-up, iteration-count selection, table output).
// Parity: gloo/benchmark/{options.h,benchmark.h,runner.h,timer.h}.
#pragma once

#include <chrono>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "glb/context.h"
#include "glb/rendezvous/context.h"
#include "glb/transport/device.h"

namespace glb {
namespace benchmark {

struct Options {
  // participation
  int contextSize = 0;
  int contextRank = 0;
  // rendezvous
  std::string redisHost;
  int redisPort = 6379;
  std::string prefix = "benchmark";
  std::string sharedPath;
  // transport
  std::string transport = "tcp";
  std::vector<std::string> tcpDevice;
  std::string pkey, cert, caFile, caPath;  // --transport=tls
  std::string ibDevice;                     // --transport=ibverbs (probe only in this build)
  int ibIndex = 0, ibPort = 1;
  bool sync = false;
  bool busyPoll = false;
  // parameters
  std::string benchmark;
  bool verify = true;
  bool showAllErrors = false;
  int inputs = 1;
  long elements = -1;  // -1: sweep
  int warmupIterationCount = 5;
  long iterationCount = -1;
  long iterationTimeNanos = 2L * 1000 * 1000 * 1000;
  int threads = 1;
  bool showNanos = false;
  bool gpuDirect = false;
  bool halfPrecision = false;
  int destinations = 1;
  int base = 2;
  int messages = 10000;
  std::string cudaAlgo = "auto";  // kernel variant for cuda_* benchmarks
  int cudaDevice = -1;            // default: rank % device count
  bool extendedSweep = false;     // sweep 1 .. 1e8 instead of 100 .. 5e6
};

Options parseOptions(int argc, char** argv);

class Timer {
 public:
  Timer() { start(); }
  void start() { begin_ = std::chrono::high_resolution_clock::now(); }
  long ns() const {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::high_resolution_clock::now() - begin_)
        .count();
  }

 private:
  std::chrono::high_resolution_clock::time_point begin_;
};

class Distribution {
 public:
  void add(long ns) { samples_.push_back(ns); }
  void merge(const Distribution& o) { samples_.insert(samples_.end(), o.samples_.begin(), o.samples_.end()); }
  size_t size() const { return samples_.size(); }
  long min() const;
  long max() const;
  long percentile(double p) const;  // sorted[p * n], as the reference (timer.h:93-95)
  long sum() const;
  void clear() { samples_.clear(); }

 private:
  mutable std::vector<long> samples_;
  mutable bool sorted_ = false;
  void sort() const;
};

// One benchmark = closure set created for a (context, options) pair.
struct Benchmark {
  std::function<void(size_t elements)> initialize;
  std::function<void()> run;
  std::function<void()> verify;                 // may be empty
  std::function<double()> deviceNs;             // optional: device time of the last run() (CUDA events)
  size_t elementSize = 4;
  double busFactor = 0.0;  // busbw = algbw * busFactor (0: not applicable)
};

using BenchmarkFactory = std::function<Benchmark(std::shared_ptr<Context>, const Options&)>;
const std::map<std::string, BenchmarkFactory>& benchmarkRegistry();

class Runner {
 public:
  explicit Runner(const Options& options);
  ~Runner();
  void run();

 private:
  void runSize(const BenchmarkFactory& factory, size_t elements);
  long broadcastValue(long v);
  void printHeader();
  void printRow(size_t elements, size_t elementSize, const Distribution& host, const Distribution& dev,
                double busFactor);

  Options options_;
  std::shared_ptr<transport::Device> device_;
  std::shared_ptr<rendezvous::Context> backing_;
  std::unique_ptr<rendezvous::ContextFactory> factory_;
};

}  // namespace benchmark
}  // namespace glb
