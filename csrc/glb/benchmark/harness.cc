#include "glb/benchmark/harness.h"

#include <getopt.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <numeric>
#include <thread>

#include "glb/barrier.h"
#include "glb/broadcast.h"
#include "glb/common/logging.h"
#include "glb/rendezvous/file_store.h"
#include "glb/rendezvous/prefix_store.h"
#include "glb/rendezvous/redis_store.h"
#include "glb/transport/ibverbs/device.h"
#include "glb/transport/tcp/device.h"
#include "glb/transport/tcp/tls/device.h"
#include "glb/transport/uv/device.h"

namespace glb {
namespace benchmark {

// ---- options ---------------------------------------------------------------------------------

static void usage(const char* argv0) {
  std::fprintf(stderr,
               "Usage: %s [OPTIONS] BENCHMARK\n\n"
               "Participation:\n"
               "  -s, --size=SIZE        Number of processes\n"
               "  -r, --rank=RANK        Rank of this process\n\n"
               "Rendezvous:\n"
               "  -h, --redis-host=HOST  Host name of Redis server\n"
               "  -p, --redis-port=PORT  Port number of Redis server\n"
               "  -x, --prefix=PREFIX    Rendezvous prefix (unique for this run)\n"
               "      --shared-path=PATH File system rendezvous with this shared path\n\n"
               "Transport:\n"
               "  -t, --transport=TRANSPORT Transport to use (tcp, tls, uv; ibverbs is probed and reported)\n"
               "      --pkey=FILE --cert=FILE --ca-file=FILE --ca-path=DIR   Credentials for --transport=tls\n"
               "      --ib-device=NAME --ib-index=N --ib-port=N              Accepted for --transport=ibverbs\n"
               "      --tcp-device=DEV[,DEV...]  Network interface(s) or address to use\n"
               "      --sync=BOOL           Switch pairs to sync mode (default: false)\n"
               "      --busy-poll=BOOL      Busy-poll in sync mode (default: false)\n\n"
               "Benchmark parameters:\n"
               "      --no-verify        Do not verify results of first iteration\n"
               "      --show-all-errors  Display all verification errors\n"
               "      --inputs           Number of input buffers\n"
               "      --elements         Number of elements per input buffer (default: sweep)\n"
               "      --extended-sweep   Sweep 1 .. 1e8 elements instead of 100 .. 5e6\n"
               "      --warmup-iters     Number of warmup iterations (default: 5)\n"
               "      --iteration-count  Number of iterations (default: derived from --iteration-time)\n"
               "      --iteration-time   Minimum time to run each size for (default: 2s)\n"
               "      --threads          Number of threads (contexts) per process (default: 1)\n"
               "      --nanos            Display timing in nanoseconds\n"
               "      --gpudirect        Accepted for compatibility (the peer path never stages)\n"
               "      --halfprecision    Use 16-bit floating point values\n"
               "      --destinations     Destinations per rank in pairwise_exchange\n"
               "      --base             Base for bcube algorithms (default: 2)\n"
               "      --messages         Messages per iteration in sendrecv stress benchmarks\n"
               "      --cuda-algo        Kernel variant for cuda_* benchmarks: auto|one_shot|two_shot|nvls\n"
               "      --cuda-device      CUDA device for this rank (default: rank %% device count)\n\n"
               "BENCHMARK is one of:\n",
               argv0);
  for (const auto& kv : benchmarkRegistry()) std::fprintf(stderr, "  %s\n", kv.first.c_str());
}

static long parseTime(const std::string& s) {
  char* end = nullptr;
  double v = std::strtod(s.c_str(), &end);
  std::string unit(end);
  if (unit == "s" || unit.empty()) return static_cast<long>(v * 1e9);
  if (unit == "ms") return static_cast<long>(v * 1e6);
  if (unit == "us") return static_cast<long>(v * 1e3);
  if (unit == "ns") return static_cast<long>(v);
  GLB_THROW(Exception, "bad time value: ", s);
}

static bool parseBool(const char* s) {
  if (s == nullptr) return true;
  std::string v(s);
  return !(v == "0" || v == "false" || v == "no" || v == "off");
}

Options parseOptions(int argc, char** argv) {
  Options o;
  enum {
    OPT_SHARED = 1000, OPT_SYNC, OPT_BUSY, OPT_TCPDEV, OPT_NOVERIFY, OPT_SHOWERR, OPT_INPUTS, OPT_ELEMENTS,
    OPT_WARMUP, OPT_ITERCOUNT, OPT_ITERTIME, OPT_THREADS, OPT_NANOS, OPT_GPUDIRECT, OPT_HALF, OPT_DEST, OPT_BASE,
    OPT_MESSAGES, OPT_CUDAALGO, OPT_CUDADEV, OPT_EXTSWEEP, OPT_PKEY, OPT_CERT, OPT_CAFILE, OPT_CAPATH,
    OPT_IBDEV, OPT_IBINDEX, OPT_IBPORT, OPT_HELP
  };
  static struct option longopts[] = {
      {"size", required_argument, nullptr, 's'},       {"rank", required_argument, nullptr, 'r'},
      {"redis-host", required_argument, nullptr, 'h'}, {"redis-port", required_argument, nullptr, 'p'},
      {"prefix", required_argument, nullptr, 'x'},     {"shared-path", required_argument, nullptr, OPT_SHARED},
      {"transport", required_argument, nullptr, 't'},  {"sync", optional_argument, nullptr, OPT_SYNC},
      {"busy-poll", optional_argument, nullptr, OPT_BUSY}, {"tcp-device", required_argument, nullptr, OPT_TCPDEV},
      {"no-verify", no_argument, nullptr, OPT_NOVERIFY},   {"show-all-errors", no_argument, nullptr, OPT_SHOWERR},
      {"inputs", required_argument, nullptr, OPT_INPUTS},  {"elements", required_argument, nullptr, OPT_ELEMENTS},
      {"warmup-iters", required_argument, nullptr, OPT_WARMUP},
      {"iteration-count", required_argument, nullptr, OPT_ITERCOUNT},
      {"iteration-time", required_argument, nullptr, OPT_ITERTIME},
      {"threads", required_argument, nullptr, OPT_THREADS}, {"nanos", no_argument, nullptr, OPT_NANOS},
      {"gpudirect", no_argument, nullptr, OPT_GPUDIRECT},   {"halfprecision", no_argument, nullptr, OPT_HALF},
      {"destinations", required_argument, nullptr, OPT_DEST}, {"base", required_argument, nullptr, OPT_BASE},
      {"messages", required_argument, nullptr, OPT_MESSAGES}, {"cuda-algo", required_argument, nullptr, OPT_CUDAALGO},
      {"cuda-device", required_argument, nullptr, OPT_CUDADEV}, {"extended-sweep", no_argument, nullptr, OPT_EXTSWEEP},
      {"pkey", required_argument, nullptr, OPT_PKEY},       {"cert", required_argument, nullptr, OPT_CERT},
      {"ca-file", required_argument, nullptr, OPT_CAFILE},  {"ca-path", required_argument, nullptr, OPT_CAPATH},
      {"ib-device", required_argument, nullptr, OPT_IBDEV}, {"ib-index", required_argument, nullptr, OPT_IBINDEX},
      {"ib-port", required_argument, nullptr, OPT_IBPORT},
      {"help", no_argument, nullptr, OPT_HELP},         {nullptr, 0, nullptr, 0}};
  int c;
  while ((c = getopt_long(argc, argv, "s:r:h:p:x:t:", longopts, nullptr)) != -1) {
    switch (c) {
      case 's': o.contextSize = std::atoi(optarg); break;
      case 'r': o.contextRank = std::atoi(optarg); break;
      case 'h': o.redisHost = optarg; break;
      case 'p': o.redisPort = std::atoi(optarg); break;
      case 'x': o.prefix = optarg; break;
      case 't': o.transport = optarg; break;
      case OPT_PKEY: o.pkey = optarg; break;
      case OPT_CERT: o.cert = optarg; break;
      case OPT_CAFILE: o.caFile = optarg; break;
      case OPT_CAPATH: o.caPath = optarg; break;
      case OPT_IBDEV: o.ibDevice = optarg; break;
      case OPT_IBINDEX: o.ibIndex = std::atoi(optarg); break;
      case OPT_IBPORT: o.ibPort = std::atoi(optarg); break;
      case OPT_SHARED: o.sharedPath = optarg; break;
      case OPT_SYNC: o.sync = parseBool(optarg); break;
      case OPT_BUSY: o.busyPoll = parseBool(optarg); break;
      case OPT_TCPDEV: {
        std::string s(optarg);
        size_t pos = 0;
        while (true) {
          auto comma = s.find(',', pos);
          o.tcpDevice.push_back(s.substr(pos, comma == std::string::npos ? comma : comma - pos));
          if (comma == std::string::npos) break;
          pos = comma + 1;
        }
        break;
      }
      case OPT_NOVERIFY: o.verify = false; break;
      case OPT_SHOWERR: o.showAllErrors = true; break;
      case OPT_INPUTS: o.inputs = std::atoi(optarg); break;
      case OPT_ELEMENTS: o.elements = std::atol(optarg); break;
      case OPT_WARMUP: o.warmupIterationCount = std::atoi(optarg); break;
      case OPT_ITERCOUNT: o.iterationCount = std::atol(optarg); break;
      case OPT_ITERTIME: o.iterationTimeNanos = parseTime(optarg); break;
      case OPT_THREADS: o.threads = std::atoi(optarg); break;
      case OPT_NANOS: o.showNanos = true; break;
      case OPT_GPUDIRECT: o.gpuDirect = true; break;
      case OPT_HALF: o.halfPrecision = true; break;
      case OPT_DEST: o.destinations = std::atoi(optarg); break;
      case OPT_BASE: o.base = std::atoi(optarg); break;
      case OPT_MESSAGES: o.messages = std::atoi(optarg); break;
      case OPT_CUDAALGO: o.cudaAlgo = optarg; break;
      case OPT_CUDADEV: o.cudaDevice = std::atoi(optarg); break;
      case OPT_EXTSWEEP: o.extendedSweep = true; break;
      case OPT_HELP:
      default:
        usage(argv[0]);
        std::exit(c == OPT_HELP ? 0 : 1);
    }
  }
  if (optind != argc - 1) {
    usage(argv[0]);
    std::exit(1);
  }
  o.benchmark = argv[optind];
  if (o.contextSize <= 0) GLB_THROW(Exception, "--size is required");
  if (o.contextRank < 0 || o.contextRank >= o.contextSize) GLB_THROW(Exception, "--rank out of range");
  if (o.sharedPath.empty() && o.redisHost.empty()) GLB_THROW(Exception, "need --shared-path or --redis-host");
  if (o.transport != "tcp" && o.transport != "tls" && o.transport != "uv" && o.transport != "ibverbs") {
    GLB_THROW(Exception, "unknown transport '", o.transport, "' (tcp, tls, uv, ibverbs)");
  }
  if (o.transport == "tls" && (o.pkey.empty() || o.cert.empty() || (o.caFile.empty() && o.caPath.empty()))) {
    GLB_THROW(Exception, "--transport=tls needs --pkey, --cert and --ca-file or --ca-path");
  }
  return o;
}

// ---- distribution -----------------------------------------------------------------------------

void Distribution::sort() const {
  if (!sorted_) {
    std::sort(samples_.begin(), samples_.end());
    sorted_ = true;
  }
}
long Distribution::min() const { sort(); return samples_.empty() ? 0 : samples_.front(); }
long Distribution::max() const { sort(); return samples_.empty() ? 0 : samples_.back(); }
long Distribution::percentile(double p) const {
  sort();
  if (samples_.empty()) return 0;
  size_t i = static_cast<size_t>(p * samples_.size());
  return samples_[std::min(i, samples_.size() - 1)];
}
long Distribution::sum() const { return std::accumulate(samples_.begin(), samples_.end(), 0L); }

// ---- runner -----------------------------------------------------------------------------------

Runner::Runner(const Options& options) : options_(options) {
  transport::tcp::attr attr;
  if (!options_.tcpDevice.empty()) {
    const std::string& d = options_.tcpDevice[options_.contextRank % options_.tcpDevice.size()];
    // An interface name ("lo", "eth0") or an address / host name.
    if (d.find('.') == std::string::npos && d.find(':') == std::string::npos) {
      attr.iface = d;
    } else {
      attr.hostname = d;
    }
  }
  if (options_.transport == "tls") {
    device_ = transport::tcp::tls::CreateDevice(attr, options_.pkey, options_.cert, options_.caFile, options_.caPath);
  } else if (options_.transport == "uv") {
    transport::uv::attr ua;
    ua.hostname = attr.hostname;
    ua.iface = attr.iface;
    device_ = transport::uv::CreateDevice(ua);
  } else if (options_.transport == "ibverbs") {
    transport::ibverbs::attr ia;
    ia.name = options_.ibDevice;
    ia.index = options_.ibIndex;
    ia.port = options_.ibPort;
    device_ = transport::ibverbs::CreateDevice(ia);  // throws with the reason on this build
  } else {
    device_ = transport::tcp::CreateDevice(attr);
  }

  std::shared_ptr<rendezvous::Store> store;
  if (!options_.redisHost.empty()) {
    store = std::make_shared<rendezvous::RedisStore>(options_.redisHost, options_.redisPort);
  } else {
    store = std::make_shared<rendezvous::FileStore>(options_.sharedPath);
  }
  store = std::make_shared<rendezvous::PrefixStore>(options_.prefix, store);
  backing_ = std::make_shared<rendezvous::Context>(options_.contextRank, options_.contextSize, options_.base);
  backing_->connectFullMesh(store, device_);
  factory_.reset(new rendezvous::ContextFactory(backing_));
}

Runner::~Runner() = default;

long Runner::broadcastValue(long v) {
  BroadcastOptions o(backing_);
  o.setOutput(&v, 1);
  o.setRoot(0);
  o.setTag(0xB0000000u);
  broadcast(o);
  return v;
}

void Runner::printHeader() {
  if (options_.contextRank != 0) return;
  std::string line(112, '=');
  std::printf("%s\n%*s\n\n", line.c_str(), static_cast<int>(56 + options_.benchmark.size() / 2),
              options_.benchmark.c_str());
  std::printf("Device:      %s\n", device_->str().c_str());
  std::printf("Options:     processes=%d, inputs=%d, threads=%d, verify=%s%s%s\n", options_.contextSize,
              options_.inputs, options_.threads, options_.verify ? "true" : "false",
              options_.sync ? ", sync=true" : "", options_.busyPoll ? ", busy-poll=true" : "");
  std::printf("\n%s\n%*s\n\n", line.c_str(), 65, "BENCHMARK RESULTS");
  const char* u = options_.showNanos ? "ns" : "us";
  std::printf("%11s %10s %10s(%s) %10s(%s) %10s(%s) %10s(%s) %10s(%s) %12s %12s %11s\n", "size (B)", "elements", "min",
              u, "p50", u, "p99", u, "max", u, "dev p50", u, "algbw(GB/s)", "busbw(GB/s)", "iterations");
  std::fflush(stdout);
}

void Runner::printRow(size_t elements, size_t elementSize, const Distribution& host, const Distribution& dev,
                      double busFactor) {
  if (options_.contextRank != 0) return;
  const double div = options_.showNanos ? 1.0 : 1000.0;
  const double bytes = static_cast<double>(elements) * elementSize;
  // Same definition as the reference (runner.cc:499-508): bytes * samples / sum(latency), in GiB/s.
  const double algbw = host.sum() > 0 ? bytes * host.size() / (host.sum() / 1e9) / (1024.0 * 1024.0 * 1024.0) : 0.0;
  char devbuf[32] = "-";
  if (dev.size() > 0) std::snprintf(devbuf, sizeof(devbuf), "%.1f", dev.percentile(0.5) / div);
  char busbuf[32] = "-";
  if (busFactor > 0) {
    const double ref = dev.size() > 0 ? static_cast<double>(dev.percentile(0.5)) : static_cast<double>(host.percentile(0.5));
    if (ref > 0) std::snprintf(busbuf, sizeof(busbuf), "%.3f", bytes / (ref / 1e9) / 1e9 * busFactor);
  }
  std::printf("%11zu %10zu %14.1f %14.1f %14.1f %14.1f %18s %12.3f %12s %11zu\n", static_cast<size_t>(bytes), elements,
              host.min() / div, host.percentile(0.5) / div, host.percentile(0.99) / div, host.max() / div, devbuf,
              algbw, busbuf, host.size());
  std::fflush(stdout);
}

void Runner::runSize(const BenchmarkFactory& factory, size_t elements) {
  // One derived context + benchmark instance per thread (reference: --threads,
  // runner.cc:282-369). Contexts are minted sequentially (the factory is collective),
  // then the threads run concurrently and their samples are merged.
  const int nthreads = std::max(1, options_.threads);
  std::vector<std::shared_ptr<Context>> contexts;
  std::vector<Benchmark> benches;
  for (int t = 0; t < nthreads; t++) {
    auto context = factory_->makeContext(device_);
    context->base = options_.base;
    if (options_.sync) {
      for (int i = 0; i < context->size; i++) {
        auto& pair = context->getPair(i);
        if (pair) pair->setSync(true, options_.busyPoll);
      }
    }
    contexts.push_back(context);
    benches.push_back(factory(context, options_));
    benches.back().initialize(elements);
  }
  Benchmark& b = benches[0];

  auto hostBarrier = [&] {
    BarrierOptions o(backing_);
    o.setTag(0xB0000001u);
    barrier(o);
  };

  if (options_.verify && b.verify) {
    for (auto& x : benches) {
      x.run();
      x.verify();
    }
    hostBarrier();
  }
  // Warm-up; its median decides the iteration count (agreed through rank 0).
  Distribution warm;
  for (int i = 0; i < options_.warmupIterationCount; i++) {
    Timer t;
    b.run();
    warm.add(t.ns());
  }
  for (size_t t = 1; t < benches.size(); t++) {
    for (int i = 0; i < options_.warmupIterationCount; i++) benches[t].run();
  }
  long iterations = options_.iterationCount;
  if (iterations <= 0) {
    long median = std::max<long>(1, warm.percentile(0.5));
    iterations = broadcastValue(std::max<long>(1, options_.iterationTimeNanos / median));
  }
  Distribution host, dev;
  hostBarrier();
  Timer total;
  while (true) {
    std::vector<Distribution> th(nthreads), td(nthreads);
    auto body = [&](int t) {
      for (long i = 0; i < iterations; i++) {
        Timer tm;
        benches[t].run();
        th[t].add(tm.ns());
        if (benches[t].deviceNs) td[t].add(static_cast<long>(benches[t].deviceNs()));
      }
    };
    if (nthreads == 1) {
      body(0);
    } else {
      std::vector<std::thread> workers;
      for (int t = 0; t < nthreads; t++) workers.emplace_back(body, t);
      for (auto& w : workers) w.join();
    }
    for (int t = 0; t < nthreads; t++) {
      host.merge(th[t]);
      dev.merge(td[t]);
    }
    if (options_.iterationCount > 0) break;
    // Keep going (x1.2) until the minimum run time has been reached on rank 0.
    long enough = broadcastValue(total.ns() >= options_.iterationTimeNanos ? 1 : 0);
    if (enough) break;
    iterations = std::max<long>(1, static_cast<long>(iterations * 0.2));
  }
  hostBarrier();
  printRow(elements, b.elementSize, host, dev, b.busFactor);
  benches.clear();
  for (auto& c : contexts) c->closeConnections();
}

void Runner::run() {
  auto it = benchmarkRegistry().find(options_.benchmark);
  if (it == benchmarkRegistry().end()) GLB_THROW(Exception, "unknown benchmark: ", options_.benchmark);
  printHeader();
  if (options_.elements >= 0) {
    runSize(it->second, static_cast<size_t>(options_.elements));
  } else {
    // {1,2,5} x 10^k sweep: 100 .. 5e6 like the reference (runner.cc:269-279), or 1 .. 1e8.
    const size_t lo = options_.extendedSweep ? 1 : 100;
    const size_t hi = options_.extendedSweep ? 100000000 : 5000000;
    for (size_t i = lo; i <= hi; i *= 10) {
      for (size_t j : {1, 2, 5}) {
        if (i * j > hi) break;
        runSize(it->second, i * j);
      }
    }
  }
  if (options_.contextRank == 0) std::printf("\n%s\n", std::string(112, '=').c_str());
}

}  // namespace benchmark
}  // namespace glb
