// glb_selftest — the C++ API exercised without Python: P threads act as ranks (real TCP
// over loopback, in-process HashStore, like the reference's BaseTest::spawn,
// gloo/test/base_test.h:117-179) and run every host collective family with closed-form
// checks. Usage: glb_selftest [P ...]   (default: 1 2 3 4 7). Exit code 0 = all passed.
// Also the binary the sanitizer builds run when results matter, not just races.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "glb/allgather.h"
#include "glb/allgather_ring.h"
#include "glb/allgatherv.h"
#include "glb/allreduce.h"
#include "glb/allreduce_bcube.h"
#include "glb/allreduce_halving_doubling.h"
#include "glb/allreduce_ring.h"
#include "glb/allreduce_ring_chunked.h"
#include "glb/alltoall.h"
#include "glb/alltoallv.h"
#include "glb/barrier.h"
#include "glb/barrier_all_to_all.h"
#include "glb/broadcast.h"
#include "glb/broadcast_one_to_all.h"
#include "glb/gather.h"
#include "glb/math.h"
#include "glb/reduce.h"
#include "glb/reduce_scatter.h"
#include "glb/reduce_scatter_halving_doubling.h"
#include "glb/rendezvous/context.h"
#include "glb/rendezvous/hash_store.h"
#include "glb/scatter.h"
#include "glb/transport/tcp/device.h"

namespace {

std::mutex gPrint;
std::atomic<int> gFailures{0};

void fail(const std::string& what, int P, int rank, const std::string& detail) {
  std::lock_guard<std::mutex> g(gPrint);
  std::printf("FAIL  P=%d rank=%d  %s: %s\n", P, rank, what.c_str(), detail.c_str());
  gFailures++;
}

#define CHECK(cond, what)                                                                \
  do {                                                                                   \
    if (!(cond)) fail(what, P, r, std::string(#cond) + " (line " + std::to_string(__LINE__) + ")"); \
  } while (0)

using Ctx = std::shared_ptr<glb::Context>;

// rank q contributes j*P + q  ->  sum over ranks = j*P*P + P(P-1)/2
double expectedSum(size_t j, int P) { return static_cast<double>(j) * P * P + P * (P - 1) / 2.0; }

void runRank(int P, int r, const std::shared_ptr<glb::rendezvous::HashStore>& store) {
  auto dev = glb::transport::tcp::CreateDevice(glb::transport::tcp::attr("127.0.0.1"));
  auto rctx = std::make_shared<glb::rendezvous::Context>(r, P);
  rctx->setTimeout(std::chrono::seconds(30));
  rctx->connectFullMesh(store, dev);
  Ctx ctx = rctx;
  auto sumf = [](void* c, const void* a, const void* b, size_t n) { glb::sum<float>(c, a, b, n); };

  for (size_t count : {size_t(0), size_t(1), size_t(7), size_t(1000), size_t(300000)}) {
    // ---- new-style allreduce, in place and out of place, both algorithms
    for (auto alg : {glb::AllreduceOptions::RING, glb::AllreduceOptions::BCUBE}) {
      std::vector<float> in(count), out(count, -1.f);
      for (size_t j = 0; j < count; j++) in[j] = static_cast<float>((j % 1000) * P + r);
      glb::AllreduceOptions o(ctx);
      o.setAlgorithm(alg);
      o.setInput(in.data(), count);
      o.setOutput(out.data(), count);
      o.setReduceFunction(sumf);
      glb::allreduce(o);
      bool ok = true, untouched = true;
      for (size_t j = 0; j < count; j++) {
        ok = ok && out[j] == static_cast<float>(expectedSum(j % 1000, P));
        untouched = untouched && in[j] == static_cast<float>((j % 1000) * P + r);
      }
      CHECK(ok, "allreduce (new-style, out of place)");
      CHECK(untouched, "allreduce leaves its input alone");
      glb::AllreduceOptions o2(ctx);
      o2.setAlgorithm(alg);
      o2.setOutput(in.data(), count);
      o2.setReduceFunction(sumf);
      glb::allreduce(o2);
      ok = true;
      for (size_t j = 0; j < count; j++) ok = ok && in[j] == out[j];
      CHECK(ok, "allreduce (new-style, in place)");
    }
    // ---- old-style allreduce classes, run twice on one instance
    auto oldStyle = [&](const char* name, std::function<std::unique_ptr<glb::Algorithm>(std::vector<float*>)> make) {
      std::vector<float> a(count), b(count);
      auto algo = make({a.data(), b.data()});
      for (int it = 0; it < 2; it++) {
        for (size_t j = 0; j < count; j++) {
          a[j] = static_cast<float>((j % 500) * 2 * P + 2 * r);
          b[j] = static_cast<float>((j % 500) * 2 * P + 2 * r + 1);
        }
        algo->run();
        bool ok = true;
        const int S = 2 * P;  // 2 pointers per rank
        for (size_t j = 0; j < count; j++) {
          const float want = static_cast<float>(static_cast<double>(j % 500) * S * S + S * (S - 1) / 2.0);
          ok = ok && a[j] == want && b[j] == want;
        }
        CHECK(ok, name);
      }
    };
    oldStyle("AllreduceRing", [&](std::vector<float*> p) {
      return std::unique_ptr<glb::Algorithm>(new glb::AllreduceRing<float>(ctx, p, count));
    });
    oldStyle("AllreduceRingChunked", [&](std::vector<float*> p) {
      return std::unique_ptr<glb::Algorithm>(new glb::AllreduceRingChunked<float>(ctx, p, count));
    });
    oldStyle("AllreduceHalvingDoubling", [&](std::vector<float*> p) {
      return std::unique_ptr<glb::Algorithm>(new glb::AllreduceHalvingDoubling<float>(ctx, p, count));
    });
    oldStyle("AllreduceBcube", [&](std::vector<float*> p) {
      return std::unique_ptr<glb::Algorithm>(new glb::AllreduceBcube<float>(ctx, p, count));
    });
  }

  // ---- allgather / allgatherv / AllgatherRing
  {
    const size_t n = 513;
    std::vector<int> in(n, r), out(n * P, -1);
    glb::AllgatherOptions o(ctx);
    o.setInput(in.data(), n);
    o.setOutput(out.data(), n * P);
    glb::allgather(o);
    bool ok = true;
    for (int q = 0; q < P; q++) ok = ok && out[q * n] == q && out[q * n + n - 1] == q;
    CHECK(ok, "allgather");
    std::vector<size_t> counts(P);
    for (int q = 0; q < P; q++) counts[q] = static_cast<size_t>(q + 1);
    std::vector<int> vin(r + 1, r), vout(static_cast<size_t>(P) * (P + 1) / 2, -1);
    glb::AllgathervOptions ov(ctx);
    ov.setInput(vin.data(), vin.size());
    ov.setOutput(vout.data(), counts);
    glb::allgatherv(ov);
    ok = true;
    size_t off = 0;
    for (int q = 0; q < P; q++) {
      for (size_t k = 0; k < counts[q]; k++) ok = ok && vout[off + k] == q;
      off += counts[q];
    }
    CHECK(ok, "allgatherv");
    std::vector<int> ring(n * P, -1);
    glb::AllgatherRing<int> ag(ctx, {in.data()}, ring.data(), n);
    ag.run();
    CHECK(ring == out, "AllgatherRing");
  }
  // ---- alltoall / alltoallv
  {
    const size_t n = 64;
    std::vector<int> in(n * P), out(n * P, -1);
    for (int q = 0; q < P; q++) std::fill(in.begin() + q * n, in.begin() + (q + 1) * n, r * 100 + q);
    glb::AlltoallOptions o(ctx);
    o.setInput(in.data(), n * P);
    o.setOutput(out.data(), n * P);
    glb::alltoall(o);
    bool ok = true;
    for (int q = 0; q < P; q++) ok = ok && out[q * n] == q * 100 + r;
    CHECK(ok, "alltoall");
    // rank r sends (q + 1) elements to rank q, so it receives (r + 1) from everyone
    std::vector<int64_t> sc(P), rc(P, r + 1);
    for (int q = 0; q < P; q++) sc[q] = q + 1;
    std::vector<int> vin(static_cast<size_t>(P) * (P + 1) / 2, r), vout(static_cast<size_t>(P) * (r + 1), -1);
    glb::AlltoallvOptions ov(ctx);
    ov.setInput(vin.data(), sc);
    ov.setOutput(vout.data(), rc);
    glb::alltoallv(ov);
    ok = true;
    for (int q = 0; q < P; q++) ok = ok && vout[static_cast<size_t>(q) * (r + 1)] == q;
    CHECK(ok, "alltoallv");
  }
  // ---- broadcast / BroadcastOneToAll / gather / scatter / reduce
  for (int root = 0; root < P; root += std::max(1, P - 1)) {
    std::vector<double> v(2049, r == root ? 3.25 : 0.0);
    glb::BroadcastOptions bo(ctx);
    bo.setOutput(v.data(), v.size());
    bo.setRoot(root);
    glb::broadcast(bo);
    CHECK(v.front() == 3.25 && v.back() == 3.25, "broadcast");
    std::vector<double> w(100, r == root ? 7.5 : 0.0);
    glb::BroadcastOneToAll<double> b1(ctx, {w.data()}, w.size(), root);
    b1.run();
    CHECK(w.front() == 7.5 && w.back() == 7.5, "BroadcastOneToAll");
    std::vector<int> gi(5, r), go(5 * P, -1);
    glb::GatherOptions go_(ctx);
    go_.setInput(gi.data(), gi.size());
    if (r == root) go_.setOutput(go.data(), go.size());
    go_.setRoot(root);
    glb::gather(go_);
    if (r == root) {
      bool ok = true;
      for (int q = 0; q < P; q++) ok = ok && go[q * 5] == q;
      CHECK(ok, "gather");
    }
    std::vector<std::vector<int>> parts(P, std::vector<int>(3));
    for (int q = 0; q < P; q++) std::fill(parts[q].begin(), parts[q].end(), q * 11);
    std::vector<int> so(3, -1);
    glb::ScatterOptions so_(ctx);
    if (r == root) {
      std::vector<int*> ptrs;
      for (auto& p : parts) ptrs.push_back(p.data());
      so_.setInputs(ptrs, 3);
    }
    so_.setOutput(so.data(), so.size());
    so_.setRoot(root);
    glb::scatter(so_);
    CHECK(so[0] == r * 11 && so[2] == r * 11, "scatter");
    std::vector<float> ri(1000, static_cast<float>(r + 1)), ro(1000, 0.f);
    glb::ReduceOptions ro_(ctx);
    ro_.setInput(ri.data(), ri.size());
    ro_.setOutput(ro.data(), ro.size());
    ro_.setRoot(root);
    ro_.setReduceFunction(sumf);
    glb::reduce(ro_);
    if (r == root) CHECK(ro[0] == P * (P + 1) / 2.0f && ro[999] == ro[0], "reduce");
  }
  // ---- reduce_scatter (new-style) and ReduceScatterHalvingDoubling with re-runs
  {
    const size_t n = 4096 + 3;
    std::vector<float> in(n), out(n, -1.f);
    for (size_t j = 0; j < n; j++) in[j] = static_cast<float>((j % 100) * P + r);
    const auto mine = glb::detail::subRange({0, n}, P, r);
    glb::ReduceScatterOptions o(ctx);
    o.setInput(in.data(), n);
    o.setOutput(out.data(), mine.len);
    o.setReduceFunction(sumf);
    glb::reduce_scatter(o);
    bool ok = true;
    for (size_t j = 0; j < mine.len; j++) ok = ok && out[j] == static_cast<float>(expectedSum((mine.off + j) % 100, P));
    CHECK(ok, "reduce_scatter (new-style)");
    std::vector<int> recv(P);
    for (int q = 0; q < P; q++) recv[q] = static_cast<int>(glb::detail::subRange({0, n}, P, q).len);
    std::vector<float> buf(n);
    glb::ReduceScatterHalvingDoubling<float> rs(ctx, {buf.data()}, n, recv);
    for (int it = 0; it < 5; it++) {
      for (size_t j = 0; j < n; j++) buf[j] = static_cast<float>((j % 100) * P + r) * (it + 1);
      rs.run();
      ok = true;
      for (size_t j = 0; j < mine.len; j++) {
        ok = ok && buf[j] == static_cast<float>(expectedSum((mine.off + j) % 100, P) * (it + 1));
      }
      CHECK(ok, "ReduceScatterHalvingDoubling (re-run)");
    }
  }
  // ---- point to point incl. recv-from-any, barriers, derived contexts
  if (P > 1) {
    const int right = (r + 1) % P, left = (r - 1 + P) % P;
    std::vector<int> a(300000, r), b(300000, -1);  // above the single-copy threshold
    auto ua = ctx->createUnboundBuffer(a.data(), a.size() * sizeof(int));
    auto ub = ctx->createUnboundBuffer(b.data(), b.size() * sizeof(int));
    ub->recv(left, 0x42);
    ua->send(right, 0x42);
    ub->waitRecv();
    ua->waitSend();
    CHECK(b.front() == left && b.back() == left, "send/recv (large)");
    int token = r, got = -1;
    auto ut = ctx->createUnboundBuffer(&token, sizeof(token));
    auto ug = ctx->createUnboundBuffer(&got, sizeof(got));
    if (r == 0) {
      std::vector<int> any;
      for (int q = 1; q < P; q++) any.push_back(q);
      long seen = 0;
      for (int q = 1; q < P; q++) {
        ug->recv(any, 0x43);
        int src = -1;
        ug->waitRecv(&src);
        CHECK(got == src, "recv-from-any delivers the sender's payload");
        seen += src;
      }
      CHECK(seen == static_cast<long>(P) * (P - 1) / 2, "recv-from-any saw every rank once");
    } else {
      ut->send(0, 0x43);
      ut->waitSend();
    }
  }
  {
    glb::BarrierOptions bo(ctx);
    glb::barrier(bo);
    glb::BarrierAllToAll b2(ctx);
    b2.run();
    b2.run();
    glb::rendezvous::ContextFactory factory(ctx);
    auto dev2 = glb::transport::tcp::CreateDevice(glb::transport::tcp::attr("127.0.0.1"));
    auto derived = factory.makeContext(dev2);
    std::vector<float> x(10, 1.f);
    glb::AllreduceOptions o(derived);
    o.setOutput(x.data(), x.size());
    o.setReduceFunction(sumf);
    glb::allreduce(o);
    CHECK(x[0] == static_cast<float>(P), "ContextFactory-derived context");
    glb::BarrierOptions b3(derived);
    glb::barrier(b3);
    derived->closeConnections();
  }
  glb::BarrierOptions fin(ctx);
  glb::barrier(fin);
  rctx->closeConnections();
}

}  // namespace

int main(int argc, char** argv) {
  std::vector<int> sizes;
  for (int i = 1; i < argc; i++) sizes.push_back(std::atoi(argv[i]));
  if (sizes.empty()) sizes = {1, 2, 3, 4, 7};
  for (int P : sizes) {
    auto store = std::make_shared<glb::rendezvous::HashStore>();
    std::vector<std::thread> ths;
    std::atomic<int> crashed{0};
    for (int r = 0; r < P; r++) {
      ths.emplace_back([&, r] {
        try {
          runRank(P, r, store);
        } catch (const std::exception& e) {
          fail("exception", P, r, e.what());
          crashed++;
        }
      });
    }
    for (auto& t : ths) t.join();
    std::printf("P=%d %s\n", P, gFailures.load() == 0 ? "ok" : "FAILED");
    if (crashed.load() > 0) break;
  }
  std::printf("%s (%d failure%s)\n", gFailures.load() == 0 ? "PASS" : "FAIL", gFailures.load(),
              gFailures.load() == 1 ? "" : "s");
  return gFailures.load() == 0 ? 0 : 1;
}
