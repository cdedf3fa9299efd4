// Single-GPU, single-process self-test of the collective kernels ("virtual-rank loopback").
//
// Nsight Compute serialises kernels, so two ranks whose kernels wait for each other can
// never be profiled together — and a box with one GPU cannot host two real ranks at full
// speed. PeerContext::loopbackComm(P) builds kernel arguments with which ONE launch plays
// rank 0 of P ranks (the flags it posts for its peers are the ones it waits for); the
// "peer" buffers are P distinct local allocations. Every hot kernel — including the
// P = 2 / 4 / 8 specialisations and, where the driver allows a one-device multicast object,
// the multimem (NVLS) kernels — runs exactly the code it runs across GPUs, and rank 0's
// part of the result has a closed form that is checked here on the host.
#include "glb/cuda/selftest.h"

#include <cmath>
#include <cstring>
#include <sstream>

#include "glb/common/utils.h"
#include "glb/cuda/kernels.h"
#include "glb/cuda/local_ops.h"
#include "glb/cuda/schedules.h"

namespace glb {
namespace cuda {

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  explicit DevBuf(size_t n) : bytes(n) { GLB_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 16))); }
  ~DevBuf() { cudaFree(p); }
  DevBuf(const DevBuf&) = delete;
};

void shareOfHost(size_t n, int parts, int r, size_t& b, size_t& e) {
  const size_t base = n / parts, rem = n % parts;
  b = r * base + std::min<size_t>(r, rem);
  e = b + base + (static_cast<size_t>(r) < rem ? 1 : 0);
}

std::vector<float> download(const void* dev, size_t count, cudaStream_t s) {
  std::vector<float> h(count);
  GLB_CUDA_CHECK(cudaMemcpyAsync(h.data(), dev, count * 4, cudaMemcpyDeviceToHost, s));
  GLB_CUDA_CHECK(cudaStreamSynchronize(s));
  return h;
}

bool close(float a, double want) { return std::fabs(static_cast<double>(a) - want) <= 1e-5 * std::fabs(want) + 1e-6; }

float bf16ToFloat(uint16_t h) {
  uint32_t w = static_cast<uint32_t>(h) << 16;
  float f;
  std::memcpy(&f, &w, 4);
  return f;
}

class Runner {
 public:
  Runner(PeerContext& pc, cudaStream_t s) : pc_(pc), s_(s) {}
  std::vector<SelfTestResult> results;

  template <typename F>
  void run(const std::string& name, F&& body) {
    SelfTestResult r;
    r.name = name;
    try {
      std::string why = body();
      GLB_CUDA_CHECK(cudaStreamSynchronize(s_));
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) why = cudaGetErrorString(e);
      r.ok = why.empty();
      r.detail = why;
    } catch (const std::exception& e) {
      r.ok = false;
      r.detail = e.what();
    }
    if (r.detail.rfind("skipped", 0) == 0) {
      r.ok = true;
      r.skipped = true;
    }
    results.push_back(r);
  }

 private:
  PeerContext& pc_;
  cudaStream_t s_;
};

}  // namespace

std::vector<SelfTestResult> loopbackSelfTest(PeerContext& pc, cudaStream_t stream, size_t count) {
  GLB_ENFORCE(count >= 4096, "loopbackSelfTest: count too small");
  DeviceGuard g(pc.device);
  pc.checkHealth();
  GLB_CUDA_CHECK(cudaStreamSynchronize(stream));
  Runner R(pc, stream);
  const int dtF = static_cast<int>(DataType::FLOAT32);
  (void)dtF;
  const size_t n = count / 4 * 4 + 3;  // not a multiple of the pack width: exercises the tails
  const PeerPtrs poolStage = pc.stagePtrs(0);
  char* myStage = static_cast<char*>(poolStage.p[pc.rank]);
  const size_t half = std::min<size_t>(roundUp(std::max<size_t>(tuning().oneShotMaxBytes, 4096), 4096), pc.stageBytes() / 4);

  for (int P : {2, 4, 8}) {
    const CommArgs ca = pc.loopbackComm(P);
    std::vector<std::unique_ptr<DevBuf>> bufs;
    PeerPtrs pp;
    std::memset(&pp, 0, sizeof(pp));
    for (int r = 0; r < P; r++) {
      bufs.emplace_back(new DevBuf(n * 4));
      pp.p[r] = bufs[r]->p;
    }
    auto fillAll = [&] {
      for (int r = 0; r < P; r++) launchFill(pp.p[r], n, DataType::FLOAT32, r, P, stream);  // value = i*P + r
    };
    const double tri = P * (P - 1) / 2.0;
    const size_t nvec = n / 4;
    size_t vb, ve;
    shareOfHost(nvec, P, 0, vb, ve);
    size_t tb, te;
    shareOfHost(n - nvec * 4, P, 0, tb, te);

    // ---- two-shot, every unroll of this P ------------------------------------------------
    for (int unroll : {1, 2, 4, 8}) {
      if ((P == 2 && unroll == 1) || (P == 4 && unroll == 8) || (P == 8 && unroll > 2)) continue;
      R.run(strcat_all("twoShotAllreduceKernel<float,", P, ",", unroll, ">"), [&]() -> std::string {
        fillAll();
        LaunchCfg cfg;
        cfg.blocks = std::min(pc.coResidentBlocks(twoShotKernelFor(DataType::FLOAT32, P, unroll)), 32);
        cfg.unroll = unroll;
        launchTwoShotAllreduce(ca, pp, n, DataType::FLOAT32, ReduceOp::SUM, 0.5f, true, LocalPtrs(), cfg, stream);
        noteLaunch();
        auto last = download(pp.p[P - 1], n, stream);
        for (size_t i = 0; i < n; i++) {
          const bool mine = (i / 4 >= vb && i / 4 < ve && i < nvec * 4) || (i >= nvec * 4 + tb && i < nvec * 4 + te);
          const double want = mine ? 0.5 * (static_cast<double>(i) * P * P + tri) : static_cast<double>(i) * P + (P - 1);
          if (!close(last[i], want)) return strcat_all("element ", i, ": got ", last[i], " want ", want);
        }
        return "";
      });
    }

    // ---- reduce-scatter / reduce pull -----------------------------------------------------
    R.run(strcat_all("reducePullKernel<float,", P, ">"), [&]() -> std::string {
      fillAll();
      std::vector<size_t> off(P + 1, 0), len(P, 0);
      for (int r = 0; r < P; r++) {
        len[r] = n / P + (static_cast<size_t>(r) < n % P ? 1 : 0);
        off[r + 1] = off[r] + len[r];
      }
      DevBuf out(len[0] * 4);
      launchReducePull(ca, pp, nullptr, out.p, off.data(), len.data(), DataType::FLOAT32, ReduceOp::SUM, 1.0f, true, false,
                       std::min(pc.maxBlocks(), 16), stream);
      noteLaunch();
      auto h = download(out.p, len[0], stream);
      for (size_t i = 0; i < len[0]; i++) {
        if (!close(h[i], static_cast<double>(i) * P * P + tri)) return strcat_all("element ", i, ": got ", h[i]);
      }
      return "";
    });

    // ---- cast epilogue (f32 in, bf16 out) -----------------------------------------------------
    R.run(strcat_all("castAllreduceKernel<float,bf16> P=", P), [&]() -> std::string {
      for (int r = 0; r < P; r++) launchFill(pp.p[r], n, DataType::FLOAT32, r % 3, 0.0, stream);  // small exact values
      std::vector<std::unique_ptr<DevBuf>> outs;
      PeerPtrs po;
      std::memset(&po, 0, sizeof(po));
      for (int r = 0; r < P; r++) {
        outs.emplace_back(new DevBuf(n * 2));
        GLB_CUDA_CHECK(cudaMemsetAsync(outs[r]->p, 0, n * 2, stream));
        po.p[r] = outs[r]->p;
      }
      launchCastAllreduce(ca, pp, nullptr, po, n, DataType::FLOAT32, DataType::BFLOAT16, ReduceOp::SUM, 2.0f, true,
                          std::min(pc.maxBlocks(), 16), stream);
      noteLaunch();
      std::vector<uint16_t> h(n);
      GLB_CUDA_CHECK(cudaMemcpyAsync(h.data(), po.p[P - 1], n * 2, cudaMemcpyDeviceToHost, stream));
      GLB_CUDA_CHECK(cudaStreamSynchronize(stream));
      double sum = 0;
      for (int r = 0; r < P; r++) sum += r % 3;
      size_t ib, ie, xb, xe;
      shareOfHost(n / 8, P, 0, ib, ie);
      shareOfHost(n - n / 8 * 8, P, 0, xb, xe);
      for (size_t i = 0; i < n; i++) {
        const bool mine = (i / 8 >= ib && i / 8 < ie && i < n / 8 * 8) || (i >= n / 8 * 8 + xb && i < n / 8 * 8 + xe);
        const double want = mine ? 2.0 * sum : 0.0;
        if (std::fabs(bf16ToFloat(h[i]) - want) > 1e-2 * std::fabs(want) + 1e-6) {
          return strcat_all("element ", i, ": got ", bf16ToFloat(h[i]), " want ", want);
        }
      }
      return "";
    });

    // ---- one-shot (pull + push) through the pool; every virtual peer aliases my pool --------
    R.run(strcat_all("oneShotAllreduceKernel<float> P=", P), [&]() -> std::string {
      const size_t m = std::min<size_t>(n, half / 4 / P) / 4 * 4 - 1;
      PeerPtrs alias;
      std::memset(&alias, 0, sizeof(alias));
      for (int r = 0; r < P; r++) alias.p[r] = myStage;
      GLB_CUDA_CHECK(cudaMemsetAsync(myStage, 0, 2 * half, stream));
      launchFill(pp.p[0], m, DataType::FLOAT32, 1.0, 1.0, stream);
      setOneShotPush(false);
      launchOneShotAllreduce(ca, pp.p[0], pp.p[1], m, DataType::FLOAT32, ReduceOp::SUM, 1.0f, alias, half, LocalPtrs(), 2, stream);
      noteLaunch();
      auto pull = download(pp.p[1], m, stream);  // P reads of my own staged copy
      setOneShotPush(true);
      GLB_CUDA_CHECK(cudaMemsetAsync(myStage, 0, 2 * half, stream));
      launchOneShotAllreduce(ca, pp.p[0], pp.p[1], m, DataType::FLOAT32, ReduceOp::SUM, 1.0f, alias, half, LocalPtrs(), 2, stream);
      noteLaunch();
      auto push = download(pp.p[1], m, stream);  // slot 0 holds my data, the other slots are zero
      for (size_t i = 0; i < m; i++) {
        if (!close(pull[i], P * (1.0 + i))) return strcat_all("pull: element ", i, ": got ", pull[i]);
        if (!close(push[i], 1.0 + i)) return strcat_all("push: element ", i, ": got ", push[i]);
      }
      return "";
    });

    // ---- data movement ---------------------------------------------------------------------
    R.run(strcat_all("broadcastKernel(direct) P=", P), [&]() -> std::string {
      fillAll();
      launchBroadcast(ca, pp, nullptr, n * 4, 0, 0, true, std::min(2 * pc.maxBlocks(), 32), 0, stream);
      noteLaunch();
      auto h = download(pp.p[P - 1], n, stream);
      for (size_t i = 0; i < n; i++) {
        if (!close(h[i], static_cast<double>(i) * P)) return strcat_all("element ", i, ": got ", h[i]);
      }
      return "";
    });
    R.run(strcat_all("gatherPushKernel(allgather) P=", P), [&]() -> std::string {
      const size_t per = n / P;
      std::vector<size_t> off(P), len(P, per * 4);
      for (int r = 0; r < P; r++) off[r] = r * per * 4;
      DevBuf in(per * 4);
      launchFill(in.p, per, DataType::FLOAT32, 7.0, 1.0, stream);
      for (int r = 0; r < P; r++) GLB_CUDA_CHECK(cudaMemsetAsync(pp.p[r], 0, n * 4, stream));
      launchGatherPush(ca, in.p, pp, nullptr, off.data(), len.data(), -1, true, std::min(2 * pc.maxBlocks(), 32), stream);
      noteLaunch();
      auto h = download(pp.p[P - 1], n, stream);
      for (size_t i = 0; i < per; i++) {
        if (!close(h[i], 7.0 + i)) return strcat_all("element ", i, ": got ", h[i]);
      }
      return h[per] == 0.0f ? "" : "wrote outside my block";
    });
    R.run(strcat_all("gatherBulkKernel(allgather, TMA) P=", P), [&]() -> std::string {
      const size_t per = n / P / 4 * 4;  // 16-byte granular blocks
      std::vector<size_t> off(P), len(P, per * 4);
      for (int r = 0; r < P; r++) off[r] = r * per * 4;
      DevBuf in(per * 4);
      launchFill(in.p, per, DataType::FLOAT32, 3.0, 1.0, stream);
      for (int r = 0; r < P; r++) GLB_CUDA_CHECK(cudaMemsetAsync(pp.p[r], 0, n * 4, stream));
      launchGatherPush(ca, in.p, pp, nullptr, off.data(), len.data(), -1, true,
                       std::min(pc.coResidentBlocks(gatherBulkKernelPtr(), 128), 16), stream, /*tma=*/true);
      noteLaunch();
      for (int r : {0, P - 1}) {
        auto h = download(pp.p[r], n, stream);
        for (size_t i = 0; i < per; i++) {
          if (!close(h[i], 3.0 + i)) return strcat_all("dst ", r, " element ", i, ": got ", h[i]);
        }
        if (h[per] != 0.0f) return "wrote outside my block";
      }
      return "";
    });
    R.run(strcat_all("alltoallPushKernel P=", P), [&]() -> std::string {
      const size_t per = n / P / 4 * 4;
      std::vector<size_t> soff(P), slen(P, per * 4), doff(P, 0);
      for (int r = 0; r < P; r++) soff[r] = r * per * 4;
      DevBuf in(per * 4 * P);
      launchFill(in.p, per * P, DataType::FLOAT32, 0.0, 1.0, stream);
      for (int r = 0; r < P; r++) GLB_CUDA_CHECK(cudaMemsetAsync(pp.p[r], 0, n * 4, stream));
      launchAlltoallPush(ca, in.p, pp, soff.data(), slen.data(), doff.data(), nullptr, -1, true,
                         std::min(2 * pc.maxBlocks(), 32), stream);
      noteLaunch();
      auto h = download(pp.p[P - 1], per, stream);  // chunk P-1 of my input lands in slot 0 of rank P-1
      for (size_t i = 0; i < per; i++) {
        if (!close(h[i], static_cast<double>((P - 1) * per + i))) return strcat_all("element ", i, ": got ", h[i]);
      }
      return "";
    });

    // ---- literal schedule executor (halving-doubling, plain and pipelined) -------------------
    R.run(strcat_all("scheduleKernel<float>(halving_doubling) P=", P), [&]() -> std::string {
      fillAll();
      for (int pipelined = 0; pipelined < 2; pipelined++) {
        Schedule sc = pipelined ? buildHalvingDoublingPipelinedSchedule(0, P, n, 4, 2) : buildHalvingDoublingSchedule(0, P, n, 4);
        DevBuf table(sc.steps.size() * sizeof(SchedStep));
        GLB_CUDA_CHECK(cudaMemcpyAsync(table.p, sc.steps.data(), sc.steps.size() * sizeof(SchedStep), cudaMemcpyHostToDevice, stream));
        PeerPtrs st = pc.stagePtrs(pc.stageBytes() / 2);
        launchSchedule(ca, pp, st, static_cast<const SchedStep*>(table.p), static_cast<int>(sc.steps.size()), scheduleBarriers(sc),
                       DataType::FLOAT32, ReduceOp::SUM, 1.0f, n, true, std::min(pc.maxBlocks(), 16), stream);
        noteLaunch();
        GLB_CUDA_CHECK(cudaStreamSynchronize(stream));
      }
      return "";  // rank 0's partial view has no closed form here; completion without a hang is the check
    });
  }

  // ---- flag-in-data kernels: one rank (no peers to hear from), fused scale + cast -------------
  {
    const CommArgs one = pc.loopbackComm(1);
    R.run("llAllreduceKernel<float,float>", [&]() -> std::string {
      const size_t m = std::min<size_t>(pc.llMaxBytes() / 4, 4099);
      DevBuf in(m * 4), out(m * 4);
      launchFill(in.p, m, DataType::FLOAT32, 1.0, 1.0, stream);
      launchLLAllreduce(one, in.p, out.p, m, DataType::FLOAT32, DataType::FLOAT32, ReduceOp::SUM, 0.25f, pc.llPtrs(),
                        pc.llSrcStride(), pc.llParityStride(), LocalPtrs(), 4, 256, stream);
      noteLaunch();
      auto h = download(out.p, m, stream);
      for (size_t i = 0; i < m; i++) {
        if (!close(h[i], 0.25 * (1.0 + i))) return strcat_all("element ", i, ": got ", h[i]);
      }
      return "";
    });
    R.run("llAllreduceKernel<float,bf16> (cast epilogue)", [&]() -> std::string {
      const size_t m = 1001;
      DevBuf in(m * 4), out(m * 2);
      launchFill(in.p, m, DataType::FLOAT32, 3.0, 0.0, stream);
      launchLLAllreduce(one, in.p, out.p, m, DataType::FLOAT32, DataType::BFLOAT16, ReduceOp::SUM, 1.0f, pc.llPtrs(),
                        pc.llSrcStride(), pc.llParityStride(), LocalPtrs(), 1, 256, stream);
      noteLaunch();
      std::vector<uint16_t> h(m);
      GLB_CUDA_CHECK(cudaMemcpyAsync(h.data(), out.p, m * 2, cudaMemcpyDeviceToHost, stream));
      GLB_CUDA_CHECK(cudaStreamSynchronize(stream));
      for (size_t i = 0; i < m; i++) {
        if (bf16ToFloat(h[i]) != 3.0f) return strcat_all("element ", i, ": got ", bf16ToFloat(h[i]));
      }
      return "";
    });
    R.run("llExchangeKernel", [&]() -> std::string {
      const size_t bytes = 1000;
      DevBuf in(bytes), out(bytes);
      GLB_CUDA_CHECK(cudaMemsetAsync(in.p, 0x5a, bytes, stream));
      GLB_CUDA_CHECK(cudaMemsetAsync(out.p, 0, bytes, stream));
      launchLLExchange(one, in.p, out.p, bytes, 0, pc.llPtrs(), pc.llSrcStride(), pc.llParityStride(), 1, 128, stream);
      noteLaunch();
      std::vector<unsigned char> h(bytes);
      GLB_CUDA_CHECK(cudaMemcpyAsync(h.data(), out.p, bytes, cudaMemcpyDeviceToHost, stream));
      GLB_CUDA_CHECK(cudaStreamSynchronize(stream));
      for (size_t i = 0; i < bytes; i++) {
        if (h[i] != 0x5a) return strcat_all("byte ", i, " wrong");
      }
      return "";
    });
  }

  // ---- pipelined kernel for plain pointers (P virtual ranks alias my pool) ---------------------
  for (int P : {2, 8}) {
    R.run(strcat_all("pipelinedAllreduceKernel<float,false,", P, ">"), [&]() -> std::string {
      const CommArgs ca = pc.loopbackComm(P);
      const size_t bulkOff = 2 * half;
      PeerPtrs st;
      std::memset(&st, 0, sizeof(st));
      for (int r = 0; r < P; r++) st.p[r] = myStage + bulkOff;
      const int blocks = std::min(pc.coResidentBlocks(pipelinedKernelFor(DataType::FLOAT32, false, P)), 8), tile = 64;
      const size_t chunkVecs = static_cast<size_t>(P) * blocks * tile;
      GLB_ENFORCE_LE(3 * chunkVecs * 16, pc.stageBytes() - bulkOff, "pool too small for the pipelined self-test");
      DevBuf in(n * 4), out(n * 4);
      launchFill(in.p, n, DataType::FLOAT32, 1.0, 0.0, stream);
      GLB_CUDA_CHECK(cudaMemsetAsync(out.p, 0, n * 4, stream));
      launchPipelinedAllreduce(ca, in.p, out.p, n, DataType::FLOAT32, ReduceOp::SUM, 1.0f, st, nullptr, tile, 256, LocalPtrs(),
                               blocks, stream);
      noteLaunch();
      auto h = download(out.p, n, stream);
      // Rank 0 reduces share 0 of every chunk (P reads of its own copy -> P x 1); the other
      // shares pass through the pool untouched.
      for (size_t i = 0; i < n; i++) {
        const size_t w = (i / 4) % chunkVecs;
        const double want = w < static_cast<size_t>(blocks) * tile ? P : 1.0;
        if (!close(h[i], want)) return strcat_all("element ", i, ": got ", h[i], " want ", want);
      }
      return "";
    });
  }

  // ---- point to point: fused send + recv through my own mailbox --------------------------------
  R.run("p2pKernel(sendrecv)", [&]() -> std::string {
    const CommArgs ca = pc.loopbackComm(2);
    const auto& o = pc.options();
    PeerPtrs mb = pc.mailboxPtrs();
    char* base = static_cast<char*>(mb.p[pc.rank]);
    // "rank 1"'s box for source 0 is my box for source 1.
    mb.p[0] = base;
    mb.p[1] = base + pc.mailboxStride();
    const size_t bytes = 3 * o.p2pSlotBytes * o.p2pSlots / 2 + 20;  // more than the ring holds: must stream
    DevBuf src(bytes), dst(bytes);
    launchFill(src.p, bytes / 4, DataType::FLOAT32, 0.0, 1.0, stream);
    GLB_CUDA_CHECK(cudaMemsetAsync(dst.p, 0, bytes, stream));
    launchP2p(ca, src.p, bytes / 4 * 4, 1, dst.p, bytes / 4 * 4, 1, mb, pc.mailboxStride(), o.p2pSlotBytes, o.p2pSlots,
              std::min(o.p2pLanes, std::max(1, pc.maxBlocks() / 2)), stream);
    noteLaunch();
    auto h = download(dst.p, bytes / 4, stream);
    for (size_t i = 0; i < bytes / 4; i++) {
      if (!close(h[i], static_cast<double>(i))) return strcat_all("element ", i, ": got ", h[i]);
    }
    return "";
  });

  for (bool tma : {false, true}) {
    R.run(tma ? "exchangeKernel<true> (zero-copy sendrecv, TMA)" : "exchangeKernel<false> (zero-copy sendrecv)",
          [&]() -> std::string {
            const CommArgs ca = pc.loopbackComm(2);
            const size_t bytes = (5u << 20) + 16 + 4;
            DevBuf src(bytes), dst(bytes);
            for (int round = 0; round < 2; round++) {  // twice: the counters must carry over
              launchFill(src.p, bytes / 4, DataType::FLOAT32, 0.0, 1.0 + round, stream);
              GLB_CUDA_CHECK(cudaMemsetAsync(dst.p, 0, bytes, stream));
              launchExchange(ca, src.p, bytes, 1, dst.p, bytes, 1, std::min(pc.maxBlocks(), 24), tma, stream);
              noteLaunch();
              auto h = download(dst.p, bytes / 4, stream);
              for (size_t i = 0; i < bytes / 4; i++) {
                if (!close(h[i], static_cast<double>(i) * (1.0 + round))) return strcat_all("round ", round, " element ", i, ": got ", h[i]);
              }
            }
            return "";
          });
  }

  R.run("peerBulkCopyKernel (put / get, TMA)", [&]() -> std::string {
    const size_t bytes = (3u << 20) + 48 + 5;  // 16-byte body + a byte tail
    DevBuf src(bytes), dst(bytes);
    launchFill(src.p, bytes / 4, DataType::FLOAT32, 0.0, 1.0, stream);
    GLB_CUDA_CHECK(cudaMemsetAsync(dst.p, 0, bytes, stream));
    launchPeerCopy(dst.p, src.p, bytes, 24, stream, /*tma=*/true);
    noteLaunch();
    auto h = download(dst.p, bytes / 4, stream);
    for (size_t i = 0; i < bytes / 4; i++) {
      if (!close(h[i], static_cast<double>(i))) return strcat_all("element ", i, ": got ", h[i]);
    }
    return "";
  });

  // ---- NVLS kernels on a one-device multicast object (when the driver allows it) ----------------
  R.run("nvlsAllreduceKernel<float,4> (1-device multicast)", [&]() -> std::string {
    std::string why;
    auto mcbuf = pc.tryAllocMulticastLoopback(n * 4, &why);
    if (!mcbuf) return "skipped: " + why;
    const CommArgs ca = pc.loopbackComm(4);
    PeerPtrs pp;
    std::memset(&pp, 0, sizeof(pp));
    for (int r = 0; r < 4; r++) pp.p[r] = mcbuf->local;
    launchFill(mcbuf->local, n, DataType::FLOAT32, 1.0, 1.0, stream);
    LaunchCfg cfg;
    cfg.blocks = std::min(pc.maxBlocks(), 32);
    cfg.unroll = 4;
    launchNvlsAllreduce(ca, mcbuf->mc, pp, n / 4 * 4, DataType::FLOAT32, 3.0f, LocalPtrs(), cfg, stream);
    noteLaunch();
    auto h = download(mcbuf->local, n / 4 * 4, stream);
    size_t vb, ve;
    shareOfHost(n / 4, 4, 0, vb, ve);
    for (size_t i = 0; i < n / 4 * 4; i++) {
      const double want = (i / 4 >= vb && i / 4 < ve) ? 3.0 * (1.0 + i) : 1.0 + i;
      if (!close(h[i], want)) return strcat_all("element ", i, ": got ", h[i], " want ", want);
    }
    return "";
  });

  // ---- local kernels ------------------------------------------------------------------------------
  R.run("localAllreduceManyKernel<float>", [&]() -> std::string {
    DevBuf a(n * 4), b(n * 4), c(n * 4);
    launchFill(a.p, n, DataType::FLOAT32, 0.0, 3.0, stream);
    launchFill(b.p, n, DataType::FLOAT32, 1.0, 3.0, stream);
    launchFill(c.p, n, DataType::FLOAT32, 2.0, 3.0, stream);
    void* all[3] = {a.p, b.p, c.p};
    launchLocalAllreduceMany(all, 3, n, DataType::FLOAT32, ReduceOp::SUM, 1.0f, stream);
    noteLaunch();
    DevBuf res(16);
    unsigned long long init[2] = {0ull, ~0ull};
    GLB_CUDA_CHECK(cudaMemcpyAsync(res.p, init, 16, cudaMemcpyHostToDevice, stream));
    launchVerify(c.p, n, DataType::FLOAT32, 3.0, 9.0, 1e-5, 1e-6, static_cast<unsigned long long*>(res.p), stream);
    unsigned long long got[2];
    GLB_CUDA_CHECK(cudaMemcpyAsync(got, res.p, 16, cudaMemcpyDeviceToHost, stream));
    GLB_CUDA_CHECK(cudaStreamSynchronize(stream));
    return got[0] == 0 ? "" : strcat_all(got[0], " mismatches, first at ", got[1] - 1);
  });

  GLB_CUDA_CHECK(cudaStreamSynchronize(stream));
  pc.checkHealth();
  return R.results;
}

// A barrier whose only peer never shows up: the kernel must give up after the timeout,
// leave the GPU usable and poison the context (the caller sees IoException).
bool loopbackTimeoutTest(PeerContext& pc, cudaStream_t stream, int timeoutMs, double* elapsedMs) {
  DeviceGuard g(pc.device);
  CommArgs ca = pc.loopbackComm(2);
  // Virtual rank 1's pad lives elsewhere: my flag for it lands there, its flag for me never comes.
  DevBuf lonely(sizeof(SignalPad));
  GLB_CUDA_CHECK(cudaMemsetAsync(lonely.p, 0, sizeof(SignalPad), stream));
  ca.sig[1] = static_cast<SignalPad*>(lonely.p);  // (ca.self stays my real pad)
  ca.timeoutNs = static_cast<unsigned long long>(timeoutMs) * 1000000ull;
  cudaEvent_t a, b;
  GLB_CUDA_CHECK(cudaEventCreate(&a));
  GLB_CUDA_CHECK(cudaEventCreate(&b));
  GLB_CUDA_CHECK(cudaEventRecord(a, stream));
  launchBarrier(ca, stream);
  GLB_CUDA_CHECK(cudaEventRecord(b, stream));
  GLB_CUDA_CHECK(cudaStreamSynchronize(stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  if (elapsedMs != nullptr) *elapsedMs = ms;
  try {
    pc.checkHealth();
  } catch (const IoException&) {
    return true;
  }
  return false;
}

// LocalOp classes (cuda/local_ops.h) on one device: memcpy, native / host reduce and
// broadcast, and the dispatchers. The NCCL flavours need one GPU per pointer and are
// exercised when `devices` names several.
std::vector<SelfTestResult> localOpsSelfTest(const std::vector<int>& devices, size_t count) {
  std::vector<SelfTestResult> out;
  auto run = [&](const std::string& name, auto&& body) {
    SelfTestResult r;
    r.name = name;
    try {
      r.detail = body();
      r.ok = r.detail.empty();
    } catch (const std::exception& e) {
      r.detail = e.what();
    }
    out.push_back(r);
  };
  const int n = static_cast<int>(devices.size());
  GLB_ENFORCE_GE(n, 1);
  std::vector<CudaDevicePointer<float>> ptrs;
  std::vector<CudaStream> streams;
  const int nptr = n >= 2 ? n : 3;  // several devices: one pointer each (what the NCCL flavours need)
  for (int i = 0; i < nptr; i++) {
    const int dev = devices[i % n];
    DeviceGuard g(dev);
    ptrs.push_back(CudaDevicePointer<float>::alloc(count));
    streams.emplace_back(dev);
  }
  auto fill = [&] {
    for (int i = 0; i < nptr; i++) {
      DeviceGuard g(ptrs[i].getDeviceID());
      launchFill(*ptrs[i], count, DataType::FLOAT32, i, nptr, *streams[i]);
    }
  };
  const double tri = nptr * (nptr - 1) / 2.0;
  auto check = [&](const float* host, double start, double stride) -> std::string {
    for (size_t i = 0; i < count; i++) {
      if (!close(host[i], start + stride * static_cast<double>(i))) return strcat_all("element ", i, ": got ", host[i]);
    }
    return "";
  };
  auto host = CudaHostPointer<float>::alloc(count);

  run("CudaLocalMemcpy (device -> host)", [&]() -> std::string {
    fill();
    CudaLocalMemcpy<float, CudaDevicePointer<float>, CudaHostPointer<float>> cp(streams[0], ptrs[0], host, 0, count);
    cp.run();
    return check(*host, 0, nptr);
  });
  run("cudaDeviceReduce + cudaDeviceBroadcast", [&]() -> std::string {
    fill();
    auto red = cudaDeviceReduce<float>(streams, ptrs, ptrs[0], CudaReductionFunction<float>::sum, 0, count);
    red->run();
    auto bc = cudaDeviceBroadcast<float>(streams, ptrs, ptrs[0], 0, count);
    bc->run();
    CudaLocalMemcpy<float, CudaDevicePointer<float>, CudaHostPointer<float>> cp(streams[nptr - 1], ptrs[nptr - 1], host, 0, count);
    cp.run();
    return check(*host, tri, static_cast<double>(nptr) * nptr);
  });
  run("cudaHostReduce (device fold) + cudaHostBroadcast", [&]() -> std::string {
    fill();
    auto red = cudaHostReduce<float>(streams, ptrs, host, CudaReductionFunction<float>::sum, 0, count);
    red->run();
    std::string why = check(*host, tri, static_cast<double>(nptr) * nptr);
    if (!why.empty()) return why;
    auto bc = cudaHostBroadcast<float>(streams, ptrs, host, 0, count);
    bc->run();
    auto back = CudaHostPointer<float>::alloc(count);
    CudaLocalMemcpy<float, CudaDevicePointer<float>, CudaHostPointer<float>> cp(streams[1], ptrs[1], back, 0, count);
    cp.run();
    return check(*back, tri, static_cast<double>(nptr) * nptr);
  });
  run("CudaLocalHostReduce (CPU fold)", [&]() -> std::string {
    fill();
    CudaLocalHostReduce<float> red(streams, ptrs, host, CudaReductionFunction<float>::sum, 0, count);
    red.run();
    return check(*host, tri, static_cast<double>(nptr) * nptr);
  });
  if (n >= 2 && n == nptr && ncclAvailable()) {
    run("CudaLocalNCCLReduce + CudaLocalNCCLBroadcast", [&]() -> std::string {
      fill();
      CudaLocalNCCLReduce<float> red(streams, ptrs, ptrs[0], CudaReductionFunction<float>::sum, 0, count);
      red.run();
      CudaLocalNCCLBroadcast<float> bc(streams, ptrs, ptrs[0], 0, count);
      bc.run();
      CudaLocalMemcpy<float, CudaDevicePointer<float>, CudaHostPointer<float>> cp(streams[n - 1], ptrs[n - 1], host, 0, count);
      cp.run();
      return check(*host, tri, static_cast<double>(nptr) * nptr);
    });
  }
  return out;
}

}  // namespace cuda
}  // namespace glb
