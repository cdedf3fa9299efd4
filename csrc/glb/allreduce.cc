#include "glb/common/trace.h"
#include "glb/allreduce.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "glb/common/utils.h"

namespace glb {

using detail::Range;
using detail::subRange;

void AllreduceOptions::setInputsRaw(const std::vector<void*>& ptrs, size_t n, size_t elemSize) {
  elements = n;
  elementSize = elemSize;
  in.clear();
  for (void* p : ptrs) in.push_back(context->createUnboundBuffer(p, n * elemSize));
}

void AllreduceOptions::setOutputsRaw(const std::vector<void*>& ptrs, size_t n, size_t elemSize) {
  elements = n;
  elementSize = elemSize;
  out.clear();
  for (void* p : ptrs) out.push_back(context->createUnboundBuffer(p, n * elemSize));
}

namespace detail {

size_t oneHopMaxBytes() {
  static const size_t n = [] {
    const char* v = std::getenv("GLB_ONEHOP_MAX");
    if (v == nullptr) v = std::getenv("GLB_ALLREDUCE_ONESHOT_MAX");
    long long b = v != nullptr ? std::atoll(v) : (16 << 10);
    return static_cast<size_t>(b < 0 ? 0 : b);
  }();
  return n;
}

std::vector<int> factorize(int n) {
  std::vector<int> f;
  for (int p = 2; p * p <= n; p++) {
    while (n % p == 0) {
      f.push_back(p);
      n /= p;
    }
  }
  if (n > 1) f.push_back(n);
  return f;
}

void ringReduceScatter(const std::shared_ptr<Context>& context, UnboundBuffer* buf, size_t elements,
                       size_t elementSize, const AllreduceOptions::Func& reduce, size_t maxSegmentSize,
                       uint64_t slot, std::chrono::milliseconds timeout) {
  // In-place ring reduce-scatter, streamed: a segment is forwarded as soon as it has been
  // reduced, so consecutive steps overlap (same scheme as the first half of ring()).
  // Afterwards rank r owns the reduced chunk (r + 1) % P.
  const int P = context->size;
  const int r = context->rank;
  if (P == 1 || elements == 0) return;
  const int right = (r + 1) % P;
  const int left = (r - 1 + P) % P;
  char* base = static_cast<char*>(buf->ptr);
  const Range all{0, elements};

  const size_t maxChunkElems = ceilDiv(elements, static_cast<size_t>(P));
  size_t segBytes = std::min(maxSegmentSize, std::max<size_t>(256u << 10, maxChunkElems * elementSize / 4));
  const size_t segElems = std::max<size_t>(1, std::min(maxChunkElems, segBytes / elementSize));
  const size_t nseg = ceilDiv(maxChunkElems, segElems);
  auto chunk = [&](int c) { return subRange(all, P, static_cast<size_t>(((c % P) + P) % P)); };
  auto segOf = [&](const Range& ch, size_t j) {
    Range s;
    s.off = ch.off + std::min(ch.len, j * segElems);
    s.len = std::min(segElems, ch.len - std::min(ch.len, j * segElems));
    return s;
  };

  constexpr size_t kDepth = 4;
  std::vector<char> tmpStorage(kDepth * segElems * elementSize);
  auto tmp = context->createUnboundBuffer(tmpStorage.data(), tmpStorage.size());
  const size_t total = static_cast<size_t>(P - 1) * nseg;
  auto post = [&](size_t k) {
    Range sg = segOf(chunk(r - static_cast<int>(k / nseg) - 1), k % nseg);
    tmp->recv(left, slot, (k % kDepth) * segElems * elementSize, sg.len * elementSize);
  };
  for (size_t k = 0; k < std::min(kDepth, total); k++) post(k);
  size_t sends = 0;
  for (size_t j = 0; j < nseg; j++) {
    Range sg = segOf(chunk(r), j);
    buf->send(right, slot, sg.off * elementSize, sg.len * elementSize);
    sends++;
  }
  for (size_t k = 0; k < total; k++) {
    const size_t s = k / nseg, j = k % nseg;
    tmp->waitRecv(timeout);
    Range sg = segOf(chunk(r - static_cast<int>(s) - 1), j);
    if (sg.len > 0) {
      char* dst = base + sg.off * elementSize;
      reduce(dst, dst, tmpStorage.data() + (k % kDepth) * segElems * elementSize, sg.len);
    }
    if (k + kDepth < total) post(k + kDepth);
    if (s + 2 < static_cast<size_t>(P)) {
      buf->send(right, slot, sg.off * elementSize, sg.len * elementSize);
      sends++;
    }
  }
  for (size_t k = 0; k < sends; k++) buf->waitSend(timeout);
}

}  // namespace detail

namespace {

// Streaming ring. Reduce-scatter and allgather form one pipeline of segments: a segment is
// forwarded the moment it has been reduced (or, in the allgather half, received) instead
// of after the whole step, so the cost is (2P-3) segment latencies plus one pass over the
// vector rather than 2(P-1) chunk latencies. `src` is where this rank's contribution is
// read from: with a single out-of-place input the first touch of every chunk reads the
// input directly (step-0 sends come from it, reductions compute out = in + received), so
// the up-front copy of the whole vector disappears.
void ring(const AllreduceOptions& opts, UnboundBuffer* out0, UnboundBuffer* src) {
  const auto& context = opts.context;
  const int P = context->size;
  const int r = context->rank;
  const int right = (r + 1) % P;
  const int left = (r - 1 + P) % P;
  const size_t es = opts.elementSize;
  const auto rsSlot = Slot::build(kAllreduceSlotPrefix, opts.tag);
  const auto agSlot = rsSlot + 1;
  const auto timeout = opts.timeout;
  const Range all{0, opts.elements};
  char* outBase = static_cast<char*>(out0->ptr);
  const char* srcBase = static_cast<const char*>(src->ptr);

  const size_t maxChunkElems = ceilDiv(opts.elements, static_cast<size_t>(P));
  // maxSegmentSize is an upper bound. Below it, aim for four segments per chunk so that the
  // arrival of one segment overlaps the reduction of the previous one, but not smaller
  // than 256 KiB (per-message cost, and the same-host single-copy threshold).
  size_t segBytes = std::min(opts.maxSegmentSize, std::max<size_t>(256u << 10, maxChunkElems * es / 4));
  const size_t segElems = std::max<size_t>(1, std::min(maxChunkElems, segBytes / es));
  const size_t nseg = ceilDiv(maxChunkElems, segElems);
  auto chunk = [&](int c) { return subRange(all, P, static_cast<size_t>(((c % P) + P) % P)); };
  auto segOf = [&](const Range& ch, size_t j) {
    Range sg;
    sg.off = ch.off + std::min(ch.len, j * segElems);
    sg.len = std::min(segElems, ch.len - std::min(ch.len, j * segElems));
    return sg;
  };

  // Landing zones for the reduce-scatter half; a zone is re-posted as soon as it is reduced.
  constexpr size_t kDepth = 4;
  std::vector<char> tmpStorage(kDepth * segElems * es);
  auto tmp = context->createUnboundBuffer(tmpStorage.data(), tmpStorage.size());
  const size_t total = static_cast<size_t>(P - 1) * nseg;
  auto postRs = [&](size_t k) {
    const size_t s = k / nseg, j = k % nseg;
    Range sg = segOf(chunk(r - static_cast<int>(s) - 1), j);
    tmp->recv(left, rsSlot, (k % kDepth) * segElems * es, sg.len * es);
  };
  for (size_t k = 0; k < std::min(kDepth, total); k++) postRs(k);

  // Sends of the allgather half go through their own handle on the output memory, so
  // that "the reduce-scatter sends are done" can be waited for separately: a rank must
  // not wait for its allgather sends before posting its allgather receives (the peer's
  // receive may only be posted after that peer has seen *its* sends complete).
  auto agOut = context->createUnboundBuffer(out0->ptr, out0->size);
  size_t srcSends = 0, outSends = 0, agSends = 0;
  for (size_t j = 0; j < nseg; j++) {  // step 0: this rank's own chunk, unreduced
    Range sg = segOf(chunk(r), j);
    src->send(right, rsSlot, sg.off * es, sg.len * es);
    srcSends++;
  }
  for (size_t k = 0; k < total; k++) {
    const size_t s = k / nseg, j = k % nseg;
    tmp->waitRecv(timeout);
    Range sg = segOf(chunk(r - static_cast<int>(s) - 1), j);
    if (sg.len > 0) {
      opts.reduce(outBase + sg.off * es, srcBase + sg.off * es, tmpStorage.data() + (k % kDepth) * segElems * es,
                  sg.len);
    }
    if (k + kDepth < total) postRs(k + kDepth);
    // Reduced: pass it on. After the last step it is final and opens the allgather half.
    if (s + 2 < static_cast<size_t>(P)) {
      out0->send(right, rsSlot, sg.off * es, sg.len * es);
      outSends++;
    } else {
      agOut->send(right, agSlot, sg.off * es, sg.len * es);
      agSends++;
    }
  }
  // The allgather half overwrites regions the reduce-scatter sends still read (in place:
  // all of them; out of place: the forwarded partial sums), so those have to be done first.
  if (src != out0) {
    for (size_t k = 0; k < srcSends; k++) src->waitSend(timeout);
  } else {
    outSends += srcSends;
  }
  for (size_t k = 0; k < outSends; k++) out0->waitSend(timeout);
  outSends = 0;

  // Allgather half: step s receives chunk (r - s) straight into the output and forwards it.
  for (size_t k = 0; k < total; k++) {
    const size_t s = k / nseg, j = k % nseg;
    Range sg = segOf(chunk(r - static_cast<int>(s)), j);
    out0->recv(left, agSlot, sg.off * es, sg.len * es);
  }
  for (size_t k = 0; k < total; k++) {
    const size_t s = k / nseg, j = k % nseg;
    out0->waitRecv(timeout);
    if (s + 2 < static_cast<size_t>(P)) {
      Range sg = segOf(chunk(r - static_cast<int>(s)), j);
      agOut->send(right, agSlot, sg.off * es, sg.len * es);
      agSends++;
    }
  }
  for (size_t k = 0; k < agSends; k++) agOut->waitSend(timeout);
}

// `src`: where this rank's contribution is read from in the first step (see ring()).
void bcube(const AllreduceOptions& opts, UnboundBuffer* out0, UnboundBuffer* src) {
  const auto& context = opts.context;
  const int P = context->size;
  const int r = context->rank;
  const auto slot = Slot::build(kAllreduceSlotPrefix, opts.tag);
  const size_t es = opts.elementSize;
  char* base = static_cast<char*>(out0->ptr);

  const std::vector<int> factors = detail::factorize(P);
  const int K = static_cast<int>(factors.size());
  // Mixed-radix digits of this rank; stride[i] = product of factors below i.
  std::vector<int> stride(K), digit(K);
  {
    int s = 1;
    for (int i = 0; i < K; i++) {
      stride[i] = s;
      digit[i] = (r / s) % factors[i];
      s *= factors[i];
    }
  }
  auto peerAt = [&](int i, int d) { return r + (d - digit[i]) * stride[i]; };

  std::vector<Range> blocks(K + 1);
  blocks[0] = Range{0, opts.elements};
  std::vector<char> tmpStorage;

  // Reduce-scatter: after step i this rank owns blocks[i + 1].
  for (int i = 0; i < K; i++) {
    const int f = factors[i];
    const Range cur = blocks[i];
    const Range mine = subRange(cur, f, digit[i]);
    blocks[i + 1] = mine;
    tmpStorage.resize(std::max<size_t>(1, static_cast<size_t>(f - 1) * mine.len * es));
    auto tmp = context->createUnboundBuffer(tmpStorage.data(), tmpStorage.size());
    int k = 0;
    for (int d = 0; d < f; d++) {
      if (d == digit[i]) continue;
      tmp->recv(peerAt(i, d), slot + i, k * mine.len * es, mine.len * es);
      k++;
    }
    // Step 0 reads the caller's input where it lies; from then on everything this rank
    // still needs is inside the block it owns in the output.
    UnboundBuffer* from = i == 0 ? src : out0;
    for (int d = 0; d < f; d++) {
      if (d == digit[i]) continue;
      const Range theirs = subRange(cur, f, d);
      from->send(peerAt(i, d), slot + i, theirs.off * es, theirs.len * es);
    }
    // Reduce contributions in arrival order.
    for (int n = 0; n < f - 1; n++) {
      int srcRank = -1;
      tmp->waitRecv(&srcRank, opts.timeout);
      int d = digit[i] + (srcRank - r) / stride[i];
      int idx = d < digit[i] ? d : d - 1;
      if (mine.len > 0) {
        char* dst = base + mine.off * es;
        const char* acc = (i == 0 && n == 0) ? static_cast<const char*>(src->ptr) + mine.off * es : dst;
        opts.reduce(dst, acc, tmpStorage.data() + idx * mine.len * es, mine.len);
      }
    }
    for (int n = 0; n < f - 1; n++) from->waitSend(opts.timeout);
  }

  // Allgather: mirror image, results land directly in the output.
  for (int i = K - 1; i >= 0; i--) {
    const int f = factors[i];
    const Range cur = blocks[i];
    const Range mine = blocks[i + 1];
    for (int d = 0; d < f; d++) {
      if (d == digit[i]) continue;
      const Range theirs = subRange(cur, f, d);
      out0->recv(peerAt(i, d), slot + K + i, theirs.off * es, theirs.len * es);
    }
    for (int d = 0; d < f; d++) {
      if (d == digit[i]) continue;
      out0->send(peerAt(i, d), slot + K + i, mine.off * es, mine.len * es);
    }
    for (int n = 0; n < f - 1; n++) out0->waitRecv(opts.timeout);
    for (int n = 0; n < f - 1; n++) out0->waitSend(opts.timeout);
  }
}

// Latency path for small vectors (UNSPECIFIED only): every rank sends its whole contribution
// to every other rank and reduces the P vectors locally, in rank order so that all ranks
// produce bit-identical results. One message hop instead of the ring's 2(P-1); the same
// idea as the CUDA one-shot kernel. Traffic is (P-1) x S per rank, hence the size limit.
void oneShot(const AllreduceOptions& opts, UnboundBuffer* out0, UnboundBuffer* src) {
  const auto& context = opts.context;
  const int P = context->size;
  const int r = context->rank;
  const size_t bytes = opts.elements * opts.elementSize;
  const auto slot = Slot::build(kAllreduceSlotPrefix, opts.tag);
  // Landing zones for the peers' vectors plus a private copy of this rank's own one (the
  // accumulator may be the very buffer it lives in).
  std::vector<char> tmpStorage(static_cast<size_t>(P) * bytes);
  auto tmp = context->createUnboundBuffer(tmpStorage.data(), tmpStorage.size());
  for (int q = 0; q < P; q++) {
    if (q != r) tmp->recv(q, slot, static_cast<size_t>(q) * bytes, bytes);
  }
  for (int k = 1; k < P; k++) src->send((r + k) % P, slot, 0, bytes);  // staggered fan-out
  std::memcpy(tmpStorage.data() + static_cast<size_t>(r) * bytes, src->ptr, bytes);
  for (int q = 1; q < P; q++) tmp->waitRecv(opts.timeout);
  char* out = static_cast<char*>(out0->ptr);
  opts.reduce(out, tmpStorage.data(), tmpStorage.data() + bytes, opts.elements);
  for (int q = 2; q < P; q++) opts.reduce(out, out, tmpStorage.data() + static_cast<size_t>(q) * bytes, opts.elements);
  for (int q = 1; q < P; q++) src->waitSend(opts.timeout);
}

}  // namespace

void allreduce(const AllreduceOptions& opts) {
  GLB_HOST_TRACE("glb::allreduce");
  const auto& context = opts.context;
  GLB_ENFORCE(context != nullptr, "allreduce: no context");
  GLB_ENFORCE(!opts.out.empty(), "allreduce: at least one output is required");
  GLB_ENFORCE(opts.elementSize > 0, "allreduce: element size not set");
  GLB_ENFORCE(static_cast<bool>(opts.reduce), "allreduce: reduce function not set");
  const size_t bytes = opts.elements * opts.elementSize;
  for (const auto& b : opts.in) GLB_ENFORCE_EQ(b->size, bytes, "allreduce: input size mismatch");
  for (const auto& b : opts.out) GLB_ENFORCE_EQ(b->size, bytes, "allreduce: output size mismatch");
  if (opts.elements == 0) return;

  // Local phase 1: fold every input into out[0].
  UnboundBuffer* out0 = opts.out[0].get();
  UnboundBuffer* src = out0;
  if (context->size > 1 && opts.in.size() == 1 && opts.in[0]->ptr != out0->ptr) {
    src = opts.in[0].get();  // both algorithms read a single input where it lies: no up-front copy
  } else if (!opts.in.empty()) {
    if (opts.in[0]->ptr != out0->ptr) std::memcpy(out0->ptr, opts.in[0]->ptr, bytes);
    for (size_t i = 1; i < opts.in.size(); i++) {
      opts.reduce(out0->ptr, out0->ptr, opts.in[i]->ptr, opts.elements);
    }
  } else {
    // In-place over the outputs: they all hold contributions.
    for (size_t i = 1; i < opts.out.size(); i++) {
      opts.reduce(out0->ptr, out0->ptr, opts.out[i]->ptr, opts.elements);
    }
  }

  if (context->size > 1) {
    switch (opts.algorithm) {
      case AllreduceOptions::UNSPECIFIED:
        if (bytes <= detail::oneHopMaxBytes() && context->size <= 64) {
          oneShot(opts, out0, src);
          break;
        }
        ring(opts, out0, src);
        break;
      case AllreduceOptions::RING:
        ring(opts, out0, src);
        break;
      case AllreduceOptions::BCUBE:
        bcube(opts, out0, src);
        break;
      default:
        GLB_THROW_INVALID_OPERATION_EXCEPTION("allreduce: unknown algorithm ", opts.algorithm);
    }
  }

  // Local phase 2: replicate to the remaining outputs.
  for (size_t i = 1; i < opts.out.size(); i++) {
    if (opts.out[i]->ptr != out0->ptr) std::memcpy(opts.out[i]->ptr, out0->ptr, bytes);
  }
}

}  // namespace glb
