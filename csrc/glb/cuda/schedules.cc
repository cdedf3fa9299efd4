#include "glb/cuda/schedules.h"

#include <cstring>

#include "glb/collectives_common.h"
#include "glb/common/utils.h"
#include "glb/mixed_radix.h"

namespace glb {
namespace cuda {

using detail::Range;
using detail::subRange;

namespace {
SchedStep makeStep(int mode, Range r, std::initializer_list<int> peers, int fromStage = 0) {
  SchedStep s;
  std::memset(&s, 0, sizeof(s));
  s.mode = mode;
  s.off = r.off;
  s.len = r.len;
  s.fromStage = fromStage;
  s.sync = 1;
  for (int p : peers) s.peers[s.npeers++] = p;
  return s;
}

// Element ranges are aligned down to 16-byte packs so every step stays vectorised;
// the last piece absorbs the remainder.
Range alignedPart(Range whole, size_t parts, size_t i, size_t packElems) {
  if (packElems <= 1) return subRange(whole, parts, i);
  const size_t packs = whole.len / packElems;
  Range p = subRange(Range{0, packs}, parts, i);
  Range out{whole.off + p.off * packElems, p.len * packElems};
  if (i == parts - 1) out.len = whole.len - p.off * packElems;
  return out;
}
}  // namespace

// Ring: P-1 rounds, each moving the WHOLE vector one hop (P·S bytes): round k folds in
// the contribution that has travelled k hops — i.e. the original vector of rank r-k,
// which that rank published in its pool before the first round.
Schedule buildRingSchedule(int rank, int size, size_t count, size_t packElems) {
  (void)packElems;
  Schedule s;
  s.name = "ring";
  s.needsStage = true;
  const Range all{0, count};
  s.steps.push_back(makeStep(SCHED_STAGE, all, {}));
  for (int k = 1; k < size; k++) {
    s.steps.push_back(makeStep(SCHED_REDUCE, all, {(rank - k + size) % size}, 1));
  }
  return s;
}

// Ring chunked: P chunks; P-1 reduce-scatter rounds then P-1 allgather rounds, each
// pulling one chunk from the left neighbour (2·S·(P-1)/P bytes, 2(P-1) steps).
Schedule buildRingChunkedSchedule(int rank, int size, size_t count, size_t packElems) {
  Schedule s;
  s.name = "ring_chunked";
  const Range all{0, count};
  const int left = (rank - 1 + size) % size;
  auto chunk = [&](int c) { return alignedPart(all, size, ((c % size) + size) % size, packElems); };
  for (int k = 0; k < size - 1; k++) s.steps.push_back(makeStep(SCHED_REDUCE, chunk(rank - k - 1), {left}));
  // After the reduce-scatter rank r owns chunk r+1... with pulls the owner of the
  // fully reduced chunk c is rank c-1+... : round k of the allgather copies chunk r-k.
  for (int k = 0; k < size - 1; k++) s.steps.push_back(makeStep(SCHED_COPY, chunk(rank - k), {left}));
  return s;
}

namespace {
// Mixed-radix hypercube with pulls: reduce-scatter then allgather.
void appendHypercube(Schedule& s, int rank, const std::vector<int>& factors, Range all, size_t packElems) {
  const int K = static_cast<int>(factors.size());
  std::vector<int> stride(K), digit(K);
  int st = 1;
  for (int i = 0; i < K; i++) {
    stride[i] = st;
    digit[i] = (rank / st) % factors[i];
    st *= factors[i];
  }
  std::vector<Range> blocks(K + 1);
  blocks[0] = all;
  for (int i = 0; i < K; i++) blocks[i + 1] = alignedPart(blocks[i], factors[i], digit[i], packElems);
  for (int i = 0; i < K; i++) {
    SchedStep step = makeStep(SCHED_REDUCE, blocks[i + 1], {});
    for (int d = 0; d < factors[i]; d++) {
      if (d != digit[i]) step.peers[step.npeers++] = rank + (d - digit[i]) * stride[i];
    }
    s.steps.push_back(step);
  }
  for (int i = K - 1; i >= 0; i--) {
    // One copy step per peer: each peer owns a different sub-block of blocks[i].
    for (int d = 0; d < factors[i]; d++) {
      if (d == digit[i]) continue;
      const int peer = rank + (d - digit[i]) * stride[i];
      s.steps.push_back(makeStep(SCHED_COPY, alignedPart(blocks[i], factors[i], d, packElems), {peer}));
    }
  }
}
}  // namespace

Schedule buildHalvingDoublingSchedule(int rank, int size, size_t count, size_t packElems) {
  Schedule s;
  s.name = "halving_doubling";
  const Range all{0, count};
  const int core = detail::largestPow2AtMost(size);
  const int extras = size - core;
  std::vector<int> factors(log2ceil(static_cast<uint32_t>(core)), 2);
  // Every rank's table must have the same number of steps (steps are barrier-separated):
  // ranks that sit a step out get an empty range.
  const Range none{0, 0};
  // fold-in: core rank e pulls the whole vector of extra rank core+e
  if (extras > 0) {
    if (rank < extras) {
      s.steps.push_back(makeStep(SCHED_REDUCE, all, {rank + core}));
    } else {
      s.steps.push_back(makeStep(SCHED_REDUCE, none, {}));
    }
  }
  if (rank < core) {
    appendHypercube(s, rank, factors, all, packElems);
  } else {
    Schedule dummy;
    appendHypercube(dummy, 0, factors, all, packElems);
    for (size_t i = 0; i < dummy.steps.size(); i++) s.steps.push_back(makeStep(SCHED_COPY, none, {}));
  }
  // fold-out: the extra pulls the finished vector back
  if (extras > 0) {
    if (rank >= core) {
      s.steps.push_back(makeStep(SCHED_COPY, all, {rank - core}));
    } else {
      s.steps.push_back(makeStep(SCHED_COPY, none, {}));
    }
  }
  return s;
}

Schedule buildHalvingDoublingPipelinedSchedule(int rank, int size, size_t count, size_t packElems, int chunks) {
  Schedule s;
  s.name = "halving_doubling_pipelined";
  chunks = std::max(1, chunks);
  // Too small to cut: every chunk must keep at least one pack per rank.
  while (chunks > 1 && count / static_cast<size_t>(chunks) < packElems * static_cast<size_t>(size)) chunks--;
  std::vector<Schedule> per(chunks);
  size_t depth = 0;
  for (int c = 0; c < chunks; c++) {
    const Range r = alignedPart(Range{0, count}, chunks, c, packElems);
    per[c] = buildHalvingDoublingSchedule(rank, size, r.len, packElems);
    for (auto& st : per[c].steps) st.off += r.off;
    depth = std::max(depth, per[c].steps.size());
  }
  // Phase t runs step t - c of chunk c for every chunk that has one.
  for (size_t t = 0; t < depth + chunks - 1; t++) {
    bool first = true;
    for (int c = 0; c < chunks; c++) {
      if (t < static_cast<size_t>(c) || t - c >= per[c].steps.size()) continue;
      SchedStep st = per[c].steps[t - c];
      st.sync = first ? 1 : 0;
      first = false;
      s.steps.push_back(st);
    }
  }
  return s;
}

int scheduleBarriers(const Schedule& s) {
  int n = 1;
  for (const auto& st : s.steps) n += st.sync ? 1 : 0;
  return n;
}

Schedule buildBcubeSchedule(int rank, int size, size_t count, int base, size_t packElems) {
  Schedule s;
  s.name = "bcube";
  // Allgather steps differ in number per rank only through the radix list, which is
  // the same everywhere, so tables line up.
  appendHypercube(s, rank, detail::radixFactors(size, std::max(2, base)), Range{0, count}, packElems);
  return s;
}

}  // namespace cuda
}  // namespace glb
