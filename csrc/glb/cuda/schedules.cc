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

namespace {
// Non-power-of-two P, the simple way: the P - 2^k surplus ranks fold their vector onto a
// partner first and pull the result back afterwards (GLB_HD_FOLD=1).
Schedule buildHalvingDoublingFoldSchedule(int rank, int size, size_t count, size_t packElems) {
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

// Non-power-of-two P, binary blocks (the reference's scheme, allreduce_halving_doubling.h:39-64
// and cuda_allreduce_halving_doubling.cc; host version: glb/binary_blocks.h), as a pull table:
//   1. halving reduce-scatter inside every block, all blocks at once (lg b0 steps; a smaller
//      block idles through the steps it does not have);
//   2. chain up, smallest block first: a rank of the larger block folds in its range from the
//      rank of the smaller block that owns the enclosing range;
//   3. chain down: a rank of the smaller block copies its range back piece by piece from the
//      b_large / b_small ranks of the larger block that hold it (one barrier phase);
//   4. doubling allgather inside every block (a smaller block's steps are the last ones).
// Ranges nest because local rank l keeps half (l >> i) & 1 in step i: the first lg b' bits of
// l give the range of local rank l mod b' in a block of b' ranks.
Schedule buildHalvingDoublingBlocksSchedule(int rank, int size, size_t count, size_t packElems) {
  Schedule s;
  s.name = "halving_doubling";
  const Range all{0, count}, none{0, 0};
  std::vector<int> bsize, bbase;
  int block = 0, me = 0;
  for (int bit = 30, start = 0; bit >= 0; bit--) {
    const int b = 1 << bit;
    if (!(size & b)) continue;
    if (rank >= start && rank < start + b) {
      block = static_cast<int>(bsize.size());
      me = rank - start;
    }
    bsize.push_back(b);
    bbase.push_back(start);
    start += b;
  }
  const int nb = static_cast<int>(bsize.size());
  const int b = bsize[block], base = bbase[block];
  const int S = static_cast<int>(log2ceil(static_cast<uint32_t>(bsize[0])));
  auto rangeOf = [&](int bsz, int l) {
    Range r = all;
    for (int d = 1; d < bsz; d <<= 1) r = alignedPart(r, 2, (l & d) ? 1 : 0, packElems);
    return r;
  };
  // 1. reduce-scatter
  std::vector<Range> owned{all};
  for (int i = 0; i < S; i++) {
    const int d = 1 << i;
    if (d < b) {
      owned.push_back(alignedPart(owned.back(), 2, (me & d) ? 1 : 0, packElems));
      s.steps.push_back(makeStep(SCHED_REDUCE, owned.back(), {base + (me ^ d)}));
    } else {
      s.steps.push_back(makeStep(SCHED_REDUCE, none, {}));
    }
  }
  const Range mine = owned.back();
  // 2. chain up
  for (int j = 0; j + 1 < nb; j++) {
    const int dst = nb - 2 - j, src = dst + 1;
    if (block == dst) {
      s.steps.push_back(makeStep(SCHED_REDUCE, mine, {bbase[src] + me % bsize[src]}));
    } else {
      s.steps.push_back(makeStep(SCHED_REDUCE, none, {}));
    }
  }
  // 3. chain down (every rank emits the same number of table entries per phase)
  for (int j = 0; j + 1 < nb; j++) {
    const int src = j, dst = j + 1;
    const int pieces = bsize[src] / bsize[dst];
    for (int m = 0; m < pieces; m++) {
      SchedStep st = block == dst ? makeStep(SCHED_COPY, rangeOf(bsize[src], me + m * b), {bbase[src] + me + m * b})
                                  : makeStep(SCHED_COPY, none, {});
      st.sync = m == 0 ? 1 : 0;
      s.steps.push_back(st);
    }
  }
  // 4. allgather
  int level = static_cast<int>(owned.size()) - 1;
  for (int i = S - 1; i >= 0; i--) {
    const int d = 1 << i;
    if (d < b) {
      s.steps.push_back(makeStep(SCHED_COPY, alignedPart(owned[level - 1], 2, (me & d) ? 0 : 1, packElems), {base + (me ^ d)}));
      level--;
    } else {
      s.steps.push_back(makeStep(SCHED_COPY, none, {}));
    }
  }
  return s;
}
}  // namespace

Schedule buildHalvingDoublingSchedule(int rank, int size, size_t count, size_t packElems) {
  const int core = detail::largestPow2AtMost(size);
  if (core != size) {
    return envFlag("HD_FOLD", false) ? buildHalvingDoublingFoldSchedule(rank, size, count, packElems)
                                     : buildHalvingDoublingBlocksSchedule(rank, size, count, packElems);
  }
  Schedule s;
  s.name = "halving_doubling";
  std::vector<int> factors(log2ceil(static_cast<uint32_t>(core)), 2);
  appendHypercube(s, rank, factors, Range{0, count}, packElems);
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
