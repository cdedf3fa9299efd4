// Halving-doubling allreduce for a rank count that is NOT a power of two: binary blocks.
//
// P is written as a sum of decreasing powers of two (7 = 4 + 2 + 1) and the ranks are cut
// into blocks of those sizes. Every block runs recursive vector halving / distance doubling
// on the whole vector by itself; then the blocks are chained:
//
//   1. reduce-scatter inside every block           (lg b steps, all blocks concurrently)
//   2. smallest block -> next larger -> ... : a rank of the smaller block owns a range that
//      is the union of the ranges of k = b_large / b_small ranks of the larger block, and
//      sends each of them its piece to fold in; the largest block ends up with the full
//      reduction, scattered over its ranks
//   3. the same chain backwards: every rank of a larger block returns its finished range to
//      the rank of the smaller block that owns the enclosing range
//   4. allgather inside every block                (lg b steps)
//
// Compared with folding the P - 2^k surplus ranks onto partners (one extra full-vector hop
// before and after, during which everybody else idles) no rank is ever idle and the extra
// traffic of a small block is its own share S / b_small, not S. This is the scheme the
// reference uses (gloo/allreduce_halving_doubling.h:39-64, 159-222, 262-317); the
// implementation here is built on unbound buffers (eager sends, receives posted per step).
#pragma once

#include <cstring>
#include <memory>
#include <vector>

#include "glb/algorithm.h"
#include "glb/collectives_common.h"
#include "glb/common/utils.h"
#include "glb/transport/unbound_buffer.h"
#include "glb/types.h"

namespace glb {
namespace detail {

struct BlockLayout {
  std::vector<int> size;   // decreasing powers of two, sum = P
  std::vector<int> base;   // first rank of each block
  int block = 0;           // my block
  int local = 0;           // my rank inside it
};

inline BlockLayout binaryBlocks(int P, int rank) {
  BlockLayout l;
  int start = 0;
  for (int bit = 30; bit >= 0; bit--) {
    const int b = 1 << bit;
    if (P & b) {
      if (rank >= start && rank < start + b) {
        l.block = static_cast<int>(l.size.size());
        l.local = rank - start;
      }
      l.size.push_back(b);
      l.base.push_back(start);
      start += b;
    }
  }
  return l;
}

// Range owned by local rank `l` of a block of `b` ranks after the halving reduce-scatter:
// bit i of l (LSB first) picks the half kept in step i. The first lg(b') bits give the range
// of local rank l mod b' in a block of b' < b ranks, which is what makes blocks nest.
inline Range blockRange(size_t count, int b, int l) {
  Range r{0, count};
  for (int d = 1; d < b; d <<= 1) r = subRange(r, 2, (l & d) ? 1 : 0);
  return r;
}

template <typename T>
class BinaryBlocksAllreduce {
 public:
  BinaryBlocksAllreduce(const std::shared_ptr<Context>& context, T* data, size_t count, const ReductionFunction<T>* fn)
      : context_(context), data_(data), count_(count), fn_(fn), lay_(binaryBlocks(context->size, context->rank)) {
    int steps = 0;
    for (int b = lay_.size[0]; b > 1; b >>= 1) steps++;
    // reduce-scatter steps, chain up, chain down, allgather steps: one slot each
    slot_ = Slot::build(kBinaryBlocksSlotPrefix, static_cast<uint32_t>(context_->nextSlot(2 * steps + 2)));
    steps_ = steps;
    scratch_.resize(std::max<size_t>(1, (count_ + 1) / 2 + 1));
    buf_ = context_->createUnboundBuffer(data_, count_ * sizeof(T));
    tmp_ = context_->createUnboundBuffer(scratch_.data(), scratch_.size() * sizeof(T));
  }

  void run() {
    const int b = lay_.size[lay_.block];
    const int me = lay_.local;
    const int base = lay_.base[lay_.block];
    const int nblocks = static_cast<int>(lay_.size.size());
    // 1. reduce-scatter inside my block
    Range cur{0, count_};
    int step = 0;
    for (int d = 1; d < b; d <<= 1, step++) {
      const int partner = base + (me ^ d);
      const Range keep = subRange(cur, 2, (me & d) ? 1 : 0), give = subRange(cur, 2, (me & d) ? 0 : 1);
      if (keep.len > 0) tmp_->recv(partner, slot_ + step, 0, keep.len * sizeof(T));
      if (give.len > 0) buf_->send(partner, slot_ + step, give.off * sizeof(T), give.len * sizeof(T));
      if (keep.len > 0) {
        tmp_->waitRecv();
        fn_->call(data_ + keep.off, scratch_.data(), keep.len);
      }
      if (give.len > 0) buf_->waitSend();
      cur = keep;
    }
    const uint64_t up = slot_ + steps_, down = slot_ + steps_ + 1;
    // 2. chain up: fold in what the next smaller block sends, then pass my range on
    if (lay_.block + 1 < nblocks) {
      const int small = lay_.size[lay_.block + 1];
      const int from = lay_.base[lay_.block + 1] + me % small;
      if (cur.len > 0) {
        tmp_->recv(from, up, 0, cur.len * sizeof(T));
        tmp_->waitRecv();
        fn_->call(data_ + cur.off, scratch_.data(), cur.len);
      }
    }
    if (lay_.block > 0) {
      const int big = lay_.size[lay_.block - 1];
      int sends = 0;
      for (int m = 0; m < big / b; m++) {
        const Range piece = blockRange(count_, big, me + m * b);
        if (piece.len == 0) continue;
        buf_->send(lay_.base[lay_.block - 1] + me + m * b, up, piece.off * sizeof(T), piece.len * sizeof(T));
        sends++;
      }
      // 3. chain down: the finished pieces come back into place
      int recvs = 0;
      std::vector<std::unique_ptr<transport::UnboundBuffer>> landing;
      for (int m = 0; m < big / b; m++) {
        const Range piece = blockRange(count_, big, me + m * b);
        if (piece.len == 0) continue;
        landing.push_back(context_->createUnboundBuffer(data_ + piece.off, piece.len * sizeof(T)));
        landing.back()->recv(lay_.base[lay_.block - 1] + me + m * b, down);
        recvs++;
      }
      for (int i = 0; i < sends; i++) buf_->waitSend();
      for (auto& l : landing) l->waitRecv();
      (void)recvs;
    }
    if (lay_.block + 1 < nblocks && cur.len > 0) {
      const int small = lay_.size[lay_.block + 1];
      buf_->send(lay_.base[lay_.block + 1] + me % small, down, cur.off * sizeof(T), cur.len * sizeof(T));
      buf_->waitSend();
    }
    // 4. allgather inside my block (the halving steps backwards)
    std::vector<Range> owned;  // my range after each halving step
    {
      Range r{0, count_};
      owned.push_back(r);
      for (int d = 1; d < b; d <<= 1) {
        r = subRange(r, 2, (me & d) ? 1 : 0);
        owned.push_back(r);
      }
    }
    step = steps_ + 2;
    int level = static_cast<int>(owned.size()) - 1;
    for (int d = b >> 1; d >= 1; d >>= 1, step++, level--) {
      const int partner = base + (me ^ d);
      const Range mine = owned[level];
      const Range theirs = subRange(owned[level - 1], 2, (me & d) ? 0 : 1);
      std::unique_ptr<transport::UnboundBuffer> land;
      if (theirs.len > 0) {
        land = context_->createUnboundBuffer(data_ + theirs.off, theirs.len * sizeof(T));
        land->recv(partner, slot_ + step);
      }
      if (mine.len > 0) buf_->send(partner, slot_ + step, mine.off * sizeof(T), mine.len * sizeof(T));
      if (land) land->waitRecv();
      if (mine.len > 0) buf_->waitSend();
    }
  }

 private:
  static constexpr uint8_t kBinaryBlocksSlotPrefix = 0x42;
  std::shared_ptr<Context> context_;
  T* data_;
  const size_t count_;
  const ReductionFunction<T>* fn_;
  BlockLayout lay_;
  uint64_t slot_ = 0;
  int steps_ = 0;
  std::vector<T> scratch_;
  std::unique_ptr<transport::UnboundBuffer> buf_, tmp_;
};

}  // namespace detail
}  // namespace glb
