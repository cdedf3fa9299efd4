"""Old-style (construct once, run many) host algorithms — numpy / CPU torch buffers.

Each class mirrors its namesake in the reference (gloo/allreduce_ring.h, ...):
buffers and slots are bound at construction, run() is reusable.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

from .. import _C
from ..types import ReduceOp, describe


def _ptrs(bufs):
    bufs = list(bufs) if isinstance(bufs, (list, tuple)) else [bufs]
    ptr0, n, dt, cuda = describe(bufs[0])
    assert not cuda, "host algorithm constructed with CUDA buffers"
    return bufs, [describe(b)[0] for b in bufs], n, dt


class _Algo:
    NAME = ""

    def __init__(self, ctx, bufs, op: ReduceOp = ReduceOp.SUM, **kw):
        self.bufs, ptrs, n, dt = _ptrs(bufs)
        self._impl = _C.make_algorithm(self.NAME, ctx, ptrs, n, int(dt), int(op), **kw)

    def run(self):
        self._impl.run()


class AllreduceRing(_Algo):
    NAME = "allreduce_ring"


class AllreduceRingChunked(_Algo):
    NAME = "allreduce_ring_chunked"


class AllreduceHalvingDoubling(_Algo):
    NAME = "allreduce_halving_doubling"


class AllreduceBcube(_Algo):
    NAME = "allreduce_bcube"


class AllreduceLocal(_Algo):
    NAME = "allreduce_local"


class BroadcastOneToAll(_Algo):
    NAME = "broadcast_one_to_all"

    def __init__(self, ctx, bufs, root: int = 0, root_pointer: int = 0):
        super().__init__(ctx, bufs, root=root, root_pointer=root_pointer)


class ReduceScatterHalvingDoubling(_Algo):
    NAME = "reduce_scatter_halving_doubling"

    def __init__(self, ctx, bufs, recv_elems: Sequence[int], op: ReduceOp = ReduceOp.SUM):
        super().__init__(ctx, bufs, op, recv_elems=list(recv_elems))


class AllgatherRing:
    def __init__(self, ctx, inputs, output):
        self.inputs, ptrs, n, dt = _ptrs(inputs)
        self.output = output
        self._impl = _C.make_algorithm("allgather_ring", ctx, ptrs, n, int(dt), out_ptr=describe(output)[0])

    def run(self):
        self._impl.run()


class BarrierAllToAll:
    def __init__(self, ctx):
        self._impl = _C.make_algorithm("barrier_all_to_all", ctx)

    def run(self):
        self._impl.run()


class BarrierAllToOne:
    def __init__(self, ctx, root: int = 0):
        self._impl = _C.make_algorithm("barrier_all_to_one", ctx, root=root)

    def run(self):
        self._impl.run()


class PairwiseExchange:
    def __init__(self, ctx, num_bytes: int, num_destinations: int):
        self._impl = _C.make_algorithm("pairwise_exchange", ctx, count=num_bytes, root=num_destinations)

    def run(self):
        self._impl.run()
