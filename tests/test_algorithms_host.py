"""Old-style host algorithms — mirrors gloo/test/allreduce_test.cc:143-299,
reduce_scatter_test.cc, allgather_test.cc, broadcast_test.cc, barrier_test.cc."""
import time

import numpy as np
import pytest

import gloo_b200 as gb
from gloo_b200.ops import algorithms as alg


def _fixture(rank, size, ptrs, count, dtype=np.float32):
    stride = size * ptrs
    return [(np.arange(count, dtype=np.float64) * stride + rank * ptrs + i).astype(dtype) for i in range(ptrs)]


def _expected(size, ptrs, count, dtype=np.float32):
    stride = size * ptrs
    return (np.arange(count, dtype=np.float64) * stride * stride + stride * (stride - 1) / 2).astype(dtype)


ALLREDUCE = [alg.AllreduceRing, alg.AllreduceRingChunked, alg.AllreduceHalvingDoubling, alg.AllreduceBcube]


@pytest.mark.parametrize("cls", ALLREDUCE)
@pytest.mark.parametrize("size", [1, 2, 3, 4, 5, 6, 7, 8, 9, 13, 16])
def test_allreduce_classes(cls, size):
    def fn(ctx):
        for count in (0, 4, 100, 1000, 10000):
            for ptrs in (1, 2):
                bufs = _fixture(ctx.rank, size, ptrs, count)
                a = cls(ctx, bufs)
                a.run()
                for b in bufs:
                    np.testing.assert_allclose(b, _expected(size, ptrs, count), rtol=1e-5)
                # run() is reusable: feed new data through the same instance
                for i, b in enumerate(bufs):
                    b[:] = _fixture(ctx.rank, size, ptrs, count)[i]
                a.run()
                for b in bufs:
                    np.testing.assert_allclose(b, _expected(size, ptrs, count), rtol=1e-5)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", [24, 32])
def test_halving_doubling_many_ranks(size):
    """The reference runs halving-doubling up to 24 and 32 ranks (allreduce_test.cc:143-299):
    24 = 16-rank core + 8 folded extras, 32 = a full 5-step hypercube."""
    def fn(ctx):
        for count in (4, 1000):
            bufs = _fixture(ctx.rank, size, 1, count)
            alg.AllreduceHalvingDoubling(ctx, bufs).run()
            np.testing.assert_allclose(bufs[0], _expected(size, 1, count), rtol=1e-5)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("base,size", [(2, 8), (3, 9), (3, 27), (4, 16), (4, 12), (3, 7)])
def test_bcube_bases(base, size):
    def fn(ctx):
        buf = _fixture(ctx.rank, size, 1, 1000)
        alg.AllreduceBcube(ctx, buf).run()
        np.testing.assert_allclose(buf[0], _expected(size, 1, 1000), rtol=1e-5)
        return True

    assert all(gb.spawn_threads(size, fn, base=base))


def test_multiple_algorithms_one_context():
    """Several instances on one context consume distinct slots (allreduce_test.cc:171-210)."""
    size = 4

    def fn(ctx):
        bufs = [_fixture(ctx.rank, size, 1, 500) for _ in range(3)]
        algos = [alg.AllreduceRingChunked(ctx, bufs[0]), alg.AllreduceHalvingDoubling(ctx, bufs[1]),
                 alg.AllreduceRing(ctx, bufs[2])]
        for a in reversed(algos):
            a.run()
        for b in bufs:
            np.testing.assert_allclose(b[0], _expected(size, 1, 500), rtol=1e-5)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("cls", ALLREDUCE)
def test_allreduce_half(cls):
    size = 4

    def fn(ctx):
        buf = [(np.arange(64) % 8 + ctx.rank).astype(np.float16)]
        cls(ctx, buf).run()
        exp = sum((np.arange(64) % 8 + r).astype(np.float64) for r in range(size))
        np.testing.assert_allclose(buf[0].astype(np.float64), exp, rtol=1e-3)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", [1, 2, 3, 4, 6, 8, 13])
def test_reduce_scatter_hd(size):
    def fn(ctx):
        for count in (size, 100, 1000, 10000):
            base, rem = divmod(count, size)
            recv = [base + (1 if r < rem else 0) for r in range(size)]
            buf = _fixture(ctx.rank, size, 1, count, np.float64)
            alg.ReduceScatterHalvingDoubling(ctx, buf, recv).run()
            off = int(np.sum(recv[:ctx.rank]))
            np.testing.assert_allclose(buf[0][:recv[ctx.rank]],
                                       _expected(size, 1, count, np.float64)[off:off + recv[ctx.rank]])
            # skewed split: everything to the last rank, nothing to the others
            recv2 = [0] * size
            recv2[-1] = count
            buf2 = _fixture(ctx.rank, size, 1, count, np.float64)
            alg.ReduceScatterHalvingDoubling(ctx, buf2, recv2).run()
            if ctx.rank == size - 1:
                np.testing.assert_allclose(buf2[0], _expected(size, 1, count, np.float64))
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", [2, 3, 6])
def test_reduce_scatter_hd_rerun_flow_control(size):
    """One instance, many runs, fresh data every run, ranks deliberately out of step.
    Without the allgather phase nothing flows back to a sender, so the engine hands out
    one credit per (step, peer): a fast rank must not refill a slow rank's landing zone
    while that rank is still reducing the previous run (found with ThreadSanitizer)."""
    count = 60000

    def fn(ctx):
        base, rem = divmod(count, size)
        recv = [base + (1 if r < rem else 0) for r in range(size)]
        off = int(np.sum(recv[:ctx.rank]))
        buf = np.zeros(count, np.float64)
        algo = alg.ReduceScatterHalvingDoubling(ctx, [buf], recv)
        for it in range(25):
            buf[:] = np.arange(count) * (it + 1) + ctx.rank
            if ctx.rank == it % size:
                time.sleep(0.002)  # this rank enters late; the others are already sending
            algo.run()
            want = np.arange(off, off + recv[ctx.rank]) * (it + 1) * size + size * (size - 1) / 2
            np.testing.assert_allclose(buf[:recv[ctx.rank]], want)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", [1, 2, 4, 7])
def test_allgather_ring(size):
    def fn(ctx):
        for inputs in (1, 3):
            count = 50
            ins = [np.full(count, ctx.rank * 10 + i, np.int32) for i in range(inputs)]
            out = np.zeros(count * inputs * size, np.int32)
            a = alg.AllgatherRing(ctx, ins, out)
            a.run()
            a.run()
            exp = np.concatenate([np.full(count, r * 10 + i, np.int32) for r in range(size) for i in range(inputs)])
            np.testing.assert_array_equal(out, exp)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", [1, 2, 5])
def test_broadcast_one_to_all(size):
    def fn(ctx):
        for root in range(size):
            bufs = [np.full(1000, ctx.rank * 10 + i, np.float32) for i in range(2)]
            a = alg.BroadcastOneToAll(ctx, bufs, root=root, root_pointer=1)
            a.run()
            for b in bufs:
                np.testing.assert_array_equal(b, np.full(1000, root * 10 + 1, np.float32))
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", [1, 2, 4, 7])
def test_barriers_and_pairwise(size):
    def fn(ctx):
        a = alg.BarrierAllToAll(ctx)
        b = alg.BarrierAllToOne(ctx, root=size - 1)
        for _ in range(3):
            a.run()
            b.run()
        if size in (2, 4):
            alg.PairwiseExchange(ctx, 4096, 1).run()
        return True

    assert all(gb.spawn_threads(size, fn))
