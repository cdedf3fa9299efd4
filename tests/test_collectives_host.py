"""New-style host collectives other than allreduce — mirrors gloo/test/{allgather,
allgatherv,alltoall,alltoallv,gather,gatherv,scatter,reduce,broadcast,barrier}_test.cc."""
import numpy as np
import pytest

import gloo_b200 as gb

SIZES = [1, 2, 4, 7]


@pytest.mark.parametrize("size", SIZES)
def test_broadcast(size):
    def fn(ctx):
        for count in (0, 1, 100, 100000):
            for root in range(size):
                buf = np.full(count, ctx.rank, np.float32)
                if ctx.rank == root:
                    src = np.arange(count, dtype=np.float32) + root
                    gb.broadcast(ctx, buf, input=src, root=root)
                else:
                    gb.broadcast(ctx, buf, root=root)
                np.testing.assert_array_equal(buf, np.arange(count, dtype=np.float32) + root)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", SIZES)
def test_allgather(size):
    def fn(ctx):
        for count in (1, 3, 1000):
            inp = np.arange(count, dtype=np.int64) + ctx.rank * 1000000
            out = np.zeros(count * size, np.int64)
            gb.allgather(ctx, out, inp)
            exp = np.concatenate([np.arange(count, dtype=np.int64) + r * 1000000 for r in range(size)])
            np.testing.assert_array_equal(out, exp)
            # in place
            out2 = np.zeros(count * size, np.int64)
            out2[ctx.rank * count:(ctx.rank + 1) * count] = inp
            gb.allgather(ctx, out2)
            np.testing.assert_array_equal(out2, exp)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", SIZES)
def test_allgatherv(size):
    def fn(ctx):
        counts = [(r * 3) % 5 for r in range(size)]  # includes zero-length ranks
        inp = np.full(counts[ctx.rank], ctx.rank, np.int32)
        out = np.full(sum(counts), -1, np.int32)
        gb.allgatherv(ctx, out, counts, inp)
        exp = np.concatenate([np.full(counts[r], r, np.int32) for r in range(size)]) if sum(counts) else out
        np.testing.assert_array_equal(out, exp)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", SIZES)
def test_alltoall(size):
    def fn(ctx):
        for count in (1, 17, 4096):
            inp = np.concatenate([np.full(count, ctx.rank * 100 + j, np.int32) for j in range(size)])
            out = np.zeros_like(inp)
            gb.alltoall(ctx, out, inp)
            exp = np.concatenate([np.full(count, j * 100 + ctx.rank, np.int32) for j in range(size)])
            np.testing.assert_array_equal(out, exp)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", SIZES)
def test_alltoallv(size):
    def fn(ctx):
        r = ctx.rank
        # rank r sends (r + j) % 3 + (1 if j == r else 0) elements to rank j
        send = [((r + j) % 3) + (1 if j == r else 0) for j in range(size)]
        recv = [((j + r) % 3) + (1 if j == r else 0) for j in range(size)]
        inp = np.concatenate([np.full(send[j], r * 100 + j, np.int64) for j in range(size)])
        out = np.full(sum(recv), -1, np.int64)
        gb.alltoallv(ctx, out, recv, inp, send)
        exp = np.concatenate([np.full(recv[j], j * 100 + r, np.int64) for j in range(size)])
        np.testing.assert_array_equal(out, exp)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", SIZES)
def test_gather_scatter(size):
    def fn(ctx):
        for root in range(size):
            inp = np.arange(10, dtype=np.float64) + ctx.rank
            out = np.zeros(10 * size) if ctx.rank == root else None
            gb.gather(ctx, inp, out, root=root)
            if ctx.rank == root:
                exp = np.concatenate([np.arange(10, dtype=np.float64) + r for r in range(size)])
                np.testing.assert_array_equal(out, exp)
            # scatter back
            got = np.zeros(10)
            ins = [np.arange(10, dtype=np.float64) * (j + 1) for j in range(size)] if ctx.rank == root else None
            gb.scatter(ctx, got, ins, root=root)
            np.testing.assert_array_equal(got, np.arange(10, dtype=np.float64) * (ctx.rank + 1))
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", SIZES)
def test_gatherv(size):
    def fn(ctx):
        counts = [(r % 3) + 1 for r in range(size)]
        inp = np.full(counts[ctx.rank], ctx.rank, np.int32)
        root = size - 1
        out = np.zeros(sum(counts), np.int32) if ctx.rank == root else None
        gb.gatherv(ctx, inp, out, counts, root=root)
        if ctx.rank == root:
            np.testing.assert_array_equal(out, np.concatenate([np.full(counts[r], r, np.int32) for r in range(size)]))
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", SIZES)
def test_reduce(size):
    def fn(ctx):
        for count in (1, 5, 1000, 100000):
            for root in (0, size - 1):
                inp = (np.arange(count, dtype=np.float64) * size + ctx.rank).astype(np.float32)
                out = np.zeros(count, np.float32)
                gb.reduce(ctx, out, inp, root=root)
                if ctx.rank == root:
                    exp = (np.arange(count, dtype=np.float64) * size * size + size * (size - 1) / 2).astype(np.float32)
                    np.testing.assert_allclose(out, exp, rtol=1e-5)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", SIZES)
def test_reduce_scatter(size):
    def fn(ctx):
        count = 1003
        inp = (np.arange(count, dtype=np.float64) + ctx.rank).astype(np.float64)
        full = np.arange(count, dtype=np.float64) * size + size * (size - 1) / 2
        # default equal split
        base, rem = divmod(count, size)
        counts = [base + (1 if r < rem else 0) for r in range(size)]
        off = np.cumsum([0] + counts)
        out = np.zeros(counts[ctx.rank])
        gb.reduce_scatter(ctx, out, inp)
        np.testing.assert_allclose(out, full[off[ctx.rank]:off[ctx.rank + 1]])
        # user-chosen counts (including a zero)
        counts2 = [0] * size
        counts2[-1] = count
        out2 = np.zeros(counts2[ctx.rank])
        gb.reduce_scatter(ctx, out2, inp, recv_counts=counts2)
        if ctx.rank == size - 1:
            np.testing.assert_allclose(out2, full)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", [1, 2, 3, 8])
def test_barrier(size):
    import time

    def fn(ctx):
        t = []
        for i in range(5):
            if ctx.rank == i % size:
                time.sleep(0.02)
            gb.barrier(ctx)
            t.append(time.monotonic())
        return t

    res = gb.spawn_threads(size, fn)
    # After each barrier nobody can be ahead of a later-sleeping rank by > the sleep.
    for i in range(5):
        ts = [r[i] for r in res]
        assert max(ts) - min(ts) < 0.02


def test_concurrent_tags():
    """Two collectives in flight on one context, told apart by tag (Slot algebra)."""
    import threading

    def fn(ctx):
        a = np.full(1000, ctx.rank + 1, np.float32)
        b = np.full(1000, 10 * (ctx.rank + 1), np.float32)
        t = threading.Thread(target=lambda: gb.allreduce(ctx, b, tag=7))
        t.start()
        gb.allreduce(ctx, a, tag=3)
        t.join()
        return float(a[0]), float(b[0])

    res = gb.spawn_threads(3, fn)
    assert all(r == (6.0, 60.0) for r in res)
