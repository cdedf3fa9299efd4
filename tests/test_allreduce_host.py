"""New-style host allreduce — mirrors gloo/test/allreduce_test.cc:301-411."""
import numpy as np
import pytest

import gloo_b200 as gb


def _fixture(rank, size, ptrs, count, dtype):
    # srcs[i][j] = j*stride + rank*ptrs + i with stride = size*ptrs (base_test.h:243-263)
    stride = size * ptrs
    return [(np.arange(count, dtype=np.float64) * stride + rank * ptrs + i).astype(dtype) for i in range(ptrs)]


def _expected(size, ptrs, count, dtype):
    stride = size * ptrs
    return (np.arange(count, dtype=np.float64) * stride * stride + stride * (stride - 1) / 2).astype(dtype)


@pytest.mark.parametrize("algo", [gb.Algorithm.RING, gb.Algorithm.BCUBE])
@pytest.mark.parametrize("size", [1, 2, 4, 7])
@pytest.mark.parametrize("ptrs", [1, 2])
def test_allreduce_sum(algo, size, ptrs):
    counts = [0, 1, 7, 100, 1000, 10000]

    def fn(ctx):
        for count in counts:
            for inplace in (True, False):
                bufs = _fixture(ctx.rank, size, ptrs, count, np.float32)
                if inplace:
                    gb.allreduce(ctx, bufs, algorithm=algo)
                    outs = bufs
                else:
                    outs = [np.zeros(count, np.float32) for _ in range(ptrs)]
                    gb.allreduce(ctx, outs, inputs=bufs, algorithm=algo)
                exp = _expected(size, ptrs, count, np.float32)
                for o in outs:
                    np.testing.assert_allclose(o, exp, rtol=1e-5)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("algo", [gb.Algorithm.RING, gb.Algorithm.BCUBE])
def test_allreduce_many_segments(algo):
    size, count = 4, 10000

    def fn(ctx):
        buf = _fixture(ctx.rank, size, 1, count, np.float32)
        gb.allreduce(ctx, buf, algorithm=algo, max_segment_size=128)
        np.testing.assert_allclose(buf[0], _expected(size, 1, count, np.float32), rtol=1e-5)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("size", [2, 3, 4, 7])
def test_default_algorithm_small_vectors_one_hop(size):
    """Without an explicit algorithm, vectors up to GLB_ALLREDUCE_ONESHOT_MAX (16 KiB) take the
    one-hop exchange: every rank reduces all contributions in rank order, so the results are
    bit-identical everywhere even for sums that are not associative in floating point."""
    def fn(ctx):
        rng = np.random.RandomState(1234 + ctx.rank)
        got = []
        for count in (1, 3, 257, 4096):
            for ptrs in (1, 2):
                bufs = [(rng.randn(count) * 10 ** rng.randint(-3, 4)).astype(np.float32) for _ in range(ptrs)]
                mine = np.sum(np.stack(bufs).astype(np.float64), axis=0)
                out = np.zeros(count, np.float32)
                gb.allreduce(ctx, out, inputs=bufs if ptrs > 1 else bufs[0])
                ref = np.zeros(count, np.float64)
                ref[:] = mine
                gb.allreduce(ctx, ref, algorithm=gb.Algorithm.RING)  # fp64 reference over the ring
                np.testing.assert_allclose(out, ref, rtol=2e-4, atol=1e-2)  # float32 accumulation
                inplace = bufs[0].copy()
                gb.allreduce(ctx, inplace)
                got.append(out.tobytes())
                got.append(inplace.tobytes())
        return got

    res = gb.spawn_threads(size, fn)
    for r in range(1, size):
        for a, b in zip(res[0][::2], res[r][::2]):
            assert a == b, "ranks disagree bitwise"


@pytest.mark.parametrize("size", [2, 3, 5])
@pytest.mark.parametrize("max_segment", [None, 128, 4096])
def test_ring_streams_segments_and_leaves_the_input_alone(size, max_segment):
    """The ring is one pipeline of segments and reads a single out-of-place input in place
    (no up-front copy): the input must come back untouched, every length (including ones
    that leave some chunks short or empty) must work in place and out of place, and large
    vectors take the single-copy transport path between the rank threads."""
    counts = [1, size - 1, size + 1, 1000, 4099, 700_001]

    def fn(ctx):
        for count in counts:
            kw = {} if max_segment is None else {"max_segment_size": max_segment}
            if max_segment == 128 and count > 10_000:
                continue  # thousands of 128-byte segments add nothing
            src = (np.arange(count, dtype=np.float64) % 997 + ctx.rank).astype(np.float64)
            keep = src.copy()
            out = np.full(count, -1.0)
            gb.allreduce(ctx, out, inputs=src, **kw)
            want = (np.arange(count, dtype=np.float64) % 997) * size + size * (size - 1) / 2
            np.testing.assert_array_equal(src, keep)
            np.testing.assert_array_equal(out, want)
            gb.allreduce(ctx, src, **kw)  # in place
            np.testing.assert_array_equal(src, want)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("dtype", [np.int8, np.uint8, np.int32, np.int64, np.uint64, np.float64, np.float16])
@pytest.mark.parametrize("op", [gb.ReduceOp.SUM, gb.ReduceOp.PRODUCT, gb.ReduceOp.MIN, gb.ReduceOp.MAX])
def test_allreduce_types_ops(dtype, op):
    size, count = 3, 257

    def fn(ctx):
        rng = np.random.RandomState(1234)
        data = [rng.randint(1, 4, size=count).astype(dtype) for _ in range(size)]
        buf = data[ctx.rank].copy()
        gb.allreduce(ctx, buf, op=op)
        f = {gb.ReduceOp.SUM: np.add, gb.ReduceOp.PRODUCT: np.multiply, gb.ReduceOp.MIN: np.minimum,
             gb.ReduceOp.MAX: np.maximum}[op]
        exp = data[0].copy()
        for d in data[1:]:
            exp = f(exp, d).astype(dtype)
        np.testing.assert_array_equal(buf, exp)
        return True

    assert all(gb.spawn_threads(size, fn))


def test_allreduce_custom_function():
    import ctypes

    size, count = 2, 64

    def fn(ctx):
        def myfn(c, a, b, n):
            ca = np.ctypeslib.as_array(ctypes.cast(a, ctypes.POINTER(ctypes.c_float)), (n,))
            cb = np.ctypeslib.as_array(ctypes.cast(b, ctypes.POINTER(ctypes.c_float)), (n,))
            cc = np.ctypeslib.as_array(ctypes.cast(c, ctypes.POINTER(ctypes.c_float)), (n,))
            cc[:] = ca + 2 * cb

        buf = np.full(count, float(ctx.rank + 1), np.float32)
        gb.allreduce(ctx, buf, op=myfn)
        return buf.copy()

    res = gb.spawn_threads(size, fn)
    np.testing.assert_array_equal(res[0], res[1])


def test_allreduce_timeout():
    # One rank never joins: the others must fail with a timeout IoError.
    def fn(ctx):
        if ctx.rank == 0:
            return "skipped"
        buf = np.ones(16, np.float32)
        with pytest.raises(gb.IoError, match="Timed out"):
            gb.allreduce(ctx, buf, timeout_ms=50)
        return "timeout"

    res = gb.spawn_threads(2, fn)
    assert res == ["skipped", "timeout"]
