"""Back-to-back random collectives on one context: every rank draws the same sequence
(shared seed) of operations, sizes and roots and checks each result against numpy. Catches
cross-talk between consecutive collectives (slot reuse, leftovers in the unexpected queue,
messages parked for the single-copy path, acknowledgements arriving late)."""
import numpy as np
import pytest

import gloo_b200 as gb

SIZES = [0, 1, 3, 17, 1000, 4099, 70_001, 300_000]


def _contrib(rank, n, salt):
    return (np.arange(n, dtype=np.float64) % 251) * (salt % 7 + 1) + rank * 3 + salt


@pytest.mark.parametrize("size,seed", [(2, 0), (3, 1), (4, 2), (5, 3), (8, 4)])
def test_random_collective_sequences(size, seed):
    steps = 60 if size <= 4 else 30

    def fn(ctx):
        rng = np.random.RandomState(seed)  # same stream on every rank
        r = ctx.rank
        for step in range(steps):
            op = rng.choice(["allreduce", "allreduce_oop", "bcube", "broadcast", "allgather", "alltoall",
                             "reduce_scatter", "reduce", "gather", "scatter", "barrier", "alltoallv"])
            n = int(rng.choice(SIZES if size <= 4 else SIZES[:-1]))
            root = int(rng.randint(size))
            salt = int(rng.randint(1000))
            total = sum(_contrib(q, n, salt) for q in range(size))
            if op in ("allreduce", "bcube"):
                x = _contrib(r, n, salt)
                gb.allreduce(ctx, x, algorithm=gb.Algorithm.BCUBE if op == "bcube" else gb.Algorithm.RING)
                np.testing.assert_allclose(x, total, err_msg=f"step {step} {op} n={n}")
            elif op == "allreduce_oop":
                x, out = _contrib(r, n, salt), np.full(n, -1.0)
                gb.allreduce(ctx, out, inputs=x)
                np.testing.assert_allclose(out, total, err_msg=f"step {step} {op} n={n}")
                np.testing.assert_array_equal(x, _contrib(r, n, salt))
            elif op == "broadcast":
                x = _contrib(root, n, salt) if r == root else np.zeros(n)
                gb.broadcast(ctx, x, root=root)
                np.testing.assert_array_equal(x, _contrib(root, n, salt))
            elif op == "allgather":
                n = min(n, 70_001)
                out = np.zeros(n * size)
                gb.allgather(ctx, out, _contrib(r, n, salt))
                np.testing.assert_array_equal(out, np.concatenate([_contrib(q, n, salt) for q in range(size)]))
            elif op == "alltoall":
                n = min(n, 70_001)
                inp = np.concatenate([_contrib(r, n, salt) + 1000 * q for q in range(size)]) if n else np.zeros(0)
                out = np.zeros(n * size)
                gb.alltoall(ctx, out, inp)
                want = np.concatenate([_contrib(q, n, salt) + 1000 * r for q in range(size)]) if n else np.zeros(0)
                np.testing.assert_array_equal(out, want)
            elif op == "alltoallv":
                counts = [[int((a * 7 + b * 3 + salt) % 5) * (n // 50 + 1) for b in range(size)] for a in range(size)]
                inp = np.concatenate([np.full(counts[r][q], r * 100 + q, np.float64) for q in range(size)])
                rc = [counts[q][r] for q in range(size)]
                out = np.zeros(sum(rc))
                gb.alltoallv(ctx, out, rc, inp, counts[r])
                want = np.concatenate([np.full(counts[q][r], q * 100 + r, np.float64) for q in range(size)])
                np.testing.assert_array_equal(out, want)
            elif op == "reduce_scatter":
                counts = [n // size + (1 if q < n % size else 0) for q in range(size)]
                out = np.zeros(counts[r])
                gb.reduce_scatter(ctx, out, _contrib(r, n, salt), recv_counts=counts)
                off = sum(counts[:r])
                np.testing.assert_allclose(out, total[off:off + counts[r]])
            elif op == "reduce":
                out = np.zeros(n)
                gb.reduce(ctx, out, _contrib(r, n, salt), root=root)
                if r == root:
                    np.testing.assert_allclose(out, total)
            elif op == "gather":
                n = min(n, 70_001)
                out = np.zeros(n * size) if r == root else None
                gb.gather(ctx, _contrib(r, n, salt), out, root=root)
                if r == root:
                    np.testing.assert_array_equal(out, np.concatenate([_contrib(q, n, salt) for q in range(size)]))
            elif op == "scatter":
                n = min(n, 70_001)
                out = np.zeros(n)
                parts = [_contrib(q, n, salt) for q in range(size)] if r == root else None
                gb.scatter(ctx, out, parts, root=root)
                np.testing.assert_array_equal(out, _contrib(r, n, salt))
            else:
                gb.barrier(ctx)
        return True

    assert all(gb.spawn_threads(size, fn, timeout_ms=60000))


@pytest.mark.parametrize("size,seed", [(2, 11), (3, 12), (4, 13), (7, 14)])
def test_interleaved_algorithm_instances(size, seed):
    """Several old-style algorithm instances alive on one context (each owns its slots and
    bound buffers), run in a random interleaving with fresh data every time - including the
    reduce-scatter-only hypercube, whose flow control is per (step, peer) credit."""
    from gloo_b200.ops import algorithms as alg

    def fn(ctx):
        rng = np.random.RandomState(seed)
        classes = [alg.AllreduceRing, alg.AllreduceRingChunked, alg.AllreduceHalvingDoubling, alg.AllreduceBcube]
        insts = []
        for _ in range(6):
            cls = classes[rng.randint(4)]
            n = int(rng.choice([1, 5, 1000, 4099, 200000]))
            buf = np.zeros(n)
            insts.append((cls(ctx, [buf]), buf, n))
        rs_n = int(rng.choice([size, 1000, 30000]))
        counts = [rs_n // size + (1 if q < rs_n % size else 0) for q in range(size)]
        rsbuf = np.zeros(rs_n)
        rs = alg.ReduceScatterHalvingDoubling(ctx, [rsbuf], counts)
        bar = alg.BarrierAllToAll(ctx)
        for _ in range(40):
            k = rng.randint(len(insts) + 2)
            salt = int(rng.randint(100))
            if k == len(insts):
                rsbuf[:] = np.arange(rs_n) + ctx.rank + salt
                rs.run()
                off = sum(counts[:ctx.rank])
                want = (np.arange(rs_n) + salt) * size + size * (size - 1) / 2
                np.testing.assert_allclose(rsbuf[:counts[ctx.rank]], want[off:off + counts[ctx.rank]])
            elif k == len(insts) + 1:
                bar.run()
            else:
                a, buf, n = insts[k]
                buf[:] = np.arange(n) % 97 + ctx.rank * 2 + salt
                a.run()
                np.testing.assert_allclose(buf, (np.arange(n) % 97 + salt) * size + size * (size - 1))
        return True

    assert all(gb.spawn_threads(size, fn, timeout_ms=60000))


@pytest.mark.parametrize("variant", ["lazy", "shared_device", "sync", "busy_poll"])
def test_random_sequences_on_other_device_modes(variant, monkeypatch):
    """The same random sequences over lazily connected pairs, one device shared by all ranks,
    and pairs in blocking / busy-polling sync mode (where unbound waits drive the sockets)."""
    orig = gb.spawn_threads

    def spawn(n, fn, **kw):
        if variant in ("lazy", "shared_device"):
            kw[variant] = True
            return orig(n, fn, **kw)

        def wrapped(ctx):
            for q in range(n):
                if q != ctx.rank:
                    ctx.get_pair(q).set_sync(True, variant == "busy_poll")
            return fn(ctx)

        return orig(n, wrapped, **kw)

    monkeypatch.setattr(gb, "spawn_threads", spawn)
    for size, seed in ((2, 21), (3, 22), (4, 23)):
        test_random_collective_sequences(size, seed)
