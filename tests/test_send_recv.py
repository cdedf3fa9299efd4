"""Point-to-point on unbound buffers — mirrors gloo/test/send_recv_test.cc:26-518 and
remote_key_test.cc:62-170 (one-sided put/get, which the reference only has on ibverbs)."""
import threading
import time

import numpy as np
import pytest

import gloo_b200 as gb
from gloo_b200 import _C


def ub(ctx, arr):
    return ctx.create_unbound_buffer(arr.ctypes.data, arr.nbytes)


@pytest.mark.parametrize("size", [2, 3, 5, 8])
def test_all_to_all(size):
    def fn(ctx):
        slot = 0x1337
        outs = [np.full(4, ctx.rank * 100 + j, np.int32) for j in range(size)]
        ins = [np.full(4, -1, np.int32) for _ in range(size)]
        bufs_in = [ub(ctx, a) for a in ins]
        bufs_out = [ub(ctx, a) for a in outs]
        for j in range(size):
            if j != ctx.rank:
                bufs_in[j].recv(j, slot)
        for j in range(size):
            if j != ctx.rank:
                bufs_out[j].send(j, slot)
        for j in range(size):
            if j != ctx.rank:
                assert bufs_in[j].wait_recv() == j
                assert bufs_out[j].wait_send() == j
                np.testing.assert_array_equal(ins[j], np.full(4, j * 100 + ctx.rank, np.int32))
        return True

    assert all(gb.spawn_threads(size, fn))


def test_offsets_and_empty_messages():
    def fn(ctx):
        peer = 1 - ctx.rank
        data = np.arange(16, dtype=np.int32) + 100 * ctx.rank
        recv = np.full(16, -1, np.int32)
        s, r = ub(ctx, data), ub(ctx, recv)
        # zero-byte message first, then a non-empty one on the same slot
        r.recv(peer, 5, 0, 0)
        s.send(peer, 5, 0, 0)
        r.wait_recv()
        s.wait_send()
        # middle 8 elements -> last 8 elements
        r.recv(peer, 5, 8 * 4, 8 * 4)
        s.send(peer, 5, 4 * 4, 8 * 4)
        r.wait_recv()
        s.wait_send()
        np.testing.assert_array_equal(recv[:8], np.full(8, -1))
        np.testing.assert_array_equal(recv[8:], np.arange(4, 12) + 100 * peer)
        return True

    assert all(gb.spawn_threads(2, fn))


def test_unexpected_messages_are_buffered_in_order():
    """Sends that arrive before the recv is posted are parked and matched FIFO."""
    def fn(ctx):
        n = 50
        if ctx.rank == 0:
            bufs = [np.full(3, i, np.int64) for i in range(n)]
            ubs = [ub(ctx, b) for b in bufs]
            for u in ubs:
                u.send(1, 9)
            for u in ubs:
                u.wait_send()
            _C.barrier(ctx, 1)
        else:
            time.sleep(0.2)  # let everything arrive first
            got = np.zeros(3, np.int64)
            u = ub(ctx, got)
            for i in range(n):
                u.recv(0, 9)
                u.wait_recv()
                assert got[0] == i
            _C.barrier(ctx, 1)
        return True

    assert all(gb.spawn_threads(2, fn))


@pytest.mark.parametrize("size", [3, 4, 8])
def test_recv_from_any(size):
    def fn(ctx):
        slot = 77
        if ctx.rank == 0:
            seen = set()
            got = np.zeros(2, np.int32)
            u = ub(ctx, got)
            for _ in range(size - 1):
                u.recv(list(range(1, size)), slot)
                src = u.wait_recv()
                assert got[0] == src and got[1] == src * 2
                seen.add(src)
            assert seen == set(range(1, size))
        else:
            d = np.array([ctx.rank, ctx.rank * 2], np.int32)
            u = ub(ctx, d)
            u.send(0, slot)
            u.wait_send()
        return True

    assert all(gb.spawn_threads(size, fn))


def test_recv_from_any_rpc_pipeline():
    """Parameter-server pattern: the server answers whoever asks, many rounds."""
    size, rounds = 4, 20

    def fn(ctx):
        if ctx.rank == 0:
            req = np.zeros(1, np.int64)
            ureq = ub(ctx, req)
            for _ in range((size - 1) * rounds):
                ureq.recv(list(range(1, size)), 1)
                src = ureq.wait_recv()
                resp = np.array([req[0] * 2], np.int64)
                ur = ub(ctx, resp)
                ur.send(src, 2)
                ur.wait_send()
        else:
            for i in range(rounds):
                req = np.array([ctx.rank * 1000 + i], np.int64)
                resp = np.zeros(1, np.int64)
                a, b = ub(ctx, req), ub(ctx, resp)
                b.recv(0, 2)
                a.send(0, 1)
                a.wait_send()
                b.wait_recv()
                assert resp[0] == 2 * (ctx.rank * 1000 + i)
        return True

    assert all(gb.spawn_threads(size, fn))


def test_abort_wait_recv_and_send():
    def fn(ctx):
        buf = np.zeros(4, np.float32)
        u = ub(ctx, buf)
        u.recv(1 - ctx.rank, 123)
        res = []
        t = threading.Thread(target=lambda: res.append(u.wait_recv(timeout_ms=10000)))
        t.start()
        time.sleep(0.05)
        u.abort_wait_recv()
        t.join()
        assert res == [None]
        # the buffer stays usable afterwards
        peer = 1 - ctx.rank
        src = np.full(4, ctx.rank + 1, np.float32)
        us = ub(ctx, src)
        u.recv(peer, 124)
        us.send(peer, 124)
        assert u.wait_recv() == peer
        us.wait_send()
        np.testing.assert_array_equal(buf, np.full(4, peer + 1, np.float32))
        return True

    assert all(gb.spawn_threads(2, fn))


def test_recv_timeout_poisons_context():
    def fn(ctx):
        buf = np.zeros(4, np.float32)
        u = ub(ctx, buf)
        if ctx.rank == 0:
            u.recv(1, 55)
            with pytest.raises(gb.TimeoutError, match="Timed out"):
                u.wait_recv(timeout_ms=30)
            # every later operation on this context fails fast
            with pytest.raises(gb.IoError):
                u.send(1, 56)
        return True

    assert all(gb.spawn_threads(2, fn))


def test_size_mismatch_detected():
    def fn(ctx):
        if ctx.rank == 0:
            big = np.zeros(100, np.int32)
            u = ub(ctx, big)
            u.send(1, 3)
            u.wait_send()
        else:
            small = np.zeros(10, np.int32)
            u = ub(ctx, small)
            time.sleep(0.1)
            with pytest.raises((gb.EnforceError, gb.IoError), match="mismatch"):
                u.recv(0, 3)
                u.wait_recv(timeout_ms=2000)
        return True

    assert all(gb.spawn_threads(2, fn))


def test_put_get_remote_key():
    """One-sided access through RemoteKey: keys exchanged with allgather, then rank r
    puts into / gets from rank r+1 without the target posting anything."""
    size = 3

    def fn(ctx):
        window = np.full(8, ctx.rank * 10, np.int64)
        uw = ub(ctx, window)
        key = uw.get_remote_key().encode()
        keys = np.zeros(64 * size, np.uint8)
        mine = np.zeros(64, np.uint8)
        mine[:len(key)] = np.frombuffer(key, np.uint8)
        gb.allgather(ctx, keys, mine)
        all_keys = [bytes(keys[i * 64:(i + 1) * 64]).rstrip(b"\0").decode() for i in range(size)]
        right = (ctx.rank + 1) % size
        left = (ctx.rank - 1) % size
        # get: read the right neighbour's window
        got = np.zeros(8, np.int64)
        ug = ub(ctx, got)
        ug.get(ctx, all_keys[right], 0, 0, 0, 64)
        assert ug.wait_recv() == right
        np.testing.assert_array_equal(got, np.full(8, right * 10))
        gb.barrier(ctx)
        # put: overwrite the second half of the right neighbour's window
        src = np.full(4, 1000 + ctx.rank, np.int64)
        us = ub(ctx, src)
        us.put(ctx, all_keys[right], 0, 0, 4 * 8, 4 * 8)
        us.wait_send()
        # a get through the same pair is ordered behind the put: use it as a flush
        ug.get(ctx, all_keys[right], 0, 0, 0, 64)
        ug.wait_recv()
        np.testing.assert_array_equal(got[4:], np.full(4, 1000 + ctx.rank))
        gb.barrier(ctx)
        np.testing.assert_array_equal(window[:4], np.full(4, ctx.rank * 10))
        np.testing.assert_array_equal(window[4:], np.full(4, 1000 + left))
        # bounds are enforced locally
        with pytest.raises(gb.EnforceError):
            us.put(ctx, all_keys[right], 0, 0, 60, 32)
        return True

    assert all(gb.spawn_threads(size, fn))


@pytest.mark.parametrize("half", [32, 1 << 17])  # small: eager; large: single-copy path
def test_bound_buffers_and_sync_mode(half):
    for sync, busy in ((False, False), (True, False), (True, True)):
        def fn(ctx):
            peer = 1 - ctx.rank
            pair = ctx.get_pair(peer)
            if sync:
                pair.set_sync(True, busy)
            src = np.arange(half, dtype=np.float32) + ctx.rank
            dst = np.zeros(2 * half, np.float32)
            slot = ctx.next_slot()
            sb = pair.create_send_buffer(slot, src.ctypes.data, src.nbytes)
            rb = pair.create_recv_buffer(slot, dst.ctypes.data, dst.nbytes)
            # write into the second half of the peer's buffer (remote offset)
            for _ in range(3):
                sb.send(0, src.nbytes, half * 4)
                rb.wait_recv()
                sb.wait_send()
            np.testing.assert_array_equal(dst[half:], np.arange(half, dtype=np.float32) + peer)
            np.testing.assert_array_equal(dst[:half], np.zeros(half, np.float32))
            return True

        assert all(gb.spawn_threads(2, fn))


def test_lazy_device_and_shared_device():
    def fn(ctx):
        buf = np.full(100, ctx.rank + 1, np.float64)
        gb.allreduce(ctx, buf)
        return float(buf[0])

    assert gb.spawn_threads(4, fn, lazy=True) == [10.0] * 4
    assert gb.spawn_threads(4, fn, shared_device=True) == [10.0] * 4


def test_large_messages_take_the_single_copy_path():
    """>= GLB_TCP_CMA_MIN bytes: header on the wire, payload pulled from the sender's address
    space (here: the same process) — unbound send/recv incl. unexpected arrival, bound
    buffers with a remote offset, put and get."""
    n = 1 << 18  # 1 MiB of float32
    before = _C.tcp_stats()["cma_messages"]

    def fn(ctx):
        peer = 1 - ctx.rank
        gb.barrier(ctx)  # the capability handshake has long finished after a round trip
        time.sleep(0.05)
        # unbound, posted first on rank 0, unexpected on rank 1
        src = np.arange(n, dtype=np.float32) + ctx.rank
        dst = np.zeros(n, np.float32)
        us, ud = ub(ctx, src), ub(ctx, dst)
        if ctx.rank == 0:
            ud.recv(peer, 7)
            gb.barrier(ctx)
            us.send(peer, 7)
        else:
            gb.barrier(ctx)
            us.send(peer, 7)
            time.sleep(0.1)
            ud.recv(peer, 7)
        assert ud.wait_recv() == peer
        assert us.wait_send() == peer
        np.testing.assert_array_equal(dst, np.arange(n, dtype=np.float32) + peer)
        # bound, remote offset
        pair = ctx.get_pair(peer)
        big = np.zeros(2 * n, np.float32)
        slot = ctx.next_slot()
        sb = pair.create_send_buffer(slot, src.ctypes.data, src.nbytes)
        rb = pair.create_recv_buffer(slot, big.ctypes.data, big.nbytes)
        sb.send(0, src.nbytes, n * 4)
        rb.wait_recv()
        sb.wait_send()
        np.testing.assert_array_equal(big[n:], np.arange(n, dtype=np.float32) + peer)
        assert not big[:n].any()
        # one-sided
        window = np.full(n, ctx.rank + 10, np.float32)
        uw = ub(ctx, window)
        key = uw.get_remote_key().encode()
        keys = np.zeros(128, np.uint8)
        mine = np.zeros(64, np.uint8)
        mine[:len(key)] = np.frombuffer(key, np.uint8)
        gb.allgather(ctx, keys, mine)
        peer_key = bytes(keys[peer * 64:(peer + 1) * 64]).rstrip(b"\0").decode()
        got = np.zeros(n, np.float32)
        ug = ub(ctx, got)
        ug.get(ctx, peer_key, 0, 0, 0, got.nbytes)
        assert ug.wait_recv() == peer
        np.testing.assert_array_equal(got, np.full(n, peer + 10, np.float32))
        gb.barrier(ctx)
        us.put(ctx, peer_key, 0, 0, 0, src.nbytes)
        us.wait_send()
        ug.get(ctx, peer_key, 0, 0, 0, got.nbytes)  # ordered behind the put
        ug.wait_recv()
        np.testing.assert_array_equal(got, src)
        gb.barrier(ctx)
        return True

    assert all(gb.spawn_threads(2, fn))
    assert _C.tcp_stats()["cma_messages"] - before >= 10


@pytest.mark.parametrize("busy", [False, True])
def test_unbound_buffers_on_sync_pairs(busy):
    """Pairs switched to sync mode are detached from the loop thread; unbound-buffer waits
    then read the socket themselves, so the new-style collectives keep working (the
    reference rejects unbound buffers on sync pairs, tcp/pair.cc)."""
    size = 3

    def fn(ctx):
        for r in range(size):
            if r != ctx.rank:
                ctx.get_pair(r).set_sync(True, busy)
        x = np.full(1000, ctx.rank + 1, np.float32)
        for _ in range(5):
            x[:] = ctx.rank + 1
            gb.allreduce(ctx, x)
            assert x[0] == 6
        big = np.full(1 << 18, ctx.rank + 1, np.float32)  # single-copy path, FIN read by the waiter
        gb.allreduce(ctx, big)
        assert big[0] == 6 and big[-1] == 6
        out = np.zeros(size * 4, np.int64)
        gb.allgather(ctx, out, np.full(4, ctx.rank, np.int64))
        np.testing.assert_array_equal(out, np.repeat(np.arange(size), 4))
        gb.barrier(ctx)
        return True

    t0 = time.time()
    assert all(gb.spawn_threads(size, fn))
    assert time.time() - t0 < 8  # the closing barrier of spawn_threads must not time out


def test_parked_large_messages_complete_out_of_order():
    """Large unbound messages whose recv is not posted yet are parked as descriptors and
    pulled by the thread that posts the recv. Post the recvs in the opposite order of the
    sends: each FIN must complete its own send (ids, not FIFO), the sender's memory must stay
    untouched until then, and a small eager message in between is unaffected."""
    n = 1 << 17  # 512 KiB

    def fn(ctx):
        peer = 1 - ctx.rank
        gb.barrier(ctx)
        time.sleep(0.05)
        if ctx.rank == 0:
            a = np.full(n, 1.0, np.float32)
            b = np.full(n, 2.0, np.float32)
            c = np.full(4, 3.0, np.float32)
            ua, ub_, uc = ub(ctx, a), ub(ctx, b), ub(ctx, c)
            ua.send(peer, 11)
            ub_.send(peer, 22)
            uc.send(peer, 33)
            assert uc.wait_send() == peer          # eager: completes at once
            # the receiver takes slot 22 first, so b's FIN arrives before a's
            assert ub_.wait_send() == peer
            assert ua.wait_send() == peer
        else:
            time.sleep(0.2)                         # both large messages are parked by now
            a, b, c = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(4, np.float32)
            ua, ub_, uc = ub(ctx, a), ub(ctx, b), ub(ctx, c)
            ub_.recv(peer, 22)
            assert ub_.wait_recv() == peer and b[0] == 2.0 and b[-1] == 2.0
            uc.recv(peer, 33)
            assert uc.wait_recv() == peer and c[0] == 3.0
            ua.recv(peer, 11)
            assert ua.wait_recv() == peer and a[0] == 1.0 and a[-1] == 1.0
        gb.barrier(ctx)
        return True

    before = _C.tcp_stats()["cma_messages"]
    assert all(gb.spawn_threads(2, fn))
    assert _C.tcp_stats()["cma_messages"] - before >= 2
