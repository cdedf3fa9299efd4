"""Runs the ibverbs transport (csrc/glb/transport/ibverbs/transport.cc) with threads as ranks.
The verbs library is whatever GLB_IBVERBS_LIB names: in CI the software provider built from
tests/fake_ibverbs/fake_ibverbs.cc (in-process RC queue pairs), on a machine with an HCA the
real libibverbs.so.1. usage: ibverbs_worker.py SIZE"""
import os
import sys
import threading
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import gloo_b200 as gb  # noqa: E402
from gloo_b200 import _C  # noqa: E402
from gloo_b200.ops import algorithms as alg  # noqa: E402


def rank_main(ctx, size):
    r = ctx.rank
    assert "ibverbs(" in str(ctx.device())
    # ---- new-style collectives (unbound buffers: eager below 8 KB, rendezvous / RDMA READ above)
    for n in (1, 100, 5000, 300_000):
        a = np.arange(n, dtype=np.float32) * size + r
        gb.allreduce(ctx, a)
        np.testing.assert_allclose(a, np.arange(n, dtype=np.float64) * size * size + size * (size - 1) / 2, rtol=1e-6)
        out = np.zeros(n * size, dtype=np.int64)
        gb.allgather(ctx, out, np.full(n, r, dtype=np.int64))
        assert out[0] == 0 and out[-1] == size - 1
        b = np.full(n, r, dtype=np.float64)
        gb.broadcast(ctx, b, root=size - 1)
        assert b[0] == size - 1
        a2a_in = np.concatenate([np.full(n, r * 100 + j, dtype=np.int32) for j in range(size)])
        a2a_out = np.zeros(n * size, dtype=np.int32)
        gb.alltoall(ctx, a2a_out, a2a_in)
        assert list(a2a_out[::n]) == [j * 100 + r for j in range(size)]
    gb.barrier(ctx)
    # ---- old-style algorithms (bound buffers: MR exchange + RDMA WRITE WITH IMMEDIATE)
    for cls in (alg.AllreduceRing, alg.AllreduceRingChunked, alg.AllreduceHalvingDoubling, alg.AllreduceBcube):
        for n in (7, 40_000):
            buf = np.arange(n, dtype=np.float32) * size + r
            algo = cls(ctx, buf)
            algo.run()
            algo.run()  # re-run: sums of sums
            np.testing.assert_allclose(buf, (np.arange(n, dtype=np.float64) * size * size + size * (size - 1) / 2) * size, rtol=1e-5)
    # ---- point to point: recv-from-any, offsets, one-sided put / get through remote keys
    if size > 1:
        slot = _C.slot_build(0x55, 1)
        if r == 0:
            seen = set()
            for _ in range(size - 1):
                buf = np.zeros(20_000, dtype=np.int32)
                ub = ctx.create_unbound_buffer(buf.ctypes.data, buf.nbytes)
                ub.recv(list(range(1, size)), slot)
                src = ub.wait_recv()
                assert buf[0] == src and buf[-1] == src
                seen.add(src)
            assert seen == set(range(1, size))
        else:
            buf = np.full(20_000, r, dtype=np.int32)
            ub = ctx.create_unbound_buffer(buf.ctypes.data, buf.nbytes)
            ub.send(0, slot)
            ub.wait_send()
        window = np.full(1024, -1, dtype=np.int64)
        wb = ctx.create_unbound_buffer(window.ctypes.data, window.nbytes)
        key = wb.get_remote_key().encode().ljust(128, b" ")
        keys = np.zeros(128 * size, dtype=np.uint8)
        gb.allgather(ctx, keys, np.frombuffer(key, dtype=np.uint8).copy())
        right = (r + 1) % size
        rkey = bytes(keys[128 * right:128 * (right + 1)]).decode().strip()
        src = np.full(8, r, dtype=np.int64)
        sb = ctx.create_unbound_buffer(src.ctypes.data, src.nbytes)
        sb.put(ctx, rkey, 1, 0, 8 * 8 * r, src.nbytes)  # my 8 values at offset 8*r of my right neighbour's window
        sb.wait_send()
        gb.barrier(ctx)
        left = (r - 1) % size
        assert window[8 * left] == left and window[8 * left + 7] == left
        got = np.zeros(8, dtype=np.int64)
        g = ctx.create_unbound_buffer(got.ctypes.data, got.nbytes)
        g.get(ctx, rkey, 2, 0, 8 * 8 * r, got.nbytes)
        g.wait_recv()
        assert got[0] == r
        try:
            sb.put(ctx, rkey, 3, 0, 1024 * 8 - 8, src.nbytes)
            raise SystemExit("out-of-range put was accepted")
        except gb.GlbError:
            pass
    gb.barrier(ctx)
    return True


def main():
    size = int(sys.argv[1])
    store = _C.HashStore()
    ok = [False] * size
    errs = []

    def run(rank):
        try:
            dev = _C.create_ibverbs_device("", 1, 0)
            ctx = _C.Context(rank, size, 2)
            ctx.set_timeout(20000)
            ctx.connect_full_mesh(store, dev)
            ok[rank] = rank_main(ctx, size)
            _C.barrier(ctx, 0xFFFFF0, 10000)
            ctx.close_connections()
        except BaseException as e:  # noqa: BLE001
            traceback.print_exc()
            errs.append(e)

    ts = [threading.Thread(target=run, args=(r,)) for r in range(size)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errs or not all(ok):
        raise SystemExit(f"FAILED: {errs}")
    print(f"IBVERBS OK {size}", flush=True)


if __name__ == "__main__":
    main()
