"""Worker process for the fork-as-ranks tests (reference: gloo/test/multiproc_test.{h,cc}).

usage: multiproc_worker.py STORE_DIR RANK SIZE MODE [ARGS...]
Exit codes: 0 ok, 10 IoError (expected when a peer dies), 1 anything else.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import gloo_b200 as gb  # noqa: E402


def main():
    store_dir, rank, size, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    args = sys.argv[5:]
    timeout_ms = int(os.environ.get("GLB_TEST_TIMEOUT_MS", "3000"))
    ctx = gb.init_context(rank, size, path=store_dir, timeout_ms=timeout_ms)
    # GLB_TEST_SYNC: 1 = blocking sync pairs, 2 = busy-polling sync pairs (reference
    # transport_test.cc runs every fault case as Async / Blocking / Polling)
    sync_mode = int(os.environ.get("GLB_TEST_SYNC", "0"))
    if sync_mode:
        for r in range(size):
            if r != rank:
                ctx.get_pair(r).set_sync(True, sync_mode == 2)
    # tell the parent we are connected
    open(os.path.join(store_dir, f"ready_{rank}"), "w").close()
    try:
        if mode == "allreduce_loop":
            n = int(args[0]) if args else 1000
            buf = np.ones(n, np.float32)
            deadline = time.time() + 60
            while time.time() < deadline:
                buf[:] = 1
                gb.allreduce(ctx, buf)
                assert buf[0] == size
        elif mode == "allreduce_once":
            buf = np.full(1000, rank + 1, np.float64)
            gb.allreduce(ctx, buf)
            assert buf[0] == size * (size + 1) / 2, buf[0]
        elif mode == "large_once":
            # Above GLB_TCP_CMA_MIN: payloads are pulled with process_vm_readv between
            # real processes; the result and the counter are both checked.
            n = int(args[0]) if args else (1 << 20)
            before = gb._C.tcp_stats()["cma_messages"]
            buf = np.full(n, rank + 1, np.float32)
            for _ in range(3):
                buf[:] = rank + 1
                gb.allreduce(ctx, buf)
                assert buf[0] == size * (size + 1) / 2 and buf[-1] == buf[0], (buf[0], buf[-1])
            want = os.environ.get("GLB_TCP_CMA", "1") != "0"
            got = gb._C.tcp_stats()["cma_messages"] > before
            assert got == want, (got, want)
        elif mode == "parked":
            # rank 0 sends a large message that rank 1 has not asked for yet (it is parked at
            # rank 1 as a descriptor of rank 0's memory); the parent then kills rank 0 and only
            # afterwards lets rank 1 post the recv, which must fail cleanly.
            big = np.ones(1 << 18, np.float32)
            u = ctx.create_unbound_buffer(big.ctypes.data, big.nbytes)
            gb.barrier(ctx)
            time.sleep(0.1)  # the capability handshake is over: large sends now go header-only
            if rank == 0:
                u.send(1, 77)
                time.sleep(120)
            else:
                while not os.path.exists(os.path.join(store_dir, "go")):
                    time.sleep(0.02)
                u.recv(0, 77)
                u.wait_recv()
                print("unexpectedly received", file=sys.stderr)
        elif mode == "stress_loop":
            # endless random collectives (same sequence on every rank) until a peer dies
            rng = np.random.RandomState(int(args[0]) if args else 0)
            deadline = time.time() + 60
            while time.time() < deadline:
                op = rng.choice(["allreduce", "allgather", "alltoall", "broadcast", "reduce_scatter", "barrier"])
                n = int(rng.choice([1, 1000, 70_000, 400_000]))
                if op == "allreduce":
                    gb.allreduce(ctx, np.ones(n, np.float32))
                elif op == "allgather":
                    gb.allgather(ctx, np.zeros(n * size, np.float32), np.ones(n, np.float32))
                elif op == "alltoall":
                    gb.alltoall(ctx, np.zeros(n * size, np.float32), np.ones(n * size, np.float32))
                elif op == "broadcast":
                    gb.broadcast(ctx, np.ones(n, np.float32), root=int(rng.randint(size)))
                elif op == "reduce_scatter":
                    gb.reduce_scatter(ctx, np.zeros(n, np.float32), np.ones(n * size, np.float32))
                else:
                    gb.barrier(ctx)
        elif mode == "sendrecv_loop":
            peer = (rank + 1) % size
            src = (rank - 1) % size
            a, b = np.ones(100, np.float32), np.zeros(100, np.float32)
            ua = ctx.create_unbound_buffer(a.ctypes.data, a.nbytes)
            ub = ctx.create_unbound_buffer(b.ctypes.data, b.nbytes)
            deadline = time.time() + 60
            while time.time() < deadline:
                ub.recv(src, 1)
                ua.send(peer, 1)
                ub.wait_recv()
                ua.wait_send()
        else:
            raise SystemExit(f"unknown mode {mode}")
    except gb.IoError as e:
        print(f"rank {rank}: IoError: {e}", file=sys.stderr)
        if os.environ.get("GLB_TEST_TRACEBACK"):
            import traceback
            traceback.print_exc()
        sys.exit(10)
    sys.exit(0)


if __name__ == "__main__":
    main()
