"""One process per GPU for the device-side failure-detection test (mirrors the reference's
transport fault tests, gloo/test/transport_test.cc:53-164, on the NVLink data plane).
usage: cuda_fault_worker.py STORE_DIR RANK SIZE MODE ALGO   (MODE: kill | stop)"""
import os
import signal
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import gloo_b200 as gb  # noqa: E402
from gloo_b200.ops import cuda as gcu  # noqa: E402

TIMEOUT_MS = 3000


def main():
    store_dir, rank, size, mode, algo = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    ctx = gb.init_context(rank, size, path=store_dir, timeout_ms=60000)
    cc = gcu.CudaContext(ctx, dev, stage_bytes=32 << 20)
    cc.set_timeout(TIMEOUT_MS)
    n = 1000 if algo == "ll" else 1 << 20
    t = torch.ones(n, device="cuda")
    if algo == "two_shot":
        cc.register(t)
    for _ in range(3):  # healthy collectives first
        cc.allreduce(t, algo=algo)
    cc.synchronize()
    assert float(t[0]) == size ** 3
    gb.barrier(ctx)
    print(f"READY {rank}", flush=True)
    if rank == size - 1:
        # the victim disappears right before the next collective
        if mode == "kill":
            os.kill(os.getpid(), signal.SIGKILL)
        else:
            os.kill(os.getpid(), signal.SIGSTOP)
        time.sleep(3600)
    t0 = time.time()
    cc.allreduce(t, algo=algo)
    try:
        cc.synchronize()
    except gb.IoError as e:
        dt = time.time() - t0
        assert dt < 2 * TIMEOUT_MS / 1000 + 2, f"took {dt:.1f}s"
        # every later call fails fast; the GPU itself is fine
        try:
            cc.allreduce(t, algo=algo)
            raise SystemExit("a poisoned context accepted another collective")
        except gb.IoError:
            pass
        x = torch.arange(1000, device="cuda").float().sum().item()
        assert x == 499500.0
        print(f"SURVIVOR {rank} OK after {dt:.2f}s: {str(e)[:120]}", flush=True)
        os._exit(0)  # the control plane to the dead peer is gone: skip orderly teardown
    raise SystemExit("the collective with a dead peer completed?!")


if __name__ == "__main__":
    main()
