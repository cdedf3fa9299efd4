#!/usr/bin/env python
"""Small deterministic workloads for Nsight Compute captures.

  local   : 1 rank, CudaAllreduceRingChunked with 2 local buffers (localReduceMany +
            localBroadcast kernels; safe for kernel replay)
  twoshot : 2 ranks as threads on cuda:0, one fused two-shot allreduce on 64 MB
            (use --replay-mode application: the kernel waits for its peer, so it cannot
            be replayed in isolation)
  oneshot : same with a 16 KB one-shot
"""
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import gloo_b200 as gb  # noqa: E402
from gloo_b200.ops import cuda as gcu  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "local"

if mode.startswith("loop_"):
    # Single-rank loopback of a collective kernel (GLB_CUDA_FORCE_KERNELS=1): same code,
    # all "peers" are this GPU, so the kernel can be replayed by ncu.
    os.environ["GLB_CUDA_FORCE_KERNELS"] = "1"
    algo = mode[len("loop_"):]
    n = (256 << 20) // 4 if algo == "two_shot" else 4096

    def fn(ctx):
        cc = gcu.CudaContext(ctx, 0, stage_bytes=8 << 20)
        t = cc.empty(n, torch.float32)
        t.fill_(1)
        for _ in range(3):
            cc.allreduce(t, algo=algo)
        torch.cuda.current_stream().synchronize()
        return float(t[0])
    print(gb.spawn_threads(1, fn, cuda_device=0))
elif mode == "local":
    def fn(ctx):
        n = 100_000_000
        ts = [torch.ones(n, device="cuda") for _ in range(2)]
        algo = gcu.CudaAllreduceRingChunked(ctx, ts)
        for _ in range(4):
            algo.run()
        return float(ts[0][0])
    print(gb.spawn_threads(1, fn, cuda_device=0))
else:
    n = (64 << 20) // 4 if mode == "twoshot" else 4096

    def fn(ctx):
        cc = gcu.CudaContext(ctx, 0, stage_bytes=8 << 20)
        t = torch.ones(n, device="cuda")
        cc.register(t)
        for _ in range(3):
            cc.allreduce(t, algo="two_shot" if mode == "twoshot" else "one_shot")
        torch.cuda.current_stream().synchronize()
        cc.pc.host_barrier()
        return float(t[0])
    print(gb.spawn_threads(2, fn, cuda_device=0))
