#!/usr/bin/env python
"""Broadcast variants on symmetric buffers, one rank per GPU (torchrun): direct push (mode 0),
scatter + allgather (1), multimem.st (2), chunk-pipelined relay (3) and NCCL on the same
buffer. Device timed, max over ranks, p50; bandwidth = bytes / time (every rank ends up with
the whole buffer). Writes one JSON document.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import gloo_b200 as gb  # noqa: E402
from gloo_b200.ops import cuda as gcu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--sizes", default="1048576,4194304,33554432,268435456,1073741824")
    ap.add_argument("--tiles", default="0,1024,4096")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    ctx = gb.init_context(rank, world, path=f"/tmp/glb_bc_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}", timeout_ms=120000)
    cc = gcu.CudaContext(ctx, local, stage_bytes=64 << 20)
    nccl = None
    try:
        nccl = gb._C.cuda.NcclComm.init_rank(ctx, local)
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print("NCCL comparator unavailable:", e)
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    F32 = int(gb.DataType.FLOAT32)

    def timed(fn, nbytes):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
        with torch.cuda.stream(stream):
            for _ in range(3):
                fn()
            for a, b in evs:
                if nbytes < (128 << 20):
                    flush.fill_(0)
                a.record(stream)
                fn()
                b.record(stream)
        stream.synchronize()
        per = np.asarray([a.elapsed_time(b) * 1e3 for a, b in evs], dtype=np.float64)
        if world > 1:
            gb.allreduce(ctx, per, op=gb.ReduceOp.MAX)
        per.sort()
        return float(per[len(per) // 2])

    rows = []
    for nbytes in [int(x) for x in args.sizes.split(",")]:
        n = nbytes // 4
        t = cc.empty(n, torch.float32)
        row = {"bytes": nbytes}
        variants = [("direct", "0", "0"), ("scatter_allgather", "1", "0")]
        if cc.nvls_available():
            variants.append(("multimem_st", "2", "0"))
        if world > 2:
            variants += [(f"relay_tile{tl}" if tl != "0" else "relay", "3", tl) for tl in args.tiles.split(",")]
        for name, mode, tile in variants:
            if mode == "0" and nbytes * (world - 1) > (2 << 30):
                continue
            os.environ["GLB_CUDA_BCAST_MODE"], os.environ["GLB_CUDA_BCAST_TILE"] = mode, tile
            t.fill_(float(rank))
            torch.cuda.synchronize()
            gb.barrier(ctx)
            us = timed(lambda: cc.broadcast(t, root=0, stream=stream), nbytes)
            stream.synchronize()
            ok = float(t[0]) == 0.0 and float(t[-1]) == 0.0 and float(t[n // 2 + 1]) == 0.0
            row[name + "_us"], row[name + "_gbs"] = round(us, 2), round(nbytes / (us * 1e-6) / 1e9, 1)
            if not ok:
                row[name + "_WRONG"] = True
            gb.barrier(ctx)
        os.environ.pop("GLB_CUDA_BCAST_MODE", None)
        os.environ.pop("GLB_CUDA_BCAST_TILE", None)
        us = timed(lambda: cc.broadcast(t, root=0, stream=stream), nbytes)
        row["auto_us"], row["auto_gbs"] = round(us, 2), round(nbytes / (us * 1e-6) / 1e9, 1)
        if nccl is not None:
            us = timed(lambda: nccl.broadcast(t.data_ptr(), t.data_ptr(), n, F32, 0, stream.cuda_stream), nbytes)
            row["nccl_us"], row["nccl_gbs"] = round(us, 2), round(nbytes / (us * 1e-6) / 1e9, 1)
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        del t
        torch.cuda.synchronize()
        gb.barrier(ctx)
    if rank == 0 and args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump({"world": world, "describe": cc.describe(), "rows": rows}, f, indent=1)
    ctx.close_connections()


if __name__ == "__main__":
    main()
