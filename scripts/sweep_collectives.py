#!/usr/bin/env python
"""Bandwidth sweep of the data-movement collectives (configs 4 and 5 of BASELINE.json):
allgather, alltoall, alltoall_v, broadcast, reduce_scatter — 1 KB .. 1 GB per-rank payload,
device-timed, max over ranks, against the NCCL comparator. torchrun, one rank per GPU.
Also the named fp16 schedules (config 3) with --schedules."""
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
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--max-bytes", type=int, default=1 << 30)
    ap.add_argument("--ops", default="allgather,alltoall,alltoall_v,broadcast,reduce_scatter")
    ap.add_argument("--schedules", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    ctx = gb.init_context(rank, world, path=f"/tmp/glb_swc_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}",
                          timeout_ms=120000)
    cc = gcu.CudaContext(ctx, local, stage_bytes=64 << 20)
    if rank == 0:
        print(cc.describe(), flush=True)
    try:
        nccl = gb._C.cuda.NcclComm.init_rank(ctx, local)
    except Exception as e:  # noqa: BLE001
        nccl = None
        if rank == 0:
            print("NCCL unavailable:", e)
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    F32 = int(gb.DataType.FLOAT32)

    def timed(fn, iters):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        with torch.cuda.stream(stream):
            for _ in range(3):
                fn()
            for a, b in evs:
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
    sizes = []
    b = 1024
    while b <= args.max_bytes:
        sizes.append(b)
        b *= 8
    if sizes[-1] != args.max_bytes:
        sizes.append(args.max_bytes)
    ops = [o for o in args.ops.split(",") if o]
    P = world
    import gc

    maxn = max(P, sizes[-1] // 4 // P * P)
    for op in ops:
        # One symmetric buffer per op, sliced per size: every allocation binds a multicast
        # object, so allocating per size would exhaust them (and waste memory).
        big = cc.empty(maxn, torch.float32)
        big.fill_(1)
        for total in sizes:
            # `total` = bytes each rank ends up holding (allgather/alltoall output, broadcast
            # buffer, reduce_scatter input), i.e. the size that bounds the traffic.
            n = max(P, total // 4 // P * P)  # float32 elements, multiple of P
            per = n // P
            iters = args.iters if total < (64 << 20) else max(5, args.iters // 4)
            row = {"op": op, "bytes": n * 4}
            try:
                if op == "allgather":
                    out = big[:n]
                    inp = torch.ones(per, device="cuda")
                    row["ours_us"] = timed(lambda: cc.allgather(out, inp, stream=stream), iters)
                    if nccl:
                        o2 = torch.empty(n, device="cuda")
                        row["nccl_us"] = timed(lambda: nccl.allgather(inp.data_ptr(), o2.data_ptr(), per, F32, stream.cuda_stream), iters)
                    factor = (P - 1) / P
                elif op in ("alltoall", "alltoall_v"):
                    out = big[:n]
                    inp = torch.ones(n, device="cuda")
                    if op == "alltoall":
                        row["ours_us"] = timed(lambda: cc.alltoall(out, inp, stream=stream), iters)
                    else:
                        # uneven but balanced split: rank r sends weight w[(j - r) % P] to rank j, so
                        # every rank sends and receives n elements in chunks of very different sizes.
                        w = [k + 1 for k in range(P)]
                        unit = [n * x // sum(w) for x in w]
                        unit[-1] += n - sum(unit)
                        send = [unit[(j - rank) % P] for j in range(P)]
                        recv = [unit[(rank - i) % P] for i in range(P)]
                        out_v = big[:n]
                        row["ours_us"] = timed(lambda: cc.alltoallv(out_v, recv, inp, send, stream=stream), iters)
                    if nccl and op == "alltoall":
                        o2 = torch.empty(n, device="cuda")
                        row["nccl_us"] = timed(lambda: nccl.alltoall(inp.data_ptr(), o2.data_ptr(), per, F32, stream.cuda_stream), iters)
                    factor = (P - 1) / P
                elif op == "broadcast":
                    buf = big[:n]
                    row["ours_us"] = timed(lambda: cc.broadcast(buf, root=0, stream=stream), iters)
                    if world > 2 and cc.nvls_available() and n * 4 >= (1 << 20):
                        for m in (0, 1):
                            os.environ["GLB_CUDA_BCAST_MODE"] = str(m)
                            row[f"ours_mode{m}_us"] = round(timed(lambda: cc.broadcast(buf, root=0, stream=stream), iters), 2)
                        os.environ.pop("GLB_CUDA_BCAST_MODE", None)
                    if nccl:
                        o2 = torch.empty(n, device="cuda")
                        row["nccl_us"] = timed(lambda: nccl.broadcast(o2.data_ptr(), o2.data_ptr(), n, F32, 0, stream.cuda_stream), iters)
                    factor = 1.0
                elif op == "reduce_scatter":
                    inp = big[:n]
                    out = torch.empty(per, device="cuda")
                    row["ours_us"] = timed(lambda: cc.reduce_scatter(out, inp, [per] * P, stream=stream), iters)
                    if world > 2 and cc.nvls_available():
                        gb._C.cuda.set_tuning({"nvls_reduce_scatter": False})
                        row["ours_p2p_us"] = round(timed(lambda: cc.reduce_scatter(out, inp, [per] * P, stream=stream), iters), 2)
                        gb._C.cuda.set_tuning({"nvls_reduce_scatter": True})
                    if nccl:
                        i2 = torch.ones(n, device="cuda")
                        row["nccl_us"] = timed(lambda: nccl.reduce_scatter(i2.data_ptr(), out.data_ptr(), per, F32, 1, stream.cuda_stream), iters)
                    factor = (P - 1) / P
                else:
                    continue
                row["ours_busbw_gbs"] = round(n * 4 / (row["ours_us"] * 1e-6) / 1e9 * factor, 2)
                if "nccl_us" in row:
                    row["nccl_busbw_gbs"] = round(n * 4 / (row["nccl_us"] * 1e-6) / 1e9 * factor, 2)
                row["ours_us"] = round(row["ours_us"], 2)
                if "nccl_us" in row:
                    row["nccl_us"] = round(row["nccl_us"], 2)
            except Exception as e:  # noqa: BLE001
                row["error"] = f"{type(e).__name__}: {str(e)[:160]}"
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
            torch.cuda.synchronize()
            gb.barrier(ctx)
        del big
        gc.collect()
    if args.schedules:
        for dtype, name in ((torch.float16, "float16"),):
            for n in (1 << 10, 1 << 16, 1 << 20, 1 << 24, 1 << 27):
                row = {"op": "allreduce_" + name, "elements": n, "bytes": n * 2}
                t = cc.empty(n, dtype)
                t.fill_(0)
                for label, cls, lit in (("auto", gcu.CudaAllreduceHalvingDoubling, False),
                                        ("halving_doubling", gcu.CudaAllreduceHalvingDoubling, True),
                                        ("bcube", gcu.CudaAllreduceBcube, True)):
                    algo = cls(ctx, t, streams=[stream], literal=lit)
                    row[label + "_us"] = round(timed(algo.run, args.iters if n < (1 << 24) else 6), 2)
                    if label == "auto":
                        row["auto_variant"] = algo.resolved_algo()
                if nccl:
                    row["nccl_us"] = round(timed(lambda: nccl.allreduce(t.data_ptr(), t.data_ptr(), n, int(gb.DataType.FLOAT16), 1, stream.cuda_stream), args.iters if n < (1 << 24) else 6), 2)
                best = min(row["auto_us"], row["halving_doubling_us"], row["bcube_us"])
                row["best_busbw_gbs"] = round(n * 2 / (best * 1e-6) / 1e9 * 2 * (P - 1) / P, 2)
                rows.append(row)
                if rank == 0:
                    print(json.dumps(row), flush=True)
                torch.cuda.synchronize()
                gb.barrier(ctx)
    if rank == 0 and args.out:
        with open(args.out, "w") as f:
            json.dump({"world": world, "describe": cc.describe(), "rows": rows}, f, indent=1)
    torch.cuda.synchronize()
    gb.barrier(ctx)
    ctx.close_connections()


if __name__ == "__main__":
    main()
