#!/usr/bin/env python
"""Device-timed sweep of every allreduce variant (and the NCCL comparator) across message
sizes. Launch with torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29511 scripts/sweep_allreduce.py --out gpurun_out/sweep8.json

Prints one table (p50 us per variant, best variant, bus GB/s) and writes JSON.
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
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--max-elements", type=int, default=100_000_000)
    ap.add_argument("--variants", default="one_shot,two_shot,nvls,staged,ring_chunked,halving_doubling,bcube,nccl")
    ap.add_argument("--blocks", default="")
    ap.add_argument("--sizes", default="", help="comma separated element counts (default: built-in sweep)")
    ap.add_argument("--tune-blocks", default="", help="e.g. 32,64,128: extra pass over --tune-sizes per block count")
    ap.add_argument("--tune-sizes", default="262144,4000000,100000000")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dtype = getattr(torch, args.dtype)
    es = torch.empty((), dtype=dtype).element_size()
    path = f"/tmp/glb_sweep_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
    ctx = gb.init_context(rank, world, path=path, timeout_ms=120000)
    cc = gcu.CudaContext(ctx, local, stage_bytes=256 << 20)
    if rank == 0:
        print(cc.describe(), flush=True)
    if args.blocks:
        gb._C.cuda.set_tuning({"max_blocks": int(args.blocks)})
    variants = [v for v in args.variants.split(",") if v]
    nccl = None
    if "nccl" in variants:
        try:
            nccl = gb._C.cuda.NcclComm.init_rank(ctx, local)
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print("NCCL comparator unavailable:", e)
            variants.remove("nccl")
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sizes = []
    n = 1
    while n <= args.max_elements:
        sizes += [n] if n == 1 else [n]
        n *= 10
    sizes = sorted(set(sizes + [x for x in (2048, 16384, 65536, 262144, 524288, 2_000_000, 5_000_000, 30_000_000)
                                if x <= args.max_elements]))
    if args.sizes:
        sizes = [int(x) for x in args.sizes.split(",")]
    rows = []
    dt_code = {torch.float32: 5, torch.float16: 7, torch.bfloat16: 8}[dtype]

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
        return float(per[len(per) // 2]), float(per[min(len(per) - 1, int(len(per) * 0.99))])

    for n in sizes:
        nbytes = n * es
        sym = cc.empty(n, dtype)
        sym.fill_(1)
        plain = torch.ones(n, dtype=dtype, device="cuda")
        reg = torch.ones(n, dtype=dtype, device="cuda")
        cc.register(reg)
        iters = args.iters if nbytes < (64 << 20) else max(8, args.iters // 3)
        row = {"elements": n, "bytes": nbytes}
        for v in variants:
            try:
                if v == "one_shot":
                    if nbytes > (256 << 10):
                        continue
                    fn = lambda: cc.allreduce(reg, algo="one_shot", stream=stream)  # noqa: E731
                elif v == "two_shot":
                    fn = lambda: cc.allreduce(reg, algo="two_shot", stream=stream)  # noqa: E731
                elif v == "nvls":
                    if not cc.nvls_available():
                        continue
                    fn = lambda: cc.allreduce(sym, algo="nvls", stream=stream)  # noqa: E731
                elif v == "staged":
                    fn = lambda: cc.allreduce(plain, stream=stream)  # noqa: E731
                elif v in ("ring_chunked", "halving_doubling", "bcube", "ring"):
                    cls = {"ring_chunked": gcu.CudaAllreduceRingChunked, "halving_doubling": gcu.CudaAllreduceHalvingDoubling,
                           "bcube": gcu.CudaAllreduceBcube, "ring": gcu.CudaAllreduceRing}[v]
                    algo = cls(ctx, reg, streams=[stream], literal=True)
                    fn = algo.run
                elif v == "nccl":
                    fn = lambda: nccl.allreduce(plain.data_ptr(), plain.data_ptr(), n, dt_code, 1, stream.cuda_stream)  # noqa: E731
                else:
                    continue
                p50, p99 = timed(fn, iters)
                row[v] = round(p50, 2)
                row[v + "_p99"] = round(p99, 2)
            except Exception as e:  # noqa: BLE001
                row[v] = None
                if rank == 0:
                    print(f"  {v} @ {n}: {type(e).__name__}: {str(e)[:200]}", flush=True)
        cand = {k: row[k] for k in variants if row.get(k)}
        if cand:
            ours = {k: t for k, t in cand.items() if k != "nccl"}
            best = min(ours, key=ours.get) if ours else None
            row["best"] = best
            if best:
                algbw = nbytes / (ours[best] * 1e-6) / 1e9
                row["best_busbw_gbs"] = round(algbw * 2 * (world - 1) / world, 2) if world > 1 else None
            if "nccl" in cand:
                row["nccl_busbw_gbs"] = round(nbytes / (cand["nccl"] * 1e-6) / 1e9 * 2 * (world - 1) / world, 2) if world > 1 else None
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        del sym, plain, reg
        torch.cuda.synchronize()
        gb.barrier(ctx)
    tune = []
    if args.tune_blocks:
        for n in [int(x) for x in args.tune_sizes.split(",")]:
            nbytes = n * es
            sym = cc.empty(n, dtype)
            sym.fill_(1)
            reg = torch.ones(n, dtype=dtype, device="cuda")
            cc.register(reg)
            for nb in [int(x) for x in args.tune_blocks.split(",")]:
                gb._C.cuda.set_tuning({"max_blocks": nb, "one_shot_blocks": min(nb, 32)})
                row = {"elements": n, "bytes": nbytes, "blocks": nb}
                iters = 20 if nbytes < (64 << 20) else 8
                if nbytes <= (256 << 10):
                    row["one_shot"] = round(timed(lambda: cc.allreduce(reg, algo="one_shot", stream=stream), iters)[0], 2)
                row["two_shot"] = round(timed(lambda: cc.allreduce(reg, algo="two_shot", stream=stream), iters)[0], 2)
                if cc.nvls_available():
                    row["nvls"] = round(timed(lambda: cc.allreduce(sym, algo="nvls", stream=stream), iters)[0], 2)
                tune.append(row)
                if rank == 0:
                    print("TUNE", json.dumps(row), flush=True)
            torch.cuda.synchronize()
            gb.barrier(ctx)
    if rank == 0 and args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump({"world": world, "dtype": args.dtype, "describe": cc.describe(), "rows": rows, "tune": tune}, f, indent=1)
    torch.cuda.synchronize()
    gb.barrier(ctx)
    ctx.close_connections()


if __name__ == "__main__":
    main()
