#!/usr/bin/env python
"""Long-context micro-benchmarks on the NVLink peer-memory path:

  * ring KV rotation (ring attention): every rank streams its KV block [S/P, H, D] (bf16) to
    the right neighbour while receiving the left neighbour's block - ``RingExchange.rotate`` =
    one fused send+recv kernel (cuda/p2p_kernels.cu);
  * Ulysses head scatter: [S/P, H, D] -> [S, H/P, D], one alltoall each way
    (``UlyssesAttention.seq_to_heads``).

Both against NCCL (grouped ncclSend/ncclRecv) on the same buffers. Device timed, max over
ranks, p50 of the iterations. Launch with torchrun, one rank per GPU.
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
from gloo_b200.parallel.strategies import RingExchange, UlyssesAttention  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--head-dim", type=int, default=128)
    ap.add_argument("--seq", default="8192,32768,131072,524288", help="global sequence lengths")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    ctx = gb.init_context(rank, world, path=f"/tmp/glb_lc_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}", timeout_ms=120000)
    cc = gcu.CudaContext(ctx, local, stage_bytes=256 << 20)
    nccl = None
    try:
        nccl = gb._C.cuda.NcclComm.init_rank(ctx, local)
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print("NCCL comparator unavailable:", e)
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    BF16, U8 = int(gb.DataType.BFLOAT16), int(gb.DataType.UINT8)

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

    ring, uly = RingExchange(ctx, cc), UlyssesAttention(ctx, cc)
    right, left = (rank + 1) % world, (rank - 1) % world
    rows = []
    for S in [int(x) for x in args.seq.split(",")]:
        shard = S // world
        kv = torch.randn(2, shard, args.heads, args.head_dim, device="cuda").to(torch.bfloat16)  # K and V blocks
        nxt = torch.empty_like(kv)
        nbytes = kv.numel() * 2
        row = {"seq": S, "shard_tokens": shard, "kv_block_bytes": nbytes}
        with torch.cuda.stream(stream):
            us = timed(lambda: ring.rotate(kv, nxt), nbytes)
        row["ring_rotate_us"] = round(us, 2)
        row["ring_rotate_gbs_per_dir"] = round(nbytes / (us * 1e-6) / 1e9, 1)
        stream.synchronize()
        gb.barrier(ctx)
        src = torch.full_like(kv, float(rank))
        with torch.cuda.stream(stream):
            ring.rotate(src, nxt)
        stream.synchronize()
        assert float(nxt.flatten()[0]) == left and float(nxt.flatten()[-1]) == left, "ring rotation delivered the wrong block"
        # the same rotation between two symmetric buffers: zero-copy exchange kernel
        ka, kb = ring.kv_buffers(kv.numel(), torch.bfloat16)
        ka.copy_(kv.view(-1))
        with torch.cuda.stream(stream):
            us = timed(lambda: ring.rotate(ka, kb), nbytes)
        row["ring_rotate_symmetric_us"] = round(us, 2)
        row["ring_rotate_symmetric_gbs_per_dir"] = round(nbytes / (us * 1e-6) / 1e9, 1)
        stream.synchronize()
        gb.barrier(ctx)
        ka.fill_(float(rank))
        torch.cuda.synchronize()
        gb.barrier(ctx)
        with torch.cuda.stream(stream):
            ring.rotate(ka, kb)
        stream.synchronize()
        assert float(kb[0]) == left and float(kb[-1]) == left, "symmetric ring rotation delivered the wrong block"
        gb.barrier(ctx)
        if nccl is not None:
            us = timed(lambda: nccl.sendrecv(kv.data_ptr(), right, nxt.data_ptr(), left, kv.numel(), BF16, stream.cuda_stream), nbytes)
            row["nccl_ring_rotate_us"] = round(us, 2)
            row["nccl_ring_rotate_gbs_per_dir"] = round(nbytes / (us * 1e-6) / 1e9, 1)
        # Ulysses: q/k/v projections of this shard, heads scattered across ranks
        x = torch.randn(shard, args.heads, args.head_dim, device="cuda").to(torch.bfloat16)
        xin = x.view(shard, world, args.heads // world, args.head_dim).transpose(0, 1).contiguous()  # [P, S/P, H/P, D]
        out = cc.empty(xin.numel(), torch.bfloat16)
        abytes = xin.numel() * 2
        with torch.cuda.stream(stream):
            us = timed(lambda: uly.seq_to_heads(xin.view(-1), out), abytes)
        row["ulysses_alltoall_us"] = round(us, 2)
        row["ulysses_busbw_gbs"] = round(abytes / (us * 1e-6) / 1e9 * (world - 1) / world, 1)
        if nccl is not None:
            us = timed(lambda: nccl.alltoall(xin.data_ptr(), out.data_ptr(), xin.numel() * 2 // world, U8, stream.cuda_stream), abytes)
            row["nccl_ulysses_alltoall_us"] = round(us, 2)
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        del kv, nxt, src, x, xin, out, ka, kb
        torch.cuda.synchronize()
        gb.barrier(ctx)
    if rank == 0 and args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump({"world": world, "heads": args.heads, "head_dim": args.head_dim, "describe": cc.describe(), "rows": rows}, f, indent=1)
    ctx.close_connections()


if __name__ == "__main__":
    main()
