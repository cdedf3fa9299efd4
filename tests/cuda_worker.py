"""One process per GPU: exercises the cross-process peer-memory path (cuMem fd passing or
cudaIpc, NVLS multicast when available). usage: cuda_worker.py STORE_DIR RANK SIZE"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import gloo_b200 as gb  # noqa: E402
from gloo_b200.ops import cuda as gcu  # noqa: E402


def main():
    store_dir, rank, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    ctx = gb.init_context(rank, size, path=store_dir, timeout_ms=60000)
    cc = gcu.CudaContext(ctx, dev, stage_bytes=32 << 20)
    if rank == 0:
        print("DESCRIBE", cc.describe(), flush=True)

    def exp_sum(n):
        return torch.arange(n, dtype=torch.float64) * size * size + size * (size - 1) / 2

    def inp(n, dtype=torch.float32):
        return (torch.arange(n, dtype=torch.float64) * size + rank).to(dtype).cuda()

    # allreduce: staged, registered, symmetric (NVLS when bound), every variant
    for n in (1, 1000, 65536, 1 << 20, 5_000_001):
        t = inp(n)
        cc.allreduce(t)
        torch.cuda.synchronize()
        torch.testing.assert_close(t.double().cpu(), exp_sum(n), rtol=1e-5, atol=0)
        t = inp(n)
        cc.register(t)
        for algo in ("one_shot", "two_shot"):
            if algo == "one_shot" and n * 4 > 256 * 1024:
                continue
            t.copy_(inp(n))
            cc.allreduce(t, algo=algo)
            torch.cuda.synchronize()
            torch.testing.assert_close(t.double().cpu(), exp_sum(n), rtol=1e-5, atol=0)
        s = cc.empty(n, torch.float32)
        s.copy_(inp(n))
        algos = ["auto", "two_shot"] + (["nvls"] if cc.nvls_available() else [])
        for algo in algos:
            s.copy_(inp(n))
            cc.allreduce(s, algo=algo)
            torch.cuda.synchronize()
            torch.testing.assert_close(s.double().cpu(), exp_sum(n), rtol=1e-5, atol=0)
    # half precision through NVLS / two-shot
    for dtype in (torch.float16, torch.bfloat16):
        n = 1 << 20
        s = cc.empty(n, dtype)
        s.copy_(((torch.arange(n) % 16) + rank).to(dtype))
        cc.allreduce(s)
        torch.cuda.synchronize()
        torch.testing.assert_close(s.double().cpu(), ((torch.arange(n) % 16) * size + size * (size - 1) / 2).double(),
                                   rtol=1e-2, atol=1e-2)
    # old-style class + literal schedules
    for cls in (gcu.CudaAllreduceRingChunked, gcu.CudaAllreduceHalvingDoubling, gcu.CudaAllreduceBcube, gcu.CudaAllreduceRing):
        for literal in (False, True):
            t = inp(300000)
            a = cls(ctx, t, literal=literal)
            a.run()
            torch.testing.assert_close(t.double().cpu(), exp_sum(300000), rtol=1e-5, atol=0)
    # data movement
    n = 100003
    b = torch.full((n,), float(rank), device="cuda")
    cc.broadcast(b, root=size - 1)
    sb = cc.empty(1 << 20, torch.float32)
    sb.fill_(float(rank))
    cc.broadcast(sb, root=0)
    out = torch.zeros(n * size, device="cuda")
    cc.allgather(out, torch.full((n,), float(rank), device="cuda"))
    a2a_in = torch.cat([torch.full((n,), float(rank * 100 + j)) for j in range(size)]).cuda()
    a2a = torch.zeros(n * size, device="cuda")
    cc.alltoall(a2a, a2a_in)
    base, rem = divmod(n, size)
    counts = [base + (1 if r < rem else 0) for r in range(size)]
    rs = torch.zeros(counts[rank], device="cuda")
    cc.reduce_scatter(rs, inp(n), counts)
    red = torch.zeros(n, device="cuda")
    cc.reduce(red, inp(n), root=0)
    torch.cuda.synchronize()
    assert float(b[0]) == size - 1 and float(sb[-1]) == 0.0
    torch.testing.assert_close(out.cpu(), torch.arange(size).repeat_interleave(n).float())
    torch.testing.assert_close(a2a.cpu(), torch.cat([torch.full((n,), float(j * 100 + rank)) for j in range(size)]))
    off = sum(counts[:rank])
    torch.testing.assert_close(rs.double().cpu(), exp_sum(n)[off:off + counts[rank]], rtol=1e-5, atol=0)
    if rank == 0:
        torch.testing.assert_close(red.double().cpu(), exp_sum(n), rtol=1e-5, atol=0)
    # ---- round 2: fused epilogues, LL, pipelined plain pointers, p2p, one-sided, CUDA graphs ----
    for algo in ["ll", "one_shot", "two_shot", "pipelined"] + (["nvls"] if cc.nvls_available() else []):
        n = 1000 if algo in ("ll", "one_shot") else 3_000_001
        t = cc.empty(n, torch.float32) if algo in ("nvls", "two_shot") else torch.empty(n, device="cuda")
        t.copy_(inp(n))
        cc.allreduce(t, algo=algo, average=True)
        cc.synchronize()
        torch.testing.assert_close(t.double().cpu(), exp_sum(n) / size, rtol=1e-5, atol=1e-6)
    a32, o16 = cc.empty(1 << 20, torch.float32), cc.empty(1 << 20, torch.bfloat16)
    a32.fill_(float(rank + 1))
    cc.allreduce(a32, out=o16, average=True)   # NVLS ld_reduce + cast when multicast-bound
    cc.synchronize()
    assert float(o16[0]) == (size + 1) / 2 and float(o16[-1]) == (size + 1) / 2
    # multi-pointer class: fold + exchange + fan-out in one launch
    ts = [cc.empty(200_000, torch.float32), torch.empty(200_000, device="cuda")]
    for i, t in enumerate(ts):
        t.fill_(float(rank * 2 + i))
    mp = gcu.CudaAllreduceRingChunked(ctx, ts)
    assert mp.launches_per_run() == 1
    mp.run()
    tot = 2 * size
    assert float(ts[0][0]) == tot * (tot - 1) / 2 and float(ts[1][-1]) == tot * (tot - 1) / 2
    # point to point ring + one-sided
    right, left = (rank + 1) % size, (rank - 1) % size
    s_t, r_t = torch.full((5_000_001,), float(rank), device="cuda"), torch.zeros(5_000_001, device="cuda")
    cc.sendrecv(s_t, right, r_t, left)
    win = cc.empty(1024, torch.float32)
    win.fill_(-1.0)
    cc.synchronize()
    cc.pc.host_barrier()
    cc.put(torch.full((8,), float(rank), device="cuda"), win, right, remote_offset=8 * rank)
    cc.synchronize()
    cc.barrier()
    cc.synchronize()
    assert float(r_t[0]) == left and float(r_t[-1]) == left and float(win[8 * left]) == left
    # zero-copy exchange into a symmetric double buffer: three ring steps, alternating halves
    m = 2_000_003
    kv = cc.empty(2 * m, torch.float32)
    kv[:m].fill_(float(rank))
    cc.synchronize()
    cc.pc.host_barrier()
    for step in range(3):
        cur, nxt = kv[(step % 2) * m:(step % 2 + 1) * m], kv[((step + 1) % 2) * m:((step + 1) % 2 + 1) * m]
        cc.exchange(cur, right, nxt, left)
        cc.synchronize()
        assert float(nxt[0]) == (rank - step - 1) % size and float(nxt[-1]) == (rank - step - 1) % size, (step, float(nxt[0]))
    # NVLS + peer-to-peer hybrid
    if cc.nvls_available() and size in (2, 4, 8):
        for n, split in ((3_000_001, 300), (5_000_000, 0), (40_000, 900)):
            t = cc.empty(n, torch.float32)
            t.copy_(inp(n))
            cc.allreduce(t, algo="hybrid", average=True, blocks=48, unroll=16, tile=split)
            cc.synchronize()
            torch.testing.assert_close(t.double().cpu(), exp_sum(n) / size, rtol=1e-5, atol=1e-6)
    # CUDA graph: fill + three allreduces captured once, replayed
    gs = torch.cuda.Stream()
    gt = cc.empty(1 << 18, torch.float32)
    galgo = gcu.CudaAllreduceRingChunked(ctx, gt, streams=[gs])
    small = torch.empty(256, device="cuda")
    F32 = int(gb.DataType.FLOAT32)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=gs):
        gb._C.cuda.fill(gt.data_ptr(), gt.numel(), F32, float(rank), 0.0, gs.cuda_stream)
        gb._C.cuda.fill(small.data_ptr(), 256, F32, 1.0, 0.0, gs.cuda_stream)
        galgo.run()
        cc.allreduce(small, stream=gs)          # LL kernel inside the graph
        galgo.run()
    for _ in range(4):
        graph.replay()
    torch.cuda.synchronize()
    cc.check_health()
    tri = size * (size - 1) / 2
    assert float(gt[0]) == tri * size and float(gt[-1]) == tri * size and float(small[7]) == size
    cc.pc.host_barrier()
    print(f"WORKER {rank} OK", flush=True)


if __name__ == "__main__":
    main()
