"""CUDA data-movement collectives over peer memory (net-new vs the reference: it has no
CUDA broadcast-by-kernel / allgather / alltoall / reduce_scatter / reduce / gather /
scatter). Registered (zero-copy, symmetric) and staged (plain tensors) flavours."""
import pytest
import torch

import gloo_b200 as gb
from gloo_b200.ops import cuda as gcu

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 3, 4, 8]


def _ctx(ctx):
    return gcu.CudaContext(ctx, 0, stage_bytes=16 << 20)


@pytest.mark.parametrize("size", SIZES)
def test_broadcast(size):
    def fn(ctx):
        cc = _ctx(ctx)
        for n in (1, 1000, 70001, 1 << 20):   # direct, and scatter+allgather for the large one
            for root in {0, size - 1}:
                t = torch.full((n,), float(ctx.rank), device="cuda")
                if ctx.rank == root:
                    t.copy_(torch.arange(n, dtype=torch.float32) + root)
                cc.broadcast(t, root=root)                      # staged
                s = cc.empty(n, torch.float32)                  # symmetric
                s.fill_(float(ctx.rank))
                if ctx.rank == root:
                    s.copy_(torch.arange(n, dtype=torch.float32) + root)
                cc.broadcast(s, root=root)
                torch.cuda.current_stream().synchronize()
                exp = torch.arange(n, dtype=torch.float32) + root
                torch.testing.assert_close(t.cpu(), exp)
                torch.testing.assert_close(s.cpu(), exp)
        cc.pc.host_barrier()
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


@pytest.mark.parametrize("size", SIZES)
def test_allgather_and_v(size):
    def fn(ctx):
        cc = _ctx(ctx)
        for n in (1, 5, 4096, 100003):
            inp = torch.arange(n, dtype=torch.float32, device="cuda") + 1000 * ctx.rank
            exp = torch.cat([torch.arange(n, dtype=torch.float32) + 1000 * r for r in range(size)])
            out = torch.zeros(n * size, device="cuda")
            cc.allgather(out, inp)                              # staged
            sym = cc.empty(n * size, torch.float32)
            cc.allgather(sym, inp)                              # registered
            torch.cuda.current_stream().synchronize()
            torch.testing.assert_close(out.cpu(), exp)
            torch.testing.assert_close(sym.cpu(), exp)
        counts = [(r * 7) % 5 for r in range(size)]            # includes empty ranks
        inp = torch.full((max(counts[ctx.rank], 1),), float(ctx.rank), device="cuda")[:counts[ctx.rank]]
        out = torch.full((max(sum(counts), 1),), -1.0, device="cuda")[:sum(counts)]
        cc.allgatherv(out, inp, counts)
        torch.cuda.current_stream().synchronize()
        exp = torch.cat([torch.full((counts[r],), float(r)) for r in range(size)]) if sum(counts) else out.cpu()
        torch.testing.assert_close(out.cpu(), exp)
        cc.pc.host_barrier()
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


@pytest.mark.parametrize("size", SIZES)
def test_alltoall_and_v(size):
    def fn(ctx):
        cc = _ctx(ctx)
        r = ctx.rank
        for n in (1, 33, 50000):
            inp = torch.cat([torch.full((n,), float(r * 100 + j)) for j in range(size)]).cuda()
            exp = torch.cat([torch.full((n,), float(j * 100 + r)) for j in range(size)])
            out = torch.zeros(n * size, device="cuda")
            cc.alltoall(out, inp)
            sym = cc.empty(n * size, torch.float32)
            cc.alltoall(sym, inp)
            torch.cuda.current_stream().synchronize()
            torch.testing.assert_close(out.cpu(), exp)
            torch.testing.assert_close(sym.cpu(), exp)
        send = [((r + j) % 3) * 5 + (1 if j == r else 0) for j in range(size)]
        recv = [((j + r) % 3) * 5 + (1 if j == r else 0) for j in range(size)]
        inp = torch.cat([torch.full((send[j],), float(r * 100 + j)) for j in range(size)]).cuda()
        out = torch.full((sum(recv),), -1.0, device="cuda")
        cc.alltoallv(out, recv, inp, send)                      # receive offsets exchanged in-kernel
        torch.cuda.current_stream().synchronize()
        exp = torch.cat([torch.full((recv[j],), float(j * 100 + r)) for j in range(size)])
        torch.testing.assert_close(out.cpu(), exp)
        cc.pc.host_barrier()
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.int32])
def test_reduce_scatter_and_reduce(size, dtype):
    def fn(ctx):
        cc = _ctx(ctx)
        for n in (size, 1000, 100003):
            base = (torch.arange(n, dtype=torch.float64) % 16)
            inp = (base + ctx.rank).to(dtype).cuda()
            full = (base * size + size * (size - 1) / 2)
            b, rem = divmod(n, size)
            counts = [b + (1 if r < rem else 0) for r in range(size)]
            off = sum(counts[:ctx.rank])
            out = torch.zeros(counts[ctx.rank], dtype=dtype, device="cuda")
            cc.reduce_scatter(out, inp, counts)                 # staged input
            sym = cc.empty(n, dtype)
            sym.copy_(inp)
            out2 = torch.zeros(counts[ctx.rank], dtype=dtype, device="cuda")
            cc.reduce_scatter(out2, sym, counts)                # registered input
            red = torch.zeros(n, dtype=dtype, device="cuda")
            cc.reduce(red, inp, root=size - 1)
            torch.cuda.current_stream().synchronize()
            exp = full[off:off + counts[ctx.rank]]
            torch.testing.assert_close(out.double().cpu(), exp, rtol=2e-2, atol=1e-2)
            torch.testing.assert_close(out2.double().cpu(), exp, rtol=2e-2, atol=1e-2)
            if ctx.rank == size - 1:
                torch.testing.assert_close(red.double().cpu(), full, rtol=2e-2, atol=1e-2)
        cc.pc.host_barrier()
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


@pytest.mark.parametrize("size", SIZES)
def test_gather_scatter(size):
    def fn(ctx):
        cc = _ctx(ctx)
        n = 777
        for root in {0, size - 1}:
            inp = torch.arange(n, dtype=torch.float32, device="cuda") + ctx.rank
            out = torch.zeros(n * size, device="cuda")
            cc.gather(out, inp, root=root)
            src = torch.cat([torch.arange(n, dtype=torch.float32) * (j + 1) for j in range(size)]).cuda()
            got = torch.zeros(n, device="cuda")
            cc.scatter(got, src if ctx.rank == root else None, root=root)
            torch.cuda.current_stream().synchronize()
            if ctx.rank == root:
                exp = torch.cat([torch.arange(n, dtype=torch.float32) + r for r in range(size)])
                torch.testing.assert_close(out.cpu(), exp)
            torch.testing.assert_close(got.cpu(), torch.arange(n, dtype=torch.float32) * (ctx.rank + 1))
        cc.pc.host_barrier()
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


def test_broadcast_class_and_topology():
    size = 3

    def fn(ctx):
        ts = [torch.full((5000,), float(ctx.rank * 10 + i), device="cuda") for i in range(2)]
        b = gcu.CudaBroadcastOneToAll(ctx, ts, root=1, root_pointer=1)
        b.run()
        for t in ts:
            assert float(t[0]) == 11.0 and float(t[-1]) == 11.0
        pc = gcu._cu.peer_context_for(ctx, 0)
        topo = pc.topology()
        assert len(topo) == size and all(d.sm_count > 0 and d.cc_major >= 9 for d in topo)
        assert pc.ranks_on_my_device() == size and pc.peer_access_everywhere()
        pc.host_barrier()
        return pc.describe()

    assert all("ranksOnDevice=3" in d for d in gb.spawn_threads(size, fn, cuda_device=0))
