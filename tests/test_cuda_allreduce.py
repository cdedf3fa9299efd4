"""CUDA allreduce over peer memory — mirrors gloo/test/cuda_allreduce_test.cc:148-349.
Ranks are threads that share cuda:0 (kernels are co-resident: the grid is capped by
the number of ranks per device), plus multi-process runs when >1 GPU is present."""
import pytest
import torch

import gloo_b200 as gb
from gloo_b200.ops import cuda as gcu

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.float16, torch.bfloat16, torch.int32, torch.int64, torch.float64, torch.uint8]


def _expected(size, count, dtype):
    # rank r contributes j*size + r  ->  sum = j*size^2 + size(size-1)/2
    j = torch.arange(count, dtype=torch.float64)
    return j * size * size + size * (size - 1) / 2


def _input(rank, size, count, dtype, device):
    j = torch.arange(count, dtype=torch.float64)
    return (j * size + rank).to(dtype).to(device)


def _tol(dtype):
    return {torch.float16: 1e-2, torch.bfloat16: 3e-2}.get(dtype, 1e-5)


@pytest.mark.parametrize("size", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("algo", ["auto", "one_shot", "two_shot"])
def test_allreduce_registered(size, algo):
    counts = [1, 3, 8, 100, 1000, 4099, 65536, 300000]

    def fn(ctx):
        cc = gcu.CudaContext(ctx, 0, stage_bytes=8 << 20)
        for count in counts:
            if algo == "one_shot" and count * 4 > 256 * 1024:
                continue
            t = _input(ctx.rank, size, count, torch.float32, "cuda:0")
            cc.register(t)
            cc.allreduce(t, algo=algo)
            torch.cuda.current_stream().synchronize()
            torch.testing.assert_close(t.double().cpu(), _expected(size, count, torch.float32), rtol=1e-5, atol=0)
        cc.pc.host_barrier()
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


@pytest.mark.parametrize("dtype", DTYPES)
def test_allreduce_staged_dtypes(dtype):
    size = 4

    def fn(ctx):
        cc = gcu.CudaContext(ctx, 0, stage_bytes=4 << 20)
        for count in [5, 257, 70001, 700001]:
            small = 7 if dtype in (torch.uint8,) else 50
            t = ((torch.arange(count, dtype=torch.float64) % small) + ctx.rank).to(dtype).cuda()
            exp = sum(((torch.arange(count, dtype=torch.float64) % small) + r).to(dtype).double() for r in range(size))
            cc.allreduce(t)  # unregistered: staged through the pool (one-shot or piecewise two-shot)
            torch.cuda.current_stream().synchronize()
            torch.testing.assert_close(t.double().cpu(), exp.to(dtype).double(), rtol=_tol(dtype), atol=_tol(dtype))
        cc.pc.host_barrier()
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


@pytest.mark.parametrize("op", [gb.ReduceOp.SUM, gb.ReduceOp.PRODUCT, gb.ReduceOp.MIN, gb.ReduceOp.MAX])
def test_allreduce_ops(op):
    size = 3

    def fn(ctx):
        cc = gcu.CudaContext(ctx, 0, stage_bytes=4 << 20)
        g = torch.Generator().manual_seed(7)
        data = [torch.randint(1, 4, (100003,), generator=g).float() for _ in range(size)]
        f = {gb.ReduceOp.SUM: torch.add, gb.ReduceOp.PRODUCT: torch.mul, gb.ReduceOp.MIN: torch.minimum,
             gb.ReduceOp.MAX: torch.maximum}[op]
        exp = data[0]
        for d in data[1:]:
            exp = f(exp, d)
        for n in (1000, 100003):
            t = data[ctx.rank][:n].cuda()
            cc.register(t)
            cc.allreduce(t, op=op)
            torch.cuda.current_stream().synchronize()
            torch.testing.assert_close(t.cpu(), exp[:n])
        cc.pc.host_barrier()
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


def test_old_style_classes_and_streams():
    size = 2
    names = [gcu.CudaAllreduceRing, gcu.CudaAllreduceRingChunked, gcu.CudaAllreduceHalvingDoubling,
             gcu.CudaAllreduceHalvingDoublingPipelined, gcu.CudaAllreduceBcube]

    def fn(ctx):
        streams = [gcu.new_stream(0) for _ in range(2)]  # dedicated: torch's pool could alias ranks
        for cls in names:
            for count in (100, 200000):
                # MultiPointer: two local buffers per rank -> 2*size contributions
                ptrs = 2
                stride = size * ptrs
                ts = [(torch.arange(count, dtype=torch.float64) * stride + ctx.rank * ptrs + i).float().cuda()
                      for i in range(ptrs)]
                algo = cls(ctx, ts)
                assert algo.uses_peer_memory()
                algo.run()
                exp = (torch.arange(count, dtype=torch.float64) * stride * stride + stride * (stride - 1) / 2)
                for t in ts:
                    torch.testing.assert_close(t.double().cpu(), exp, rtol=1e-5, atol=0)
                # async variant: user streams + delayed initialisation
                ts2 = [torch.empty(count, device="cuda") for _ in range(ptrs)]
                algo2 = cls(ctx, ts2, streams=streams)
                for i, s in enumerate(streams):
                    gb._C.cuda.spin(200000, s.cuda_stream)
                    gb._C.cuda.fill(ts2[i].data_ptr(), count, int(gb.DataType.FLOAT32), float(ctx.rank * ptrs + i),
                                    float(stride), s.cuda_stream)
                algo2.run()
                for s in streams:
                    s.synchronize()
                for t in ts2:
                    torch.testing.assert_close(t.double().cpu(), exp, rtol=1e-5, atol=0)
        gcu._cu.peer_context_for(ctx, 0).host_barrier()
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


def test_host_workspace_fallback():
    size = 2

    def fn(ctx):
        t = _input(ctx.rank, size, 5000, torch.float32, "cuda:0")
        algo = gcu.CudaAllreduceRingChunked(ctx, t, host_workspace=True)
        assert not algo.uses_peer_memory()
        algo.run()
        torch.testing.assert_close(t.double().cpu(), _expected(size, 5000, torch.float32), rtol=1e-5, atol=0)
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


def test_symmetric_tensor_and_barrier():
    size = 2

    def fn(ctx):
        cc = gcu.CudaContext(ctx, 0, stage_bytes=4 << 20)
        t = cc.empty(4096, torch.float32)
        t.copy_(_input(ctx.rank, size, 4096, torch.float32, "cuda:0"))
        cc.barrier()
        cc.allreduce(t, algo="two_shot")
        torch.cuda.current_stream().synchronize()
        torch.testing.assert_close(t.double().cpu(), _expected(size, 4096, torch.float32), rtol=1e-5, atol=0)
        cc.pc.host_barrier()
        return cc.describe()

    res = gb.spawn_threads(size, fn, cuda_device=0)
    assert all("PeerContext" in r for r in res)


@pytest.mark.parametrize("size", [2, 3, 4, 7, 8])
def test_literal_schedules(size):
    """The named schedules executed literally over peer pointers (one kernel each)."""
    classes = [gcu.CudaAllreduceRing, gcu.CudaAllreduceRingChunked, gcu.CudaAllreduceHalvingDoubling,
               gcu.CudaAllreduceBcube]

    def fn(ctx):
        for cls in classes:
            for dtype in (torch.float32, torch.float16):
                for count in (1, 100, 4099, 300000):
                    small = 8
                    t = ((torch.arange(count, dtype=torch.float64) % small) + ctx.rank).to(dtype).cuda()
                    algo = cls(ctx, t, literal=True)
                    assert algo.resolved_algo() in ("ring", "ring_chunked", "halving_doubling", "bcube")
                    algo.run()
                    exp = (torch.arange(count, dtype=torch.float64) % small) * size + size * (size - 1) / 2
                    torch.testing.assert_close(t.double().cpu(), exp, rtol=1e-3, atol=1e-3)
        gcu._cu.peer_context_for(ctx, 0).host_barrier()
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


@pytest.mark.parametrize("size,inputs", [(1, 2), (1, 1), (2, 1), (3, 2)])
def test_host_allreduce_pipeline(size, inputs):
    """CudaHostAllreduce: pinned host -> (H2D | fused allreduce | D2H pipeline) -> pinned host,
    checked against an fp64 sum on the host; run twice to cover buffer reuse across runs and
    with a piece count that does not divide the length."""
    count = 300_003

    def fn(ctx):
        cc = gcu.CudaContext(ctx, 0, stage_bytes=8 << 20)
        hins = [torch.empty(count, dtype=torch.float32).pin_memory() for _ in range(inputs)]
        hout = torch.empty(count, dtype=torch.float32).pin_memory()
        op = gcu.CudaHostAllreduce(ctx, cc, hins, hout, chunks=5)
        assert len(op.bounds) >= 5 and op.bounds[-1][1] == count
        base = torch.arange(count, dtype=torch.float64) % 1000
        for it in range(2):
            for i, h in enumerate(hins):
                h.copy_((base + (ctx.rank * inputs + i) * (it + 1)).float())
            hout.fill_(-1.0)
            op.run()
            torch.cuda.current_stream().synchronize()
            total = size * inputs
            want = base * total + (it + 1) * total * (total - 1) / 2
            torch.testing.assert_close(hout.double(), want, rtol=1e-6, atol=0)
        cc.pc.host_barrier()
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))
