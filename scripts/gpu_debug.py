#!/usr/bin/env python
"""Run small GPU scenarios, each in its own subprocess with a timeout, and report
PASS / FAIL / HANG per scenario (a hung collective kernel must not take the rest down).
usage: python scripts/gpu_debug.py [case ...]"""
import os
import subprocess
import sys
import time

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run_case(name):
    import torch

    import gloo_b200 as gb
    from gloo_b200.ops import cuda as gcu

    size = int(os.environ.get("DBG_SIZE", "2"))

    def log(ctx, *a):
        print(f"[{name} r{ctx.rank}]", *a, flush=True)

    def cc_of(ctx):
        return gcu.CudaContext(ctx, 0, stage_bytes=8 << 20)

    def sync():
        torch.cuda.current_stream().synchronize()

    def fn(ctx):
        r = ctx.rank
        if name == "oneshot_reg":
            cc = cc_of(ctx); t = torch.full((1000,), float(r + 1), device="cuda"); cc.register(t)
            cc.allreduce(t, algo="one_shot"); sync(); assert float(t[0]) == size * (size + 1) / 2
        elif name == "twoshot_reg":
            cc = cc_of(ctx); t = torch.full((1 << 20,), float(r + 1), device="cuda"); cc.register(t)
            cc.allreduce(t, algo="two_shot"); sync(); assert float(t[-1]) == size * (size + 1) / 2
        elif name == "bcast_staged_small":
            cc = cc_of(ctx); t = torch.full((4096,), float(r + 1), device="cuda")
            cc.broadcast(t, root=size - 1); sync(); assert float(t[0]) == size
        elif name == "bcast_staged_large":
            cc = cc_of(ctx); t = torch.full((1 << 20,), float(r + 1), device="cuda")
            cc.broadcast(t, root=0); sync(); assert float(t[-1]) == 1.0
        elif name == "bcast_sym":
            cc = cc_of(ctx); t = cc.empty(1 << 18, torch.float32); t.fill_(float(r + 1))
            cc.broadcast(t, root=0); sync(); assert float(t[-1]) == 1.0
        elif name == "allgather_staged":
            cc = cc_of(ctx); out = torch.zeros(size * 128, device="cuda")
            cc.allgather(out, torch.full((128,), float(r), device="cuda")); sync()
            assert float(out[-1]) == size - 1
        elif name == "allgather_sym":
            cc = cc_of(ctx); out = cc.empty(size * 4096, torch.float32)
            cc.allgather(out, torch.full((4096,), float(r), device="cuda")); sync()
            assert float(out[-1]) == size - 1
        elif name == "alltoall_staged":
            cc = cc_of(ctx); inp = torch.cat([torch.full((64,), float(r * 10 + j)) for j in range(size)]).cuda()
            out = torch.zeros(size * 64, device="cuda"); cc.alltoall(out, inp); sync()
            assert float(out[0]) == float(r)
        elif name == "reduce_scatter_staged":
            cc = cc_of(ctx); inp = torch.full((size * 100,), float(r + 1), device="cuda")
            out = torch.zeros(100, device="cuda"); cc.reduce_scatter(out, inp); sync()
            assert float(out[0]) == size * (size + 1) / 2
        elif name == "barrier":
            cc = cc_of(ctx)
            for _ in range(5):
                cc.barrier()
            sync()
        elif name == "smoke_seq":
            cc = cc_of(ctx)
            for count, algo in ((1000, "one_shot"), (1 << 20, "two_shot")):
                t = torch.full((count,), float(r + 1), device="cuda"); cc.register(t)
                cc.allreduce(t, algo=algo); sync(); log(ctx, "allreduce", algo, float(t[0]))
            b = torch.full((4096,), float(r + 1), device="cuda")
            cc.broadcast(b, root=1); sync(); log(ctx, "broadcast", float(b[0]))
            out = torch.empty(size * 128, device="cuda")
            cc.allgather(out, torch.full((128,), float(r), device="cuda")); sync(); log(ctx, "allgather", float(out[128]))
        elif name == "old_loop":
            streams = [gcu.new_stream(0) for _ in range(2)]
            names = [gcu.CudaAllreduceRing, gcu.CudaAllreduceRingChunked, gcu.CudaAllreduceHalvingDoubling]
            for cls in names:
                for count in (100, 200000):
                    ts = [torch.full((count,), float(r * 2 + i + 1), device="cuda") for i in range(2)]
                    log(ctx, cls.__name__, count, "ctor sync")
                    algo = cls(ctx, ts)
                    log(ctx, "run sync")
                    algo.run()
                    log(ctx, "sync done", float(ts[0][0]))
                    ts2 = [torch.empty(count, device="cuda") for _ in range(2)]
                    algo2 = cls(ctx, ts2, streams=streams)
                    for i, s in enumerate(streams):
                        gb._C.cuda.spin(200000, s.cuda_stream)
                        gb._C.cuda.fill(ts2[i].data_ptr(), count, int(gb.DataType.FLOAT32), float(r * 2 + i + 1), 0.0, s.cuda_stream)
                    log(ctx, "run async")
                    algo2.run()
                    for s in streams:
                        s.synchronize()
                    log(ctx, "async done", float(ts2[0][0]))
            gcu._cu.peer_context_for(ctx, 0).host_barrier()
        elif name.startswith("old_"):
            ptrs = 2 if "2ptr" in name else 1
            count = 200000 if "big" in name else 100
            ts = [torch.full((count,), float(r * ptrs + i + 1), device="cuda") for i in range(ptrs)]
            log(ctx, "constructing")
            if "async" in name:
                streams = [gcu.new_stream(0) for _ in range(ptrs)]
                a = gcu.CudaAllreduceRingChunked(ctx, ts, streams=streams)
            else:
                a = gcu.CudaAllreduceRingChunked(ctx, ts)
            log(ctx, "constructed", a.resolved_algo())
            reps = 3 if "rep" in name else 1
            for i in range(reps):
                a.run()
                log(ctx, "run", i, "returned")
            if "async" in name:
                for s in streams:
                    s.synchronize()
            tot = size * ptrs
            exp = tot * (tot + 1) / 2
            got = float(ts[0][0])
            log(ctx, "value", got, "expected", exp)
            assert reps > 1 or got == exp
            gcu._cu.peer_context_for(ctx, 0).host_barrier()
        else:
            raise SystemExit(f"unknown case {name}")
        if not name.startswith("old_"):
            cc.pc.host_barrier()
        log(ctx, "case done")
        return True

    assert all(gb.spawn_threads(size, fn, cuda_device=0))


CASES = ["barrier", "oneshot_reg", "twoshot_reg", "bcast_staged_small", "bcast_staged_large", "bcast_sym",
         "allgather_staged", "allgather_sym", "alltoall_staged", "reduce_scatter_staged",
         "old_1ptr", "old_1ptr_big", "old_2ptr", "old_2ptr_big", "old_1ptr_async", "old_2ptr_async_big", "old_2ptr_rep"]

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        _run_case(sys.argv[2])
        sys.exit(0)
    cases = sys.argv[1:] or CASES
    timeout = int(os.environ.get("DBG_TIMEOUT", "45"))
    for c in cases:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--case", c], capture_output=True, text=True, timeout=timeout)
            status = "PASS" if p.returncode == 0 else "FAIL"
            tail = (p.stdout + p.stderr)[-2500:] if (status == "FAIL" or os.environ.get("DBG_VERBOSE")) else ""
        except subprocess.TimeoutExpired as e:
            status = "HANG"
            tail = ((e.stdout or b"").decode(errors="replace") + (e.stderr or b"").decode(errors="replace"))[-2500:]
        print(f"=== {c}: {status} ({time.time() - t0:.1f}s)", flush=True)
        if tail:
            print(tail, flush=True)
