#!/usr/bin/env python
"""Headline benchmark: cuda_allreduce_ring_chunked, float32, on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 20 --warmup 5
    ... bench.py --impl reference ...   # unmodified pytorch/gloo benchmark_cuda from baseline/_ref

A *step* is one `CudaAllreduceRingChunked.run()` on a 1e8-element float32 buffer per
GPU (400 MB, larger than the 126 MB L2, so back-to-back steps cannot hit in cache).
Timing: CUDA events on the launching stream around exactly K steps, barrier +
synchronize on both sides, max over ranks. `value` is the allreduce bus bandwidth
(algbw * 2(N-1)/N, GB/s per GPU — the figure BASELINE.json asks for, against the
900 GB/s/direction NVLink roofline). With N=1 a one-buffer allreduce is the
identity, so the N=1 run uses two local buffers per rank (`--inputs 2`, which the
reference's benchmark supports too) and reports algorithm bandwidth instead.
The JSON also carries a latency sweep (p50/p99, L2 flushed between iterations),
an end-to-end number (pinned-host H2D of the input and D2H of the result every
step) and the clocks seen while timing.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# `value` is the whole-job aggregate the driver contract asks for: the per-GPU bus bandwidth
# (algbw x 2(N-1)/N, the number NCCL-style tables quote and `busbw_gbs` keeps) summed over the
# N GPUs. At N=1 there is no bus: the step is the local reduce + re-broadcast of the two
# buffers and the value is its algorithm bandwidth.
METRIC = ("cuda_allreduce_ring_chunked float32 aggregate bus bandwidth (GB/s summed over the N GPUs = "
          "N x algbw x 2(N-1)/N; algbw at N=1)")
PUBLISHED_BUSBW_GBS = 2.8  # 20 MB allreduce_ring_chunked, 4 machines, 40 GbE (derived busbw)
HEADLINE_ELEMENTS = 100_000_000
SWEEP = [1, 10, 100, 1_000, 10_000, 100_000, 1_000_000, 10_000_000]


def env_int(name, dflt):
    return int(os.environ.get(name, dflt))


def rendezvous_dir(tag: str) -> str:
    port = os.environ.get("MASTER_PORT", "0")
    launcher = os.getppid() if "RANK" in os.environ else os.getpid()
    return f"/tmp/glb_bench_{tag}_{port}_{launcher}"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# our implementation
# ------------------------------------------------------------------------------------------

def run_ours(args):
    import numpy as np
    import torch

    import gloo_b200 as gb
    from gloo_b200.ops import cuda as gcu

    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # Keep this rank (its pinned staging buffers and the transport's I/O thread) on the
    # socket its GPU hangs off.
    from gloo_b200.utils.affinity import bind_to_gpu
    cpus = bind_to_gpu(local)

    store = gb.FileStore(rendezvous_dir("ours"))
    ctx = gb.init_context(rank, world, store=store, device=gb.create_device("127.0.0.1"), timeout_ms=120000)
    inputs = 2 if world == 1 else 1
    cc = gcu.CudaContext(ctx, local, stage_bytes=64 << 20)
    stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2

    def host_max(values):
        arr = np.asarray(values, dtype=np.float64).copy()
        if world > 1:
            gb.allreduce(ctx, arr, op=gb.ReduceOp.MAX)
        return arr

    def sync_all():
        torch.cuda.synchronize()
        gb.barrier(ctx)

    def make(elements, symmetric=True):
        # The harness owns the buffers (as the reference's benchmark does). By default they
        # come from the library's symmetric allocator: peer-mapped and, on >2 GPUs, bound to
        # an NVSwitch multicast object so the reduction can run inside the switch. Plain
        # cudaMalloc'ed tensors (symmetric=False) are registered through cudaIpc instead.
        if symmetric:
            ts = [cc.empty(elements, torch.float32) for _ in range(inputs)]
        else:
            ts = [torch.empty(elements, dtype=torch.float32, device=dev) for _ in range(inputs)]
        algo = gcu.CudaAllreduceRingChunked(ctx, ts, streams=[stream] * inputs)
        return ts, algo

    def fill(ts, elements):
        total = world * inputs
        for i, t in enumerate(ts):
            gb._C.cuda.fill(t.data_ptr(), elements, int(gb.DataType.FLOAT32), float(rank * inputs + i), float(total),
                            stream.cuda_stream)

    def verify(ts, elements):
        total = world * inputs
        n = min(elements, 1 << 16)
        idx = torch.arange(n, dtype=torch.float64)
        exp = idx * total * total + total * (total - 1) / 2
        for t in ts:
            got = t[:n].double().cpu()
            if not torch.allclose(got, exp, rtol=1e-5):
                raise AssertionError(f"allreduce verification failed at {elements} elements")

    # ---- headline: K back-to-back steps, device timed ------------------------------------
    E = args.elements
    ts, algo = make(E)
    with torch.cuda.stream(stream):
        fill(ts, E)
        algo.run()
    stream.synchronize()
    verify(ts, E)
    for _ in range(args.warmup):
        algo.run()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = gb._C.cuda.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record(stream)
    for _ in range(args.steps):
        algo.run()
    ev1.record(stream)
    stream.synchronize()
    sync_all()
    launches = gb._C.cuda.launch_count() - launches0
    total_ms = float(host_max([ev0.elapsed_time(ev1)])[0])
    ms_per_step = total_ms / args.steps
    size_bytes = E * 4
    algbw = size_bytes / (ms_per_step * 1e-3) / 1e9
    busbw = algbw * 2 * (world - 1) / world if world > 1 else None
    resolved = algo.resolved_algo()

    # ---- end to end: pinned host -> device, allreduce, device -> pinned host, every step ----
    # The call a user with host-resident data makes: gcu.CudaHostAllreduce, which pipelines
    # the PCIe legs with the NVLink reduction piece by piece (same kernels as above).
    del ts, algo
    hins = [torch.empty(E, dtype=torch.float32).pin_memory() for _ in range(inputs)]
    for h in hins:
        h.fill_(1.0)
    hout = torch.empty(E, dtype=torch.float32).pin_memory()
    e2e = gcu.CudaHostAllreduce(ctx, cc, hins, hout, chunks=args.e2e_chunks)
    def e2e_step():
        with torch.cuda.stream(stream):
            e2e.run()
    for _ in range(min(3, args.warmup)):
        e2e_step()
    sync_all()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record(stream)
    for _ in range(args.steps):
        e2e_step()
    eb.record(stream)
    stream.synchronize()
    sync_all()
    e2e_ms = float(host_max([ea.elapsed_time(eb)])[0]) / args.steps
    # The sampler ran through both timed regions (device-only steps and end-to-end steps).
    clocks = sampler.stop() if rank == 0 else None
    assert abs(float(hout[0]) - world * inputs) < 1e-3, "e2e result mismatch"
    e2e_algbw = size_bytes / (e2e_ms * 1e-3) / 1e9
    e2e_val = e2e_algbw * 2 * (world - 1) / world if world > 1 else e2e_algbw
    e2e_launches = e2e.launches_per_run
    del hins, hout, e2e

    # ---- same size on plain cudaMalloc'ed buffers (cudaIpc-registered) and the NCCL comparator ----
    def time_k(fn):
        for _ in range(3):
            fn()
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            fn()
        b.record(stream)
        stream.synchronize()
        sync_all()
        return float(host_max([a.elapsed_time(b)])[0]) / args.steps

    extra = {}
    if world > 1:
        pts, palgo = make(E, symmetric=False)
        with torch.cuda.stream(stream):
            fill(pts, E)
        pms = time_k(palgo.run)
        extra["plain_cudamalloc_buffers"] = {"ms_per_step": round(pms, 5), "variant": palgo.resolved_algo(),
                                             "busbw_gbs": round(size_bytes / (pms * 1e-3) / 1e9 * 2 * (world - 1) / world, 3)}
        try:
            nccl = gb._C.cuda.NcclComm.init_rank(ctx, local)
            ptr = pts[0].data_ptr()
            nms = time_k(lambda: nccl.allreduce(ptr, ptr, E, int(gb.DataType.FLOAT32), 1, stream.cuda_stream))
            extra["nccl_comparator"] = {"ms_per_step": round(nms, 5), "version": gb._C.cuda.nccl_version(),
                                        "busbw_gbs": round(size_bytes / (nms * 1e-3) / 1e9 * 2 * (world - 1) / world, 3)}
            del nccl
        except Exception as e:  # noqa: BLE001
            extra["nccl_comparator"] = {"unavailable": str(e)[:200]}
        del pts, palgo

    # ---- latency sweep: per-iteration events, L2 flushed between iterations ----------------
    sweep = []
    if not args.no_sweep:
        iters = max(args.steps, 20)
        for n in SWEEP:
            ts, algo = make(n)
            with torch.cuda.stream(stream):
                fill(ts, n)
                algo.run()
            stream.synchronize()
            verify(ts, n)
            for _ in range(max(3, args.warmup)):
                algo.run()
            sync_all()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
            with torch.cuda.stream(stream):
                for a, b in evs:
                    flush.fill_(0)
                    a.record(stream)
                    algo.run()
                    b.record(stream)
            stream.synchronize()
            sync_all()
            per = host_max([a.elapsed_time(b) * 1e3 for a, b in evs])  # us, max over ranks per iteration
            per.sort()
            p50, p99 = float(per[len(per) // 2]), float(per[min(len(per) - 1, int(len(per) * 0.99))])
            ab = n * 4 / (p50 * 1e-6) / 1e9
            sweep.append({"elements": n, "bytes": n * 4, "p50_us": round(p50, 2), "p99_us": round(p99, 2),
                          "min_us": round(float(per[0]), 2), "algbw_gbs": round(ab, 3),
                          "busbw_gbs": round(ab * 2 * (world - 1) / world, 3) if world > 1 else None,
                          "variant": algo.resolved_algo()})
            del ts, algo
    sync_all()
    if rank == 0:
        per_gpu = busbw if world > 1 else algbw
        value = per_gpu * world
        out = {
            "metric": METRIC,
            "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": round(per_gpu / PUBLISHED_BUSBW_GBS, 2),
            "vs_baseline_basis": "per-GPU bus bandwidth / 2.8 GB/s (BASELINE.md: derived busbw of the published "
                                 "20 MB allreduce_ring_chunked row, 4 machines over 40 GbE)",
            "per_gpu_gbs": round(per_gpu, 3), "dtype": "fp32", "data": "synthetic",
            "impl": "gloo_b200",
            "config": {"model": "cuda_allreduce_ring_chunked", "elements": E, "bytes_per_gpu": size_bytes,
                       "inputs_per_rank": inputs, "global_batch": world * inputs, "seq_len": E,
                       "parallelism": f"allreduce x{world}", "cpu_affinity_cpus": len(cpus), "kernel_variant": resolved,
                       "buffers": "library symmetric allocator (peer-mapped; NVLS multicast-bound when >2 GPUs)",
                       "l2": "inputs larger than L2 (400 MB > 126 MB) for the headline; 256 MB flush between sweep iterations",
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "algbw_gbs": round(algbw, 3), "busbw_gbs": round(busbw, 3) if busbw else None,
            "roofline": {"nvlink_gbs_per_dir_nominal": 900, "nvlink_gbs_per_dir_measured": 770,
                         "frac_of_measured": round(busbw / 770, 3) if busbw else None},
            "clocks": clocks,
            "e2e": {"value": round(e2e_val * world, 3), "per_gpu_gbs": round(e2e_val, 3), "unit": "GB/s",
                    "ms_per_step": round(e2e_ms, 4),
                    "h2d_bytes_per_step": size_bytes * inputs, "d2h_bytes_per_step": size_bytes,
                    "api": "gloo_b200.ops.cuda.CudaHostAllreduce (chunked H2D | allreduce | D2H pipeline)",
                    "pieces": int(e2e_launches), "gpu_launches_per_step": int(e2e_launches)},
            "gpu_launches": int(launches),
            **extra,
            "sweep": sweep,
        }
        print(json.dumps(out), flush=True)
    ctx.close_connections()
    shutil.rmtree(rendezvous_dir("ours"), ignore_errors=True)


# ------------------------------------------------------------------------------------------
# reference arm: unmodified pytorch/gloo benchmark_cuda
# ------------------------------------------------------------------------------------------

ROW = re.compile(r"^\s*(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+([0-9.]+)\s+(\d+)\s*$")


def run_reference(args):
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    local = env_int("LOCAL_RANK", 0)
    binary = os.path.join(ROOT, "baseline", "_ref", "bin", "benchmark_cuda")
    if not os.path.exists(binary):
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/bin/benchmark_cuda missing "
                              "(pip cannot install the CMake-only reference; build it with baseline/build_reference.sh)"}))
        return
    # The reference cannot place two inputs on one GPU (its local reduce goes through NCCL,
    # which needs unique devices: nccl.cu:111-116), so its 1-GPU run uses one buffer and
    # does D2H + H2D of it; our 1-GPU run reduces and re-broadcasts two buffers (more work).
    inputs = 1
    base = rendezvous_dir("ref")
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = str(local)

    def one(elements, iters, warmup):
        d = f"{base}_{elements}"
        os.makedirs(d, exist_ok=True)
        cmd = [binary, "--size", str(world), "--rank", str(rank), "--shared-path", d, "--transport", "tcp",
               "--tcp-device", "lo", "--elements", str(elements), "--iteration-count", str(iters),
               "--warmup-iters", str(warmup), "--inputs", str(inputs), "--no-verify", "--nanos",
               "cuda_allreduce_ring_chunked"]
        p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1500)
        row = None
        for ln in p.stdout.splitlines():
            m = ROW.match(ln)
            if m:
                row = [float(x) for x in m.groups()]
        if p.returncode != 0 or (rank == 0 and row is None):
            raise RuntimeError(f"reference benchmark failed (rc={p.returncode}): {p.stderr[-400:]} {p.stdout[-400:]}")
        return row

    try:
        head = one(args.elements, args.steps, args.warmup)
        sweep = []
        if not args.no_sweep:
            for n in (1_000, 100_000, 1_000_000, 10_000_000):
                r = one(n, max(args.steps, 20), max(args.warmup, 3))
                if rank == 0:
                    p50 = r[3] / 1e3
                    ab = n * 4 / (p50 * 1e-6) / 1e9
                    sweep.append({"elements": n, "bytes": n * 4, "p50_us": round(p50, 2), "p99_us": round(r[4] / 1e3, 2),
                                  "min_us": round(r[2] / 1e3, 2), "algbw_gbs": round(ab, 4),
                                  "busbw_gbs": round(ab * 2 * (world - 1) / world, 4) if world > 1 else None})
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": f"reference run failed: {str(e)[:300]}"}))
        return
    if rank == 0:
        size_bytes, _, mn, p50, p99, mx, bw_gib, iters = head
        # Mean latency from the reference's own bandwidth column (GiB/s over the sum of
        # latencies, runner.cc:499-508); fall back to p50 when the column underflows.
        mean_ns = size_bytes / (bw_gib * (1 << 30)) * 1e9 if bw_gib > 0 else p50
        ms = mean_ns / 1e6
        algbw = size_bytes / (ms * 1e-3) / 1e9
        per_gpu = algbw * 2 * (world - 1) / world if world > 1 else algbw
        value = per_gpu * world
        print(json.dumps({
            "metric": METRIC,
            "value": round(value, 4), "unit": "GB/s", "n_gpus": world, "steps": int(iters), "warmup": args.warmup,
            "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": round(per_gpu / PUBLISHED_BUSBW_GBS, 3), "per_gpu_gbs": round(per_gpu, 4),
            "dtype": "fp32", "data": "synthetic",
            "impl": "reference",
            "config": {"model": "cuda_allreduce_ring_chunked", "elements": args.elements, "bytes_per_gpu": int(size_bytes),
                       "inputs_per_rank": inputs, "global_batch": world * inputs, "seq_len": args.elements,
                       "parallelism": f"allreduce x{world}",
                       "timing": "reference's own harness: host wall clock on rank 0 (runner.cc:641-645), "
                                 "mean from its bandwidth column",
                       "path": "pytorch/gloo benchmark_cuda (CudaHostWorkspace: GPU->pinned host->TCP loopback->CPU reduce)"},
            "p50_us": round(p50 / 1e3, 2), "p99_us": round(p99 / 1e3, 2), "sweep": sweep,
        }), flush=True)
    for d in glob.glob(base + "_*"):
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elements", type=int, default=HEADLINE_ELEMENTS)
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=16, help="pieces of the host<->device pipeline in the e2e run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
