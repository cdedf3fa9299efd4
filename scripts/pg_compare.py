"""torch.distributed on CPU tensors: backend "gloo" (PyTorch's ProcessGroupGloo over pytorch/gloo)
vs backend "glb" (this library), same script, same machine.

  python scripts/pg_compare.py [--world 4] [--backends gloo,glb]

Spawns WORLD processes per backend; rank 0 prints p50 latency (us) per op and size.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(backend, path, rank, world):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    if backend == "glb":
        import gloo_b200.parallel.process_group  # noqa: F401
    dist.init_process_group(backend, init_method=f"file://{path}", rank=rank, world_size=world)
    res = {}
    for name in ("all_reduce", "broadcast", "all_gather", "all_to_all"):
        for n in (1, 1024, 65536, 1 << 20, 1 << 22):
            x = torch.ones(n)
            if name == "all_gather":
                out = torch.empty(n * world)
                fn = lambda: dist.all_gather_into_tensor(out, x)  # noqa: E731
            elif name == "all_to_all":
                if n < world:
                    continue
                xi = torch.ones(n // world * world)
                out = torch.empty_like(xi)
                fn = lambda: dist.all_to_all_single(out, xi)  # noqa: E731
            elif name == "broadcast":
                fn = lambda: dist.broadcast(x, src=0)  # noqa: E731
            else:
                fn = lambda: dist.all_reduce(x)  # noqa: E731
            iters = 200 if n <= 65536 else (30 if n <= (1 << 20) else 12)
            for _ in range(5):
                fn()
            dist.barrier()
            ts = []
            for _ in range(iters):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            ts.sort()
            res[f"{name}/{n}"] = ts[len(ts) // 2] * 1e6
    dist.barrier()
    if rank == 0:
        print("RESULT " + json.dumps(res), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=4)
    ap.add_argument("--backends", default="gloo,glb")
    ap.add_argument("--worker", nargs=4)
    a = ap.parse_args()
    if a.worker:
        worker(a.worker[0], a.worker[1], int(a.worker[2]), int(a.worker[3]))
        return
    table = {}
    for be in a.backends.split(","):
        path = os.path.join(tempfile.mkdtemp(prefix="pgcmp_"), "init")
        procs = [subprocess.Popen([sys.executable, __file__, "--worker", be, path, str(r), str(a.world)],
                                  stdout=subprocess.PIPE, text=True) for r in range(a.world)]
        outs = [p.communicate(timeout=900)[0] for p in procs]
        line = [l for l in outs[0].splitlines() if l.startswith("RESULT ")]
        table[be] = json.loads(line[0][7:]) if line else {}
    keys = list(next(iter(table.values())).keys())
    print(f"world={a.world}  p50 latency in us (float32 elements)")
    print(f"{'op/elements':28s}" + "".join(f"{b:>12s}" for b in table))
    for k in keys:
        print(f"{k:28s}" + "".join(f"{table[b].get(k, float('nan')):12.1f}" for b in table))


if __name__ == "__main__":
    main()
