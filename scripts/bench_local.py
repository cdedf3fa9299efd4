#!/usr/bin/env python
"""Launch-shape sweep of the single-rank fused step (localAllreduceManyKernel: every buffer :=
sum of all, one pass) on the benchmark's buffers (2 x 400 MB fp32): CTAs per SM x packs per
thread x tiled / grid-stride. Device timed, p50 of 20; HBM bytes = 2 reads + 2 writes of 400 MB.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import gloo_b200 as gb  # noqa: E402

cu = gb._C.cuda
n = 100_000_000
a, b = torch.ones(n, device="cuda"), torch.ones(n, device="cuda")
F32 = int(gb.DataType.FLOAT32)
stream = torch.cuda.current_stream().cuda_stream
rows = []
for tiled in (False, True):
    for ctas in (2, 4, 6, 8):
        for unroll in (1, 2, 4):
            cu.set_local_shape(ctas, unroll, tiled)
            for _ in range(3):
                cu.local_allreduce_many([a.data_ptr(), b.data_ptr()], n, F32, 1, 0.5, stream)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
            for x, y in evs:
                x.record()
                cu.local_allreduce_many([a.data_ptr(), b.data_ptr()], n, F32, 1, 0.5, stream)
                y.record()
            torch.cuda.synchronize()
            per = sorted(x.elapsed_time(y) for x, y in evs)
            ms = per[len(per) // 2]
            ok = float(a[0]) == 1.0 and float(b[-1]) == 1.0 and float(a[n // 3]) == 1.0
            rows.append({"tiled": tiled, "ctas_per_sm": ctas, "unroll": unroll, "ms": round(ms, 4),
                         "hbm_gbs": round(4 * n * 4 / (ms * 1e-3) / 1e9, 1), "ok": ok})
            print(json.dumps(rows[-1]), flush=True)
best = min(rows, key=lambda r: r["ms"])
print("best:", json.dumps(best))
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump({"rows": rows, "best": best}, f, indent=1)
