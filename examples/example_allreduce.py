"""Host allreduce over the TCP transport (counterpart of gloo/examples/example_allreduce.cc).

  python examples/example_allreduce.py <rank> <size> <rendezvous dir>
"""
import sys

import numpy as np

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import gloo_b200 as gb  # noqa: E402

rank, size, path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
ctx = gb.init_context(rank, size, path=path)            # FileStore rendezvous + full TCP mesh
data = np.full(4, rank, np.int32)
out = np.zeros(4, np.int32)
gb.allreduce(ctx, out, inputs=data)                      # new-style, out of place
print(f"rank {rank}: {out}")
algo = gb.ops.algorithms.AllreduceRingChunked(ctx, data)  # old-style, in place, reusable
algo.run()
print(f"rank {rank}: {data}")
gb.barrier(ctx)
ctx.close_connections()
