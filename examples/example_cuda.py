"""CUDA collectives over NVLink peer memory. One process per GPU:

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/example_cuda.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import gloo_b200 as gb  # noqa: E402
from gloo_b200.ops import cuda as gcu  # noqa: E402

rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
ctx = gb.init_context(rank, size, path=f"/tmp/glb_example_{os.environ['MASTER_PORT']}_{os.getppid()}")
cc = gcu.CudaContext(ctx, local)
if rank == 0:
    print(cc.describe())
    for d in cc.topology():
        print(" ", d.hostname, d.device, d.pci_bus_id, "multicast" if d.multicast_supported else "")

grads = cc.empty(1 << 24, torch.bfloat16)        # symmetric: zero-copy, NVLS-capable
grads.fill_(rank)
cc.allreduce(grads)                               # one fused kernel on the current stream
shard = torch.empty(grads.numel() // size, dtype=grads.dtype, device="cuda")
cc.reduce_scatter(shard, grads)                   # ZeRO-style
cc.allgather(grads, shard)
tokens = torch.randn(size * 1024, 64, device="cuda")
routed = torch.empty_like(tokens)
cc.alltoall(routed, tokens)                       # MoE dispatch / Ulysses
torch.cuda.synchronize()
print(f"rank {rank}: allreduce -> {float(grads[0]):.0f}")
gb.barrier(ctx)
ctx.close_connections()
