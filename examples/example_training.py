"""A data-parallel training loop on top of the library: gradients live in symmetric buckets and are
averaged by the fused NVLink kernels while backward is still running (GradientBucketer); on a machine
without GPUs the same script runs on CPU tensors over the TCP transport.

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/example_training.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

import gloo_b200 as gb  # noqa: E402
from gloo_b200.parallel import DataParallel, GradientBucketer  # noqa: E402

rank, size = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
steps = int(os.environ.get("STEPS", 20))
ctx = gb.init_context(rank, size, path=f"/tmp/glb_train_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}")
cuda = torch.cuda.is_available()
cc = None
if cuda:
    from gloo_b200.ops import cuda as gcu

    torch.cuda.set_device(local)
    cc = gcu.CudaContext(ctx, local)
dev = torch.device("cuda", local) if cuda else torch.device("cpu")

torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(256, 1024), torch.nn.GELU(), torch.nn.Linear(1024, 1024), torch.nn.GELU(),
                            torch.nn.Linear(1024, 16)).to(dev)
DataParallel(ctx, cc).broadcast_parameters(model.parameters())           # same start everywhere
sync = GradientBucketer(ctx, cc, model.parameters(), bucket_bytes=1 << 20)  # p.grad = views into symmetric buckets
opt = torch.optim.SGD(model.parameters(), lr=0.05)

g = torch.Generator().manual_seed(1234)
w_true = torch.randn(256, 16, generator=g)
first = last = None
for step in range(steps):
    x = torch.randn(64, 256, generator=g)
    y = x @ w_true
    xs, ys = x[rank::size].to(dev), y[rank::size].to(dev)                   # this rank's shard of the batch
    loss = torch.nn.functional.mse_loss(model(xs), ys)
    loss.backward()                                                         # buckets are reduced as they fill
    sync.finish()                                                           # compute stream waits for the results
    opt.step()
    sync.zero_grad()                                                        # in place: the views stay
    last = float(loss.detach())
    first = last if first is None else first
if rank == 0:
    print(f"loss {first:.4f} -> {last:.4f} over {steps} steps on {size} rank(s), {'cuda' if cuda else 'cpu'}")
gb.barrier(ctx)
ctx.close_connections()
