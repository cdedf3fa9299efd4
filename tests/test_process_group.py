"""torch.distributed backend "glb": every c10d collective on CPU tensors, point-to-point,
sub-groups and DistributedDataParallel, with real processes (the consumer-facing counterpart
of ProcessGroupGloo, which is how most users reach pytorch/gloo)."""
import os
import subprocess
import sys
import tempfile

import pytest

WORKER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pg_worker.py")


@pytest.mark.parametrize("size", [1, 2, 3])
def test_torch_distributed_backend(size):
    init = os.path.join(tempfile.mkdtemp(prefix="glb_pg_"), "init")
    procs = [subprocess.Popen([sys.executable, WORKER, init, str(r), str(size)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(size)]
    outs = [p.communicate(timeout=240) for p in procs]
    for r, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in out, (r, p.returncode, out[-500:], err[-2000:])


@pytest.mark.gpu
def test_torch_distributed_backend_cuda():
    """CUDA tensors through the same backend: one process per GPU when there are several,
    otherwise a single rank (plumbing only: a one-rank collective is the identity)."""
    import torch

    size = min(torch.cuda.device_count(), 2)  # the configuration measured on hardware
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pg_cuda_worker.py")
    init = os.path.join(tempfile.mkdtemp(prefix="glb_pg_cuda_"), "init")
    procs = [subprocess.Popen([sys.executable, worker, init, str(r), str(size)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(size)]
    outs = [p.communicate(timeout=240) for p in procs]
    for r, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in out, (r, p.returncode, out[-500:], err[-2000:])


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="written after the round's GPU minutes were spent; not yet run on hardware")
def test_gradient_bucketer_overlapped_multiprocess():
    """GradientBucketer with one process per GPU: buckets are launched from the autograd hooks while backward
    runs (kept at the very end of the GPU tests: it is the newest consumer)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    size = 2
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bucketer_worker.py")
    d = tempfile.mkdtemp(prefix="glb_bucketer_")
    procs = [subprocess.Popen([sys.executable, worker, d, str(r), str(size)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(size)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"WORKER {r} OK" in out, (r, p.returncode, out[-3000:])
