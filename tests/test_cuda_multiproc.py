"""Process-per-GPU runs of the CUDA collectives (needs >= 2 GPUs; the single-GPU CI box
covers the same kernels with threads as ranks)."""
import os
import subprocess
import sys
import tempfile

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
WORKER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_worker.py")


def _run(size, env_extra=None):
    d = tempfile.mkdtemp(prefix="glb_cuda_mp_")
    env = dict(os.environ, **(env_extra or {}))
    procs = [subprocess.Popen([sys.executable, WORKER, d, str(r), str(size)], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(size)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
        outs.append(o)
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"WORKER {r} OK" in o, f"rank {r} failed:\n{o[-3000:]}"
    return outs


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("mode", ["vmm", "ipc"])
def test_all_gpus(mode):
    n = torch.cuda.device_count()
    size = 8 if n >= 8 else (4 if n >= 4 else 2)
    outs = _run(size, {"GLB_CUDA_VMM": "1" if mode == "vmm" else "0"})
    print(outs[0][:400])


BENCH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gloo_b200", "bin", "glb_benchmark")


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2 or not os.path.exists(BENCH),
                    reason="needs >= 2 GPUs and the benchmark binary")
@pytest.mark.parametrize("name", ["cuda_allreduce_ring_chunked", "cuda_allreduce_halving_doubling_pipelined", "cuda_broadcast_one_to_all",
                                  "cuda_allgather", "cuda_alltoall", "cuda_reduce_scatter", "cuda_sendrecv", "cuda_exchange"])
def test_benchmark_cli_cuda(name):
    """The CLI's CUDA benchmarks, one process per GPU (the reference's benchmark_cuda invocation)."""
    d = tempfile.mkdtemp(prefix="glb_cli_cuda_")
    size = 2
    procs = [subprocess.Popen([BENCH, "--size", str(size), "--rank", str(r), "--shared-path", d, "--elements", "1000000",
                               "--iteration-count", "10", name],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(size)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert [p.returncode for p in procs] == [0] * size, outs
    assert any(l.split() and l.split()[0] == "4000000" for l in outs[0].splitlines()), outs[0]
