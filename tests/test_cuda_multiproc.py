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
