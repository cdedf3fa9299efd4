"""Device-side failure detection across processes: one of the ranks is SIGKILLed / SIGSTOPped
right before a collective; the survivors' kernels must give up within the device timeout,
the host must see IoError within 2x the timeout and the GPU must stay usable.
(reference behaviour on its transports: gloo/test/transport_test.cc:53-164)."""
import os
import signal
import subprocess
import sys
import tempfile

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
WORKER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_fault_worker.py")


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("mode", ["kill", "stop"])
@pytest.mark.parametrize("algo", ["two_shot", "ll", "pipelined"])
def test_survivors_raise_when_a_peer_disappears(mode, algo):
    size = min(4, torch.cuda.device_count())
    d = tempfile.mkdtemp(prefix="glb_cuda_fault_")
    procs = [subprocess.Popen([sys.executable, WORKER, d, str(r), str(size), mode, algo], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(size)]
    outs = []
    try:
        for r, p in enumerate(procs[:-1]):
            o, _ = p.communicate(timeout=120)
            outs.append(o)
            assert p.returncode == 0 and f"SURVIVOR {r} OK" in o, f"rank {r}:\n{o[-3000:]}"
    finally:
        for p in procs:
            if p.poll() is None:
                p.send_signal(signal.SIGCONT)
                p.kill()
        procs[-1].communicate()
    print(outs[0][-300:])
