"""Fault injection with processes as ranks — mirrors gloo/test/transport_test.cc:53-386
and multiproc_test.{h,cc}: SIGKILL a rank mid-collective -> survivors raise IoError
quickly; SIGSTOP a rank -> survivors hit the timeout path."""
import os
import signal
import subprocess
import sys
import tempfile
import time

import pytest

WORKER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multiproc_worker.py")


def launch(size, mode, *args, timeout_ms=3000, extra_env=None):
    d = tempfile.mkdtemp(prefix="glb_mp_")
    env = dict(os.environ, GLB_TEST_TIMEOUT_MS=str(timeout_ms), **(extra_env or {}))
    procs = [subprocess.Popen([sys.executable, WORKER, d, str(r), str(size), mode, *map(str, args)], env=env,
                              stderr=subprocess.PIPE, text=True) for r in range(size)]
    return d, procs


def wait_ready(d, size, timeout=60):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if all(os.path.exists(os.path.join(d, f"ready_{r}")) for r in range(size)):
            return
        time.sleep(0.02)
    raise AssertionError("workers did not come up")


def reap(procs, timeout):
    out = []
    for p in procs:
        try:
            p.wait(timeout=timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            p.wait()
        out.append(p.returncode)
    return out


@pytest.mark.parametrize("size", [2, 4])
def test_healthy_run(size):
    d, procs = launch(size, "allreduce_once")
    assert reap(procs, 60) == [0] * size


@pytest.mark.parametrize("cma", ["1", "0"])
@pytest.mark.parametrize("size", [2, 3])
def test_large_messages_between_processes(size, cma):
    """Same-host single-copy path (process_vm_readv) on and off; the worker asserts the
    path it expected was the one that carried the payload."""
    d, procs = launch(size, "large_once", 1 << 20, timeout_ms=20000, extra_env={"GLB_TCP_CMA": cma})
    codes = reap(procs, 90)
    assert codes == [0] * size, (codes, [p.stderr.read()[-400:] for p in procs])


@pytest.mark.parametrize("size", [2, 3, 4])
@pytest.mark.parametrize("mode", ["allreduce_loop", "sendrecv_loop", "allreduce_loop_large"])
def test_sigkill_is_detected(size, mode):
    timeout_ms = 3000
    args = ()
    if mode == "allreduce_loop_large":  # dies while peers pull from / wait for FIN of its memory
        mode, args = "allreduce_loop", (1 << 19,)
    d, procs = launch(size, mode, *args, timeout_ms=timeout_ms)
    wait_ready(d, size)
    time.sleep(0.3)
    t0 = time.time()
    procs[0].send_signal(signal.SIGKILL)
    codes = reap(procs, 2 * timeout_ms / 1000 + 10)
    elapsed = time.time() - t0
    assert codes[0] == -signal.SIGKILL
    assert all(c == 10 for c in codes[1:]), (codes, [p.stderr.read()[-300:] for p in procs[1:]])
    assert elapsed < 2 * timeout_ms / 1000 + 5


@pytest.mark.parametrize("sync", ["1", "2"])
@pytest.mark.parametrize("mode", ["allreduce_loop", "sendrecv_loop"])
def test_sigkill_is_detected_in_sync_modes(mode, sync):
    """Blocking (poll) and busy-polling pairs: the waiting user thread reads the socket itself
    and must notice the dead peer just like the loop thread does in async mode."""
    size, timeout_ms = 3, 3000
    d, procs = launch(size, mode, timeout_ms=timeout_ms, extra_env={"GLB_TEST_SYNC": sync})
    wait_ready(d, size)
    time.sleep(0.3)
    t0 = time.time()
    procs[0].send_signal(signal.SIGKILL)
    codes = reap(procs, 2 * timeout_ms / 1000 + 10)
    assert codes[0] == -signal.SIGKILL
    assert all(c == 10 for c in codes[1:]), (codes, [p.stderr.read()[-300:] for p in procs[1:]])
    assert time.time() - t0 < 2 * timeout_ms / 1000 + 5


@pytest.mark.parametrize("sync", ["1", "2"])
def test_healthy_run_in_sync_modes(sync):
    d, procs = launch(3, "large_once", 1 << 18, timeout_ms=20000, extra_env={"GLB_TEST_SYNC": sync})
    codes = reap(procs, 90)
    assert codes == [0] * 3, (codes, [p.stderr.read()[-400:] for p in procs])


def test_sender_dies_while_its_message_is_parked():
    """Single-copy path: a large message whose recv is not posted yet exists only as a
    descriptor of the sender's memory. If the sender dies first, posting the recv must raise
    IoError (the pull fails, or the pair has already noticed the EOF) - never hang or crash."""
    d, procs = launch(2, "parked", timeout_ms=3000)
    wait_ready(d, 2)
    time.sleep(0.5)
    procs[0].send_signal(signal.SIGKILL)
    procs[0].wait()
    time.sleep(0.2)
    open(os.path.join(d, "go"), "w").close()
    codes = reap(procs, 20)
    assert codes[0] == -signal.SIGKILL
    assert codes[1] == 10, (codes, procs[1].stderr.read()[-500:])


@pytest.mark.parametrize("size", [2, 3])
def test_sigstop_hits_timeout(size):
    timeout_ms = 1500
    d, procs = launch(size, "allreduce_loop", timeout_ms=timeout_ms)
    wait_ready(d, size)
    time.sleep(0.3)
    procs[0].send_signal(signal.SIGSTOP)
    t0 = time.time()
    codes = reap(procs[1:], 4 * timeout_ms / 1000 + 10)
    elapsed = time.time() - t0
    procs[0].send_signal(signal.SIGKILL)
    procs[0].wait()
    assert all(c == 10 for c in codes), codes
    assert elapsed >= timeout_ms / 1000 * 0.5
    errs = " ".join(p.stderr.read() for p in procs[1:])
    assert "Timed out" in errs or "timeout" in errs.lower() or "closed" in errs.lower()


@pytest.mark.parametrize("seed,size,sync", [(1, 3, "0"), (2, 4, "0"), (3, 3, "1"), (4, 2, "2")])
def test_kill_during_random_collectives(seed, size, sync):
    """A rank dies at an arbitrary moment while all ranks run a random mix of collectives,
    small and large (eager, single-copy, parked messages, acknowledgements in flight): every
    survivor must come back with IoError promptly - no hang, no crash."""
    import random

    rnd = random.Random(seed)
    d, procs = launch(size, "stress_loop", seed, timeout_ms=3000, extra_env={"GLB_TEST_SYNC": sync})
    wait_ready(d, size)
    time.sleep(rnd.uniform(0.05, 0.5))
    victim = rnd.randrange(size)
    t0 = time.time()
    procs[victim].send_signal(signal.SIGKILL)
    codes = reap(procs, 20)
    assert codes[victim] == -signal.SIGKILL
    survivors = [c for i, c in enumerate(codes) if i != victim]
    assert all(c == 10 for c in survivors), (codes, [p.stderr.read()[-300:] for i, p in enumerate(procs) if i != victim])
    assert time.time() - t0 < 12
