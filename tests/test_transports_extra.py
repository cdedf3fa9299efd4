"""Transport names a reference user may ask for besides tcp: uv (the socket transport on the portable poll(2) reactor),
ibverbs (probe + precise error), MPI bootstrap (run against a thread-world MPI), and the
benchmark CLI's --transport switch (reference: gloo/benchmark/options.cc:149-180)."""
import os
import sys
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

import gloo_b200 as gb
from gloo_b200 import _C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "gloo_b200", "bin", "glb_benchmark")


def test_uv_device_runs_new_style_collectives():
    store = gb.HashStore()
    import threading

    out = [None] * 3

    def rank(r):
        dev = _C.create_uv_device("127.0.0.1")
        assert str(dev).startswith("uv(")
        ctx = gb.Context(r, 3)
        ctx.connect_full_mesh(store, dev)
        x = np.full(100, r + 1, np.float64)
        gb.allreduce(ctx, x)
        got = np.zeros(3, np.int32)
        gb.allgather(ctx, got, np.array([r], np.int32))
        gb.barrier(ctx)
        out[r] = (float(x[0]), got.tolist())

    ths = [threading.Thread(target=rank, args=(r,)) for r in range(3)]
    [t.start() for t in ths]
    [t.join(60) for t in ths]
    assert out == [(6.0, [0, 1, 2])] * 3


def test_uv_poll_reactor_interoperates_with_tcp_and_survives_churn():
    """Ranks on the poll(2) reactor and ranks on epoll in one job (same wire protocol); large payloads,
    point to point, then contexts torn down and rebuilt several times on the same devices (descriptor
    removal while the loop thread is polling a snapshot)."""
    import threading

    size = 4
    devs = [_C.create_uv_device("127.0.0.1") if r % 2 == 0 else gb.create_device("127.0.0.1") for r in range(size)]
    errors = []
    for round_ in range(3):
        store = gb.HashStore()
        out = [None] * size

        def rank(r):
            try:
                ctx = gb.Context(r, size)
                ctx.connect_full_mesh(store, devs[r])
                for n in (1, 1000, 3_000_001):
                    x = np.arange(n, dtype=np.float32) + r
                    gb.allreduce(ctx, x)
                    np.testing.assert_allclose(x[[0, -1]], (np.arange(n, dtype=np.float64)[[0, -1]] * size + 6))
                a = np.full(70_000, float(r), np.float64)
                b = np.zeros_like(a)
                right, left = (r + 1) % size, (r - 1) % size
                ub_s, ub_r = ctx.create_unbound_buffer(a.ctypes.data, a.nbytes), ctx.create_unbound_buffer(b.ctypes.data, b.nbytes)
                ub_r.recv(left, 7)
                ub_s.send(right, 7)
                ub_r.wait_recv()
                ub_s.wait_send()
                assert b[0] == left and b[-1] == left
                gb.barrier(ctx)
                ctx.close_connections()
                out[r] = True
            except Exception as e:  # noqa: BLE001
                errors.append((round_, r, repr(e)))

        ths = [threading.Thread(target=rank, args=(r,)) for r in range(size)]
        [t.start() for t in ths]
        [t.join(120) for t in ths]
        assert not errors and out == [True] * size, (errors, out)


def _build_fake_ibverbs():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    out = os.path.join(tempfile.mkdtemp(prefix="glb_fakeibv_"), "libfakeibverbs.so")
    subprocess.check_call([cxx, "-shared", "-fPIC", "-std=c++17", "-O1", "-pthread", f"-I{os.path.join(ROOT, 'csrc')}",
                           os.path.join(ROOT, "tests", "fake_ibverbs", "fake_ibverbs.cc"), "-o", out])
    return out


@pytest.mark.parametrize("size", [2, 3, 4])
def test_ibverbs_transport_over_software_verbs(size):
    """The verbs data path (RC queue pairs, eager / rendezvous unbound messages, bound buffers over
    RDMA WRITE WITH IMMEDIATE, remote-key put / get) against tests/fake_ibverbs: an in-process
    software provider with the libibverbs ABI (the image has neither rdma-core nor an HCA).
    Parity: gloo/test/{send_recv,remote_key,allreduce}_test.cc run with Transport::IBVERBS."""
    lib = _build_fake_ibverbs()
    env = dict(os.environ, GLB_IBVERBS_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ibverbs_worker.py"), str(size)], env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and f"IBVERBS OK {size}" in out.stdout, (out.stdout[-2000:], out.stderr[-4000:])


def test_ibverbs_on_a_real_hca():
    """Same worker against the real libibverbs when the box has one and an RDMA device."""
    p = _C.ibverbs_probe()
    if not p["library"] or not p["devices"]:
        pytest.skip(f"no RDMA device: {p['detail']}")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ibverbs_worker.py"), "2"], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0 and "IBVERBS OK 2" in out.stdout, (out.stdout[-2000:], out.stderr[-4000:])


def test_ibverbs_probe_and_error():
    p = _C.ibverbs_probe()
    assert set(p) == {"library", "devices", "peer_memory_module", "detail"} and p["detail"]
    if hasattr(_C, "ibverbs_device_names"):
        assert _C.ibverbs_device_names() == p["devices"]
    with pytest.raises(gb.InvalidOperationError) as e:
        _C.create_ibverbs_device()
    assert "ibverbs" in str(e.value) and ("tcp" in str(e.value))


def run_bench(size, *args, env=None, timeout=120):
    d = tempfile.mkdtemp(prefix="glb_cli_")
    procs = [subprocess.Popen([BENCH, "--size", str(size), "--rank", str(r), "--shared-path", d, *args],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
             for r in range(size)]
    outs = [p.communicate(timeout=timeout)[0] for p in procs]
    return [p.returncode for p in procs], outs


@pytest.mark.skipif(not os.path.exists(BENCH), reason="benchmark binary not built")
@pytest.mark.parametrize("transport", ["tcp", "uv"])
def test_benchmark_cli(transport):
    codes, outs = run_bench(2, "--transport", transport, "--elements", "1000", "--iteration-count", "20",
                            "allreduce_ring_chunked")
    assert codes == [0, 0], outs
    row = [l for l in outs[0].splitlines() if l.strip().startswith("4000")]
    assert row, outs[0]  # size (B) = 1000 float32
    assert int(row[0].split()[-1]) == 20  # iterations column


@pytest.mark.skipif(not os.path.exists(BENCH), reason="benchmark binary not built")
def test_benchmark_cli_sync_threads_and_errors():
    codes, outs = run_bench(2, "--sync=true", "--busy-poll=true", "--threads", "2", "--elements", "500",
                            "--iteration-count", "10", "allreduce_halving_doubling")
    assert codes == [0, 0], outs
    codes, outs = run_bench(1, "--transport", "ibverbs", "--elements", "10", "allreduce_ring")
    assert codes[0] != 0 and "ibverbs transport" in outs[0]
    codes, outs = run_bench(1, "--transport", "carrier-pigeon", "--elements", "10", "allreduce_ring")
    assert codes[0] != 0 and "unknown transport" in outs[0]


@pytest.mark.skipif(not os.path.exists(BENCH) or shutil.which("openssl") is None or not _C.tls_available(),
                    reason="needs the benchmark binary and OpenSSL")
def test_benchmark_cli_tls():
    from test_tls import make_ca, make_cert

    d = tempfile.mkdtemp(prefix="glb_cli_tls_")
    ca_key, ca_crt = make_ca(d, "bench")
    key, crt = make_cert(d, "rank", ca_key, ca_crt)
    codes, outs = run_bench(2, "--transport", "tls", "--pkey", key, "--cert", crt, "--ca-file", ca_crt,
                            "--elements", "1000", "--iteration-count", "10", "new_allreduce_ring")
    assert codes == [0, 0], outs
    codes, outs = run_bench(1, "--transport", "tls", "--elements", "10", "allreduce_ring")
    assert codes[0] != 0 and "--pkey" in outs[0]


@pytest.mark.skipif(shutil.which("g++") is None and not os.path.exists("/usr/bin/g++"), reason="no host compiler")
def test_mpi_bootstrap_against_thread_world_mpi():
    """csrc/glb/mpi/context.cc is only compiled when mpi.h exists (not in this image). Build it
    here against tests/fake_mpi (threads as the MPI world) and run the whole bootstrap:
    MPI_Allreduce(MAX) + MPI_Allgather of the rendezvous blobs -> TCP full mesh -> allreduce."""
    lib = os.path.join(ROOT, "gloo_b200", "lib")
    if not os.path.exists(os.path.join(lib, "libglb.so")):
        pytest.skip("libglb.so not built")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = os.path.join(tempfile.mkdtemp(prefix="glb_mpi_"), "mpi_bootstrap")
    fm = os.path.join(ROOT, "tests", "fake_mpi")
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-pthread", "-DGLB_USE_MPI=1", "-DGLB_USE_CUDA=1",
                           f"-I{fm}", f"-I{os.path.join(ROOT, 'csrc')}", "-I/usr/local/cuda/include",
                           os.path.join(fm, "mpi_bootstrap_main.cc"), os.path.join(fm, "fake_mpi.cc"),
                           os.path.join(ROOT, "csrc", "glb", "mpi", "context.cc"),
                           f"-L{lib}", "-lglb", f"-Wl,-rpath,{lib}", "-o", exe])
    for p in (2, 5):
        out = subprocess.run([exe, str(p)], capture_output=True, text=True, timeout=60)
        assert out.returncode == 0 and out.stdout.startswith(f"OK {p * (p + 1) // 2}"), (out.stdout, out.stderr)


SELFTEST = os.path.join(ROOT, "gloo_b200", "bin", "glb_selftest")


@pytest.mark.skipif(not os.path.exists(SELFTEST), reason="selftest binary not built")
def test_cpp_selftest():
    """The C++ API without Python in the loop: threads as ranks, every host collective family,
    closed-form checks (csrc/glb/benchmark/selftest_main.cc)."""
    out = subprocess.run([SELFTEST, "1", "2", "3", "5", "8"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "PASS (0 failures)" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])


def test_device_on_a_real_interface():
    """Bind the first non-loopback interface that has an address (what a multi-node job does
    with GLB_SOCKET_IFNAME) and run a collective through it."""
    import threading

    cands = [n for n in sorted(os.listdir("/sys/class/net")) if n != "lo"]
    dev = None
    for name in cands:
        try:
            dev = gb.create_device(hostname="", iface=name)
            break
        except gb.GlbError:
            continue
    if dev is None:
        pytest.skip("no non-loopback interface with an address")
    assert "127.0.0.1" not in str(dev)
    store, out = gb.HashStore(), [None, None]

    def rank(r):
        ctx = gb.Context(r, 2)
        ctx.connect_full_mesh(store, gb.create_device(hostname="", iface=name))
        x = np.full(1 << 17, r + 1, np.float32)  # large enough for the single-copy path
        gb.allreduce(ctx, x)
        gb.barrier(ctx)
        out[r] = float(x[-1])

    ths = [threading.Thread(target=rank, args=(r,)) for r in range(2)]
    [t.start() for t in ths]
    [t.join(60) for t in ths]
    assert out == [3.0, 3.0]


def test_cpp_example_quick_collective_then_close():
    """examples/example_reduce.cc: connect, one tiny reduce, closeConnections(), exit - ranks
    finish at different times, so late protocol chatter (capability frames) and the abortive
    close of a fast rank must not turn into errors on the slower ones."""
    lib = os.path.join(ROOT, "gloo_b200", "lib")
    if not os.path.exists(os.path.join(lib, "libglb.so")):
        pytest.skip("libglb.so not built")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = os.path.join(tempfile.mkdtemp(prefix="glb_ex_"), "example_reduce")
    subprocess.check_call([cxx, "-std=c++17", "-pthread", f"-I{os.path.join(ROOT, 'csrc')}",
                           os.path.join(ROOT, "examples", "example_reduce.cc"), f"-L{lib}", "-lglb",
                           f"-Wl,-rpath,{lib}", "-o", exe])
    for _ in range(8):
        d = tempfile.mkdtemp(prefix="glb_ex_rdv_")
        procs = [subprocess.Popen([exe, str(r), "3", d], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
                 for r in range(3)]
        outs = [p.communicate(timeout=60) for p in procs]
        assert [p.returncode for p in procs] == [0, 0, 0], outs
        assert "sum = 6" in outs[0][0]


@pytest.mark.skipif(not os.path.exists(BENCH), reason="benchmark binary not built")
def test_every_host_benchmark_name_runs_and_verifies():
    """Each benchmark the CLI lists (the reference's names plus the new-style additions) runs on
    two ranks with result verification on; cuda_* names need a GPU and are covered by -m gpu."""
    res = subprocess.run([BENCH, "--help"], capture_output=True, text=True)
    listing = res.stdout + res.stderr
    names = [l.strip() for l in listing.split("BENCHMARK is one of:")[1].splitlines() if l.strip()]
    host = [n for n in names if not n.startswith("cuda_")]
    assert len(host) >= 25, names
    bad = []
    for n in host:
        codes, outs = run_bench(2, "--elements", "500", "--iteration-count", "3", "--inputs", "2", n, timeout=60)
        if codes != [0, 0]:
            bad.append((n, codes, outs[0][-300:]))
    assert not bad, bad


def test_ipv6_loopback_device():
    """The transport is address-family agnostic: bind ::1 and run a collective over it."""
    import threading

    try:
        dev = gb.create_device(hostname="::1")
    except gb.GlbError:
        pytest.skip("IPv6 loopback not available")
    assert "[::1]" in str(dev)
    store, out = gb.HashStore(), [None, None]

    def rank(r):
        ctx = gb.Context(r, 2)
        ctx.connect_full_mesh(store, gb.create_device(hostname="::1"))
        x = np.full(1 << 17, r + 1, np.float32)
        gb.allreduce(ctx, x)
        gb.barrier(ctx)
        out[r] = float(x[-1])

    ths = [threading.Thread(target=rank, args=(r,)) for r in range(2)]
    [t.start() for t in ths]
    [t.join(60) for t in ths]
    assert out == [3.0, 3.0]
