"""Stores, Slot algebra, float16 conversion, CPU math, ContextFactory, Linux probes —
mirrors gloo/test/{math,linux,context_factory,memory}_test.cc and the store semantics
documented in docs/rendezvous.md."""
import os
import struct
import tempfile
import threading
import time

import numpy as np
import pytest

import gloo_b200 as gb
from gloo_b200 import _C


# ---- stores -------------------------------------------------------------------------------

def _exercise_store(make):
    s = make()
    s.set("a", b"hello")
    assert s.get("a") == b"hello"
    with pytest.raises(gb.GlbError):
        s.set("a", b"again")  # keys are write-once
    # blocking get across threads
    out = []
    t = threading.Thread(target=lambda: out.append(make().get("later") if make is not None else None))
    s2 = s
    t = threading.Thread(target=lambda: out.append(s2.get("later")))
    t.start()
    time.sleep(0.05)
    s.set("later", b"x" * 1000)
    t.join()
    assert out == [b"x" * 1000]
    with pytest.raises(gb.IoError):
        s.wait(["never"], 50)
    assert s.add("ctr", 5) == 5
    assert s.add("ctr", -2) == 3
    s.append("log", b"ab")
    s.append("log", b"cd")
    assert s.get("log") == b"abcd"
    assert s.multi_get(["a", "later"]) == [b"hello", b"x" * 1000]


def test_hash_store():
    _exercise_store(gb.HashStore)


def test_file_store():
    d = tempfile.mkdtemp(prefix="glb_fs_")
    _exercise_store(lambda: gb.FileStore(d))
    # a second instance (another "process") sees the same keys
    assert gb.FileStore(d).get("a") == b"hello"
    assert len(os.listdir(d)) >= 3


def test_prefix_store():
    base = gb.HashStore()
    a, b = gb.PrefixStore("A", base), gb.PrefixStore("B", base)
    a.set("k", b"1")
    b.set("k", b"2")
    assert a.get("k") == b"1" and b.get("k") == b"2"
    assert base.get("A/k") == b"1"


def test_python_store_subclass():
    """A store written in Python (e.g. wrapping torch's TCPStore) can drive rendezvous."""
    class DictStore(gb.Store):
        data = {}
        lock = threading.Lock()

        def set(self, key, value):
            with self.lock:
                self.data[key] = bytes(value)

        def get(self, key):
            self.wait([key], 30000)
            with self.lock:
                return self.data[key]

        def wait(self, keys, timeout_ms):
            t0 = time.time()
            while True:
                with self.lock:
                    if all(k in self.data for k in keys):
                        return
                if (time.time() - t0) * 1000 > timeout_ms:
                    raise RuntimeError("timeout")
                time.sleep(0.001)

    res = [None, None]

    def run(rank):
        ctx = gb.init_context(rank, 2, store=DictStore())
        buf = np.full(10, rank + 1.0)
        gb.allreduce(ctx, buf)
        res[rank] = buf[0]
        gb.barrier(ctx)
        ctx.close_connections()

    ts = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert res == [3.0, 3.0]


# ---- Slot / types / math ----------------------------------------------------------------------

def test_slot_algebra():
    s = _C.slot_build(4, 0x12345678, 0)
    assert s >> 56 == 4
    assert (s >> 24) & 0xFFFFFFFF == 0x12345678
    assert _C.slot_build(4, 7, 100) == _C.slot_build(4, 7, 0) + 100
    with pytest.raises(gb.EnforceError):
        _C.slot_build(4, 7, 1 << 24)  # delta overflow
    # distinct prefixes / tags never collide
    assert len({_C.slot_build(p, t, 0) for p in range(1, 10) for t in range(50)}) == 9 * 50


def test_float16_conversion_matches_numpy():
    rng = np.random.RandomState(0)
    vals = np.concatenate([rng.randn(2000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1, 100, 7e4)])
    vals = np.concatenate([vals, np.array([0.0, -0.0, np.inf, -np.inf, 65504, 65520, 6e-8, 5.96e-8, 2.98e-8], np.float32)])
    for v in vals:
        bits = _C.float_to_half_bits(float(v))
        with np.errstate(over="ignore"):  # values beyond 65504 round to inf on purpose
            exp = np.float32(v).astype(np.float16).view(np.uint16)
        assert bits == int(exp), (v, hex(bits), hex(int(exp)))
        assert _C.half_bits_to_float(bits) == float(np.uint16(bits).view(np.float16))
    assert np.isnan(_C.half_bits_to_float(_C.float_to_half_bits(float("nan"))))
    # bfloat16: round to nearest even on the upper 16 bits
    for v in (1.0, 1.00390625, 3.14159, -2.5e-3, 1e38):
        u = struct.unpack("<I", struct.pack("<f", v))[0]
        exp = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
        assert _C.float_to_bfloat_bits(v) == exp


@pytest.mark.parametrize("dtype", [np.float16, np.float32, np.int32, np.int64, np.uint8, np.float64])
@pytest.mark.parametrize("op", list(gb.ReduceOp)[:4])
def test_cpu_reduce_functions(dtype, op):
    from gloo_b200.types import describe

    rng = np.random.RandomState(1)
    for n in (0, 1, 7, 8, 9, 1000, 1003):
        a = rng.randint(1, 5, n).astype(dtype)
        b = rng.randint(1, 5, n).astype(dtype)
        c = np.zeros(n, dtype)
        _, _, dt, _ = describe(np.zeros(1, dtype))
        _C.reduce_local(c.ctypes.data, a.ctypes.data, b.ctypes.data, n, int(dt), int(op))
        f = {gb.ReduceOp.SUM: np.add, gb.ReduceOp.PRODUCT: np.multiply, gb.ReduceOp.MIN: np.minimum,
             gb.ReduceOp.MAX: np.maximum}[op]
        np.testing.assert_array_equal(c, f(a, b).astype(dtype))
    assert isinstance(_C.has_simd_half(), bool)


def test_factorize():
    for n in range(1, 200):
        f = _C.factorize(n)
        assert int(np.prod(f)) == n if f else n == 1


# ---- contexts -----------------------------------------------------------------------------------

def test_context_factory():
    """Derived contexts are minted over the backing mesh without touching the store."""
    size = 4

    def fn(ctx):
        fac = gb.ContextFactory(ctx)
        dev = gb.create_device()
        derived = [fac.make_context(dev) for _ in range(3)]
        vals = []
        for i, c in enumerate(derived):
            assert (c.rank, c.size) == (ctx.rank, ctx.size)
            buf = np.full(5, (ctx.rank + 1) * (i + 1), np.int64)
            gb.allreduce(c, buf)
            vals.append(int(buf[0]))
            gb.barrier(c)
        for c in derived:
            c.close_connections()
        return vals

    res = gb.spawn_threads(size, fn)
    assert all(r == [10, 20, 30] for r in res)


def test_context_accessors_and_timeouts():
    def fn(ctx):
        assert ctx.get_timeout() == 30000
        ctx.set_timeout(1234)
        assert ctx.get_timeout() == 1234
        a, b = ctx.next_slot(), ctx.next_slot(5)
        assert b == a + 1 and ctx.next_slot() == b + 5
        assert "tcp" in str(ctx.device())
        p = ctx.get_pair(1 - ctx.rank)
        assert p.is_connected() and p.local_rank() == ctx.rank  # both ranks live on this host
        return True

    assert all(gb.spawn_threads(2, fn))


def test_many_slots_do_not_leak():
    """10k distinct tags on one context (memory_test.cc, enabled here)."""
    import resource

    def fn(ctx):
        buf = np.ones(4, np.float32)
        for i in range(300):
            gb.allreduce(ctx, buf, tag=i)
            buf[:] = 1
        before = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
        for i in range(300, 5300):
            gb.allreduce(ctx, buf, tag=i)
            buf[:] = 1
        after = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
        return after - before

    growth_kb = max(gb.spawn_threads(2, fn))
    assert growth_kb < 20000


# ---- linux probes ------------------------------------------------------------------------------------

def test_linux_probes():
    ifs = _C.list_interfaces()
    assert "lo" in ifs
    assert _C.interface_to_bus_id("lo") == ""
    assert _C.interface_speed("lo") in (-1,) or _C.interface_speed("lo") > 0
    assert isinstance(_C.kernel_modules(), list)
    nets = _C.pci_devices(0x020000, 0xFF0000)
    for b in nets:
        assert _C.pci_distance(b, b) == 0
    assert _C.pci_distance("ffff:ff:ff.f", "ffff:ff:ff.e") == -1
    assert _C.hostname()


def test_redis_store_against_resp_server():
    """RedisStore speaks RESP itself (no hiredis): write-once SETNX keys, GET, EXISTS-polling
    wait with timeout, binary-safe values, the v2 API, and a full-mesh rendezvous through it
    (reference: gloo/rendezvous/redis_store.cc:35-119)."""
    import subprocess
    import sys

    from fake_redis import resp_call

    here = os.path.dirname(os.path.abspath(__file__))
    srv = subprocess.Popen([sys.executable, os.path.join(here, "fake_redis.py")], stdout=subprocess.PIPE, text=True)
    try:
        port = int(srv.stdout.readline())
        s = gb.RedisStore("127.0.0.1", port)
        blob = bytes(range(256)) * 3 + b"\r\n$5\r\n"  # framing characters inside the value
        s.set("k", blob)
        assert s.get("k") == blob
        with pytest.raises(gb.GlbError):
            s.set("k", b"again")  # keys are write-once
        assert s.check(["k"]) and not s.check(["k", "missing"])
        t0 = time.time()
        with pytest.raises(gb.IoError):
            s.wait(["missing"], 150)
        assert 0.1 < time.time() - t0 < 2.0
        threading.Timer(0.1, lambda: resp_call(port, b"SET", b"late", b"x")).start()
        s.wait(["late"], 5000)
        assert s.has_extended_api()
        assert s.add("ctr", 5) == 5 and s.add("ctr", -2) == 3
        s.append("log", b"ab")
        s.append("log", b"cd")
        assert s.get("log") == b"abcd"
        assert s.multi_get(["k", "log"]) == [blob, b"abcd"]

        # rendezvous through the store, one client connection per rank, namespaced by PrefixStore
        def rank(r, out):
            store = gb.PrefixStore("job42", gb.RedisStore("127.0.0.1", port))
            ctx = gb.Context(r, 3)
            ctx.connect_full_mesh(store, gb.create_device())
            x = np.full(16, r + 1, np.float32)
            gb.allreduce(ctx, x)
            out[r] = float(x[0])
            gb.barrier(ctx)

        out = [None] * 3
        ths = [threading.Thread(target=rank, args=(r, out)) for r in range(3)]
        [t.start() for t in ths]
        [t.join(60) for t in ths]
        assert out == [6.0, 6.0, 6.0]
        assert any(k.startswith(b"job42/") for k in resp_call(port, b"KEYS", b"*"))
        assert {b"SETNX", b"GET", b"EXISTS"} <= set(resp_call(port, b"_STATS"))
    finally:
        srv.kill()
        srv.wait()


def test_connect_timeouts():
    """Reference tcp_test.cc:12 / base_test: (1) a peer that never publishes its address makes
    rendezvous fail after the context timeout; (2) an address that is published but not served
    (connection refused, retried) fails after the timeout as well, and with
    connection retries disabled it fails on the first attempt."""
    # (1) rank 1 never arrives
    store = gb.HashStore()
    ctx = gb.Context(0, 2)
    ctx.set_timeout(300)
    t0 = time.time()
    with pytest.raises(gb.IoError):
        ctx.connect_full_mesh(store, gb.create_device())
    assert 0.25 < time.time() - t0 < 5

    # (2) rank 0's blob from a finished run points at a port nobody listens on any more
    d = tempfile.mkdtemp(prefix="glb_stale_")

    def once(rank):
        c = gb.init_context(rank, 2, path=d, timeout_ms=5000)
        gb.barrier(c)
        c.close_connections()

    ths = [threading.Thread(target=once, args=(r,)) for r in range(2)]
    [t.start() for t in ths]
    [t.join(30) for t in ths]
    # a fresh rank 1 finds rank 0's stale address in its store and dials a dead port
    stale = gb.HashStore()
    stale.set("0", gb.FileStore(d).get("0"))
    c1 = gb.Context(1, 2)
    c1.set_timeout(400)
    t0 = time.time()
    with pytest.raises(gb.IoError):
        c1.connect_full_mesh(stale, gb.create_device())
    assert time.time() - t0 < 10


# ---- property tests -------------------------------------------------------------------------------

try:
    from hypothesis import given, settings
    from hypothesis import strategies as st
except ImportError:  # pragma: no cover - hypothesis is optional
    given = None

if given is not None:

    @settings(max_examples=500, deadline=None)
    @given(st.floats(width=32, allow_nan=False))
    def test_half_conversion_property(v):
        """float -> half is numpy's round-to-nearest-even for every finite or infinite float32,
        and half -> float is exact."""
        bits = _C.float_to_half_bits(float(v))
        with np.errstate(over="ignore"):
            want = np.float32(v).astype(np.float16)
        assert bits == int(want.view(np.uint16))
        assert _C.half_bits_to_float(bits) == float(want) or (np.isnan(want) and np.isnan(_C.half_bits_to_float(bits)))

    @settings(max_examples=300, deadline=None)
    @given(st.floats(width=32, allow_nan=False, allow_infinity=False))
    def test_bfloat_conversion_property(v):
        """float -> bfloat16 rounds to nearest even on the upper 16 bits (torch agrees)."""
        import torch

        want = torch.tensor([v], dtype=torch.float32).to(torch.bfloat16).view(torch.int16).item() & 0xFFFF
        assert _C.float_to_bfloat_bits(float(v)) == want

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 100000))
    def test_factorize_property(n):
        f = _C.factorize(n)
        assert all(p >= 2 for p in f) and f == sorted(f)
        assert int(np.prod(f, dtype=np.int64)) == n if f else n == 1
        assert all(all(p % q for q in range(2, int(p ** 0.5) + 1)) for p in f)  # primes


def test_metrics_snapshot_and_prometheus_endpoint():
    """utils.metrics: counters move with traffic and are served in Prometheus text format."""
    import socket
    import urllib.request

    import numpy as np

    from gloo_b200.utils import metrics

    before = metrics.snapshot()
    assert {"glb_tcp_cma_messages_total", "glb_tcp_cma_bytes_total", "glb_tcp_spin_budget_us"} <= set(before)

    def fn(ctx):
        x = np.ones(1 << 20, np.float32)           # 4 MB: above the single-copy threshold between same-host ranks
        gb.allreduce(ctx, x)
        return float(x[0])

    assert gb.spawn_threads(2, fn) == [2.0, 2.0]
    after = metrics.snapshot()
    assert after["glb_tcp_cma_bytes_total"] >= before["glb_tcp_cma_bytes_total"]
    pytest.importorskip("prometheus_client")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    metrics.start_exporter(port)
    body = urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=10).read().decode()
    assert "glb_tcp_cma_bytes_total" in body and "glb_build_info" in body and 'cuda_arch=' in body


def test_chrome_trace_of_collective_calls(tmp_path):
    """GLB_TRACE_FILE: every collective call becomes a chrome-trace complete event (read at process exit)."""
    import json
    import subprocess
    import sys

    code = (
        "import numpy as np, gloo_b200 as gb\n"
        "assert gb._C.trace_enabled()\n"
        "def fn(ctx):\n"
        "    x = np.ones(1000, np.float32)\n"
        "    for _ in range(3):\n"
        "        gb.allreduce(ctx, x)\n"
        "    gb.barrier(ctx)\n"
        "    out = np.zeros(2000, np.float32)\n"
        "    gb.allgather(ctx, out, np.ones(1000, np.float32))\n"
        "    return True\n"
        "assert all(gb.spawn_threads(2, fn))\n"
        "print('flushed', gb._C.trace_flush())\n")
    env = dict(os.environ, GLB_TRACE_FILE=str(tmp_path / "trace_%r.json"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    files = list(tmp_path.iterdir())
    assert len(files) == 1 and files[0].name.startswith("trace_") and "%r" not in files[0].name
    events = json.loads(files[0].read_text().rstrip().rstrip(",") + "]")
    names = [e["name"] for e in events]
    assert names.count("glb::allreduce") == 6 and "glb::allgather" in names and "glb::barrier" in names
    assert all(e["ph"] == "X" and e["dur"] >= 0 and e["pid"] > 0 for e in events)
    assert len({e["tid"] for e in events}) >= 2          # two rank threads
    assert not gb._C.trace_enabled()                     # off in this process
