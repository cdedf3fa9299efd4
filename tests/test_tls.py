"""TLS transport — mirrors gloo/test/tls_tcp_test.cc:25-91 (+ openssl_utils.cc, which
generates the certificates): mutual authentication works with a common CA and fails
with a foreign one."""
import os
import shutil
import subprocess
import tempfile
import threading

import numpy as np
import pytest

import gloo_b200 as gb
from gloo_b200 import _C

pytestmark = pytest.mark.skipif(not _C.tls_available() or shutil.which("openssl") is None,
                                reason="OpenSSL not available")


def make_ca(d, name):
    key, crt = os.path.join(d, f"{name}_ca.key"), os.path.join(d, f"{name}_ca.crt")
    subprocess.check_call(["openssl", "req", "-x509", "-newkey", "rsa:2048", "-nodes", "-keyout", key, "-out", crt,
                           "-subj", f"/CN={name}-ca", "-days", "2"], stderr=subprocess.DEVNULL)
    return key, crt


def make_cert(d, name, ca_key, ca_crt):
    key, csr, crt = (os.path.join(d, f"{name}.{e}") for e in ("key", "csr", "crt"))
    subprocess.check_call(["openssl", "req", "-newkey", "rsa:2048", "-nodes", "-keyout", key, "-out", csr,
                           "-subj", f"/CN={name}"], stderr=subprocess.DEVNULL)
    subprocess.check_call(["openssl", "x509", "-req", "-in", csr, "-CA", ca_crt, "-CAkey", ca_key, "-CAcreateserial",
                           "-out", crt, "-days", "2"], stderr=subprocess.DEVNULL)
    return key, crt


def run_ranks(devs, fn, timeout_ms=5000):
    store = gb.HashStore()
    size = len(devs)
    res, errs = [None] * size, [None] * size

    def run(r):
        try:
            ctx = _C.Context(r, size, 2)
            ctx.set_timeout(timeout_ms)
            ctx.connect_full_mesh(store, devs[r]())
            res[r] = fn(ctx)
            gb.barrier(ctx)
            ctx.close_connections()
        except BaseException as e:  # noqa: BLE001
            errs[r] = e

    ts = [threading.Thread(target=run, args=(r,)) for r in range(size)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return res, errs


def test_tls_allreduce_and_p2p():
    d = tempfile.mkdtemp(prefix="glb_tls_")
    ca_key, ca_crt = make_ca(d, "good")
    certs = [make_cert(d, f"rank{r}", ca_key, ca_crt) for r in range(3)]
    devs = [lambda r=r: _C.create_tls_device("127.0.0.1", certs[r][0], certs[r][1], ca_crt) for r in range(3)]

    def fn(ctx):
        assert str(ctx.device()).startswith("tls+tcp")
        out = []
        for n in (1, 1000, 3_000_000):  # the last one spans many TLS records and partial writes
            buf = np.full(n, ctx.rank + 1, np.float32)
            gb.allreduce(ctx, buf)
            out.append(float(buf[0]) == 6.0 and float(buf[-1]) == 6.0)
        return all(out)

    res, errs = run_ranks(devs, fn, timeout_ms=20000)
    assert errs == [None] * 3, errs
    assert res == [True] * 3


def test_tls_rejects_foreign_ca():
    d = tempfile.mkdtemp(prefix="glb_tls_")
    good_key, good_crt = make_ca(d, "good")
    evil_key, evil_crt = make_ca(d, "evil")
    c0 = make_cert(d, "rank0", good_key, good_crt)
    c1 = make_cert(d, "rank1", evil_key, evil_crt)  # signed by a CA rank 0 does not trust
    devs = [lambda: _C.create_tls_device("127.0.0.1", c0[0], c0[1], good_crt),
            lambda: _C.create_tls_device("127.0.0.1", c1[0], c1[1], evil_crt)]
    res, errs = run_ranks(devs, lambda ctx: True, timeout_ms=2000)
    assert all(isinstance(e, gb.IoError) for e in errs), errs


def test_tls_bad_files():
    d = tempfile.mkdtemp(prefix="glb_tls_")
    ca_key, ca_crt = make_ca(d, "good")
    key, crt = make_cert(d, "rank0", ca_key, ca_crt)
    with pytest.raises(gb.GlbError):
        _C.create_tls_device("127.0.0.1", os.path.join(d, "missing.key"), crt, ca_crt)
    with pytest.raises(gb.GlbError):
        _C.create_tls_device("127.0.0.1", ca_key, crt, ca_crt)  # key does not match the certificate
    with pytest.raises(gb.GlbError):
        _C.create_tls_device("127.0.0.1", key, crt, "", "")      # no CA
