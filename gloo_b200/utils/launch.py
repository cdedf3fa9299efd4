"""Launch / rendezvous helpers.

* ``spawn_threads``: threads-as-ranks harness (the reference's BaseTest::spawn,
  gloo/test/base_test.h:117-179): N threads, each with its own device + context,
  rendezvous through one in-process HashStore, real TCP over loopback.
* ``init_context``: build a connected context from a store (FileStore path,
  explicit Store, or the torchrun environment).
"""
from __future__ import annotations

import os
import threading
import traceback
from typing import Any, Callable, List, Optional

from .. import _C


def create_device(hostname: str = "127.0.0.1", iface: str = "", lazy: bool = False, num_loops: int = 1,
                  nvl: bool = False, cuda_device: int = -1):
    """A transport device. ``nvl=True`` wraps the TCP device in the NVLink peer-memory
    transport: CUDA pointers become unbound buffers whose send / recv / put / get run over
    NVLink (``hasGPUDirect``), host pointers keep using TCP."""
    dev = _C.create_tcp_device(hostname, iface, lazy, num_loops)
    if nvl:
        dev = _C.create_nvl_device(dev, cuda_device)
    return dev


def init_context(rank: int, size: int, store=None, path: Optional[str] = None, prefix: Optional[str] = None,
                 device=None, base: int = 2, timeout_ms: Optional[int] = None):
    """Create a full-mesh connected context.

    Exactly one of ``store`` / ``path`` must be given; ``prefix`` namespaces the
    keys so one store can serve several contexts.
    """
    if store is None:
        if path is None:
            raise ValueError("init_context needs a store or a FileStore path")
        store = _C.FileStore(path)
    if prefix is not None:
        store = _C.PrefixStore(prefix, store)
    if device is None:
        device = create_device()
    ctx = _C.Context(rank, size, base)
    if timeout_ms is not None:
        ctx.set_timeout(timeout_ms)
    ctx.connect_full_mesh(store, device)
    return ctx


def spawn_threads(size: int, fn: Callable[..., Any], *args, base: int = 2, timeout_ms: int = 30000,
                  lazy: bool = False, shared_device: bool = False, cuda_device: Optional[int] = None,
                  nvl: bool = False, **kwargs) -> List[Any]:
    """Run ``fn(ctx, *args, **kwargs)`` on ``size`` threads acting as ranks; returns per-rank results.

    ``cuda_device``: make that GPU current in every rank thread and give each rank
    its own CUDA stream. Ranks that share a process must never share a stream: the
    collective kernels of different ranks wait for each other on the device, so
    they have to be able to run concurrently (the legacy default stream would
    serialise them and deadlock).
    """
    store = _C.HashStore()
    results: List[Any] = [None] * size
    errors: List[Optional[BaseException]] = [None] * size
    shared = create_device(lazy=lazy) if shared_device else None

    def run(rank: int):
        try:
            dev = shared if shared is not None else create_device(lazy=lazy)
            if nvl:
                import torch

                torch.cuda.set_device(cuda_device or 0)
                dev = _C.create_nvl_device(dev, cuda_device or 0)
            ctx = _C.Context(rank, size, base)
            ctx.set_timeout(timeout_ms)
            ctx.connect_full_mesh(store, dev)
            try:
                if cuda_device is not None:
                    import torch

                    torch.cuda.set_device(cuda_device)
                    # A dedicated stream (torch.cuda.Stream() comes from a shared pool).
                    raw = _C.cuda.create_stream(cuda_device)
                    try:
                        with torch.cuda.stream(torch.cuda.ExternalStream(raw, device=cuda_device)):
                            results[rank] = fn(ctx, *args, **kwargs)
                            torch.cuda.current_stream().synchronize()
                    finally:
                        _C.cuda.destroy_stream(raw)
                else:
                    results[rank] = fn(ctx, *args, **kwargs)
            finally:
                # Leave together so nobody tears down sockets a peer still needs.
                try:
                    _C.barrier(ctx, 0xFFFFFF, min(timeout_ms, 10000))
                except Exception:
                    pass
                ctx.close_connections()
        except BaseException as e:  # noqa: BLE001
            errors[rank] = e
            traceback.print_exc()

    threads = [threading.Thread(target=run, args=(r,), name=f"rank{r}") for r in range(size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    return results


def env_rank_size():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
