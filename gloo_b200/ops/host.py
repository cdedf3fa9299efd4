"""Host-buffer collectives (new-style API) for numpy arrays and CPU torch tensors.

Each function mirrors a free function of the reference (gloo/allreduce.h, ...) and
is a thin shim over the native implementation in csrc/glb/*.cc.
"""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence, Union

from .. import _C
from ..types import Algorithm, DataType, ReduceOp, describe, element_size

Buffer = Any


def _as_list(x) -> List[Any]:
    return list(x) if isinstance(x, (list, tuple)) else [x]


def _host(x):
    ptr, n, dt, cuda = describe(x)
    if cuda:
        raise TypeError("host collective called with a CUDA tensor; use gloo_b200.ops.cuda")
    return ptr, n, dt


def allreduce(ctx, outputs: Union[Buffer, Sequence[Buffer]], inputs: Union[None, Buffer, Sequence[Buffer]] = None,
              op: Union[ReduceOp, Callable] = ReduceOp.SUM, algorithm: Algorithm = Algorithm.UNSPECIFIED,
              tag: int = 0, timeout_ms: int = -1, max_segment_size: int = 0) -> None:
    """Element-wise reduce across ranks (and across the local buffers); every output gets the result.

    With ``inputs=None`` the operation is in place on ``outputs``. ``op`` may be a
    ReduceOp or a callable ``fn(c_ptr, a_ptr, b_ptr, n)`` (custom reduction).
    """
    outs = _as_list(outputs)
    ins = _as_list(inputs) if inputs is not None else []
    optr, n, dt = _host(outs[0])
    optrs = [optr] + [_host(o)[0] for o in outs[1:]]
    iptrs = [_host(i)[0] for i in ins]
    custom = None
    opv = op
    if callable(op) and not isinstance(op, ReduceOp):
        custom, opv = op, ReduceOp.CUSTOM
    _C.allreduce(ctx, iptrs, optrs, n, int(dt), int(opv), int(algorithm), tag, timeout_ms, max_segment_size, custom)


def reduce(ctx, output: Buffer, input: Optional[Buffer] = None, root: int = 0,
           op: Union[ReduceOp, Callable] = ReduceOp.SUM, tag: int = 0, timeout_ms: int = -1) -> None:
    optr, n, dt = _host(output)
    iptr = _host(input)[0] if input is not None else 0
    custom = None
    opv = op
    if callable(op) and not isinstance(op, ReduceOp):
        custom, opv = op, ReduceOp.CUSTOM
    _C.reduce(ctx, iptr, optr, n, int(dt), int(opv), root, tag, timeout_ms, custom)


def reduce_scatter(ctx, output: Buffer, input: Buffer, recv_counts: Optional[Sequence[int]] = None,
                   op: ReduceOp = ReduceOp.SUM, tag: int = 0, timeout_ms: int = -1) -> None:
    iptr, n, dt = _host(input)
    optr, _, _ = _host(output)
    _C.reduce_scatter(ctx, iptr, optr, n, list(recv_counts or []), int(dt), int(op), tag, timeout_ms)


def broadcast(ctx, output: Buffer, input: Optional[Buffer] = None, root: int = 0, tag: int = 0,
              timeout_ms: int = -1) -> None:
    optr, n, dt = _host(output)
    iptr = _host(input)[0] if input is not None else 0
    _C.broadcast(ctx, iptr, optr, n * element_size(dt), root, tag, timeout_ms)


def allgather(ctx, output: Buffer, input: Optional[Buffer] = None, tag: int = 0, timeout_ms: int = -1) -> None:
    optr, on, dt = _host(output)
    es = element_size(dt)
    if input is not None:
        iptr, inn, _ = _host(input)
    else:
        iptr, inn = 0, 0
    _C.allgather(ctx, iptr, inn * es, optr, on * es, tag, timeout_ms)


def allgatherv(ctx, output: Buffer, counts: Sequence[int], input: Optional[Buffer] = None, tag: int = 0,
               timeout_ms: int = -1) -> None:
    optr, _, dt = _host(output)
    iptr = _host(input)[0] if input is not None else 0
    _C.allgatherv(ctx, iptr, optr, list(counts), element_size(dt), tag, timeout_ms)


def alltoall(ctx, output: Buffer, input: Buffer, tag: int = 0, timeout_ms: int = -1) -> None:
    iptr, n, dt = _host(input)
    optr, _, _ = _host(output)
    _C.alltoall(ctx, iptr, optr, n * element_size(dt), tag, timeout_ms)


def alltoallv(ctx, output: Buffer, out_counts: Sequence[int], input: Buffer, in_counts: Sequence[int],
              tag: int = 0, timeout_ms: int = -1) -> None:
    iptr, _, dt = _host(input)
    optr, _, _ = _host(output)
    _C.alltoallv(ctx, iptr, list(in_counts), optr, list(out_counts), element_size(dt), tag, timeout_ms)


def gather(ctx, input: Buffer, output: Optional[Buffer] = None, root: int = 0, tag: int = 0,
           timeout_ms: int = -1) -> None:
    iptr, n, dt = _host(input)
    optr = _host(output)[0] if output is not None else 0
    _C.gather(ctx, iptr, n * element_size(dt), optr, root, tag, timeout_ms)


def gatherv(ctx, input: Buffer, output: Optional[Buffer] = None, counts: Optional[Sequence[int]] = None,
            root: int = 0, tag: int = 0, timeout_ms: int = -1) -> None:
    iptr, n, dt = _host(input)
    optr = _host(output)[0] if output is not None else 0
    _C.gatherv(ctx, iptr, n, optr, list(counts or []), element_size(dt), root, tag, timeout_ms)


def scatter(ctx, output: Buffer, inputs: Optional[Sequence[Buffer]] = None, root: int = 0, tag: int = 0,
            timeout_ms: int = -1) -> None:
    optr, n, dt = _host(output)
    iptrs = [_host(i)[0] for i in (inputs or [])]
    _C.scatter(ctx, iptrs, optr, n * element_size(dt), root, tag, timeout_ms)


def barrier(ctx, tag: int = 0, timeout_ms: int = -1) -> None:
    _C.barrier(ctx, tag, timeout_ms)
