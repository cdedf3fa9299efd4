"""Consumers of the collectives — the parallelism strategies of SURVEY §2.18.

The reference implements none of these (it is the substrate they call); each helper
here is a thin, explicit mapping from a strategy to the collective it consumes:

  DataParallel        gradient allreduce in size-capped buckets (+ broadcast of params)
  GradientBucketer    gradients live in symmetric buckets (views), allreduce overlapped with backward
  ZeroShard           ZeRO/FSDP: reduce_scatter gradients, allgather parameters
  ZeroOptimizer       ZeRO-1: flat symmetric param / grad buffers, optimizer state for 1/P of them
  TensorParallel      Megatron column/row-parallel linear: allgather / allreduce; ColumnParallelLinear /
                      RowParallelLinear layers with the collectives as autograd operators (f / g)
  SequenceParallel    Megatron-SP: reduce_scatter + allgather along the sequence
  MoEDispatcher       expert parallel dispatch/combine: alltoallv
  UlyssesAttention    head<->sequence alltoall
  RingExchange        ring-attention / pipeline neighbour send-recv (KV rotation), ring attention forward
  PipelineParallel    GPipe schedule over point-to-point activation / gradient sends

Everything accepts either a CUDA ``CudaContext`` (NVLink kernels) or a host context
(TCP); tensors decide which path runs.
"""
from .strategies import (  # noqa: F401
    ColumnParallelLinear,
    DataParallel,
    GradientBucketer,
    MoEDispatcher,
    PipelineParallel,
    RingExchange,
    RowParallelLinear,
    SequenceParallel,
    TensorParallel,
    UlyssesAttention,
    ZeroOptimizer,
    ZeroShard,
)
