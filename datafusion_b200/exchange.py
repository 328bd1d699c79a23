"""The partition exchange between GPUs: the stand-in for RepartitionExec's channels
(reference datafusion/physical-plan/src/repartition/mod.rs:1097-1145, 2138-2225) when both join inputs
(or the partial aggregate states) must be co-partitioned across the GPUs of one box
(PartitionMode::Partitioned, hash_join/exec.rs:1312-1325).

  local pass  : libdfgpu dfgpu_hash_partition_device — rows -> n contiguous per-destination regions (CUDA)
  counts      : one tiny all-to-all of the per-destination row counts
  payload     : ONE all-to-all-v per column over NVLink (NCCL through torch.distributed)

torch is plumbing only (process group + NCCL); `backend="gloo"` with CPU tensors exercises the same
host logic in the world_size-2 CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import capi as D

_NP = {D.INT8: "|i1", D.UINT8: "|u1", D.INT16: "<i2", D.UINT16: "<u2", D.INT32: "<i4", D.UINT32: "<u4", D.INT64: "<i8", D.UINT64: "<u8",
       D.FLOAT32: "<f4", D.FLOAT64: "<f8", D.DATE32: "<i4", D.DATE64: "<i8", D.TIMESTAMP: "<i8"}


class _CudaView:
    """zero-copy __cuda_array_interface__ view of a device pointer owned by libdfgpu"""

    def __init__(self, ptr: int, n: int, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2, "strides": None}
        self._owner = owner


def plan_all_to_all(send_counts: Sequence[int], recv_counts: Sequence[int]):
    """split sizes + offsets of one all-to-all-v (pure host logic, unit-tested on CPU)"""
    send_counts = [int(x) for x in send_counts]; recv_counts = [int(x) for x in recv_counts]
    send_offs = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64)
    recv_offs = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
    return send_counts, recv_counts, send_offs, recv_offs


def exchange_counts(dist, send_counts, device):
    import torch
    world = dist.get_world_size()
    s = torch.tensor(list(send_counts), dtype=torch.int64, device=device)
    r = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(r, s)
    return [int(x) for x in r.tolist()]


def all_to_all_columns(dist, send_tensors, send_counts, recv_counts):
    """one all_to_all_single (all-to-all-v) per column tensor; returns the received tensors"""
    import torch
    out = []
    total = int(sum(recv_counts))
    for t in send_tensors:
        r = torch.empty(total, dtype=t.dtype, device=t.device)
        dist.all_to_all_single(r, t, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts))
        out.append(r)
    return out


class ExchangedBatch:
    def __init__(self, ctx, tensors, types, rows, keep):
        self.ctx, self.tensors, self.types, self.rows, self._keep = ctx, tensors, types, rows, keep

    def columns(self) -> List[D.Column]:
        cols = []
        for t, ty in zip(self.tensors, self.types):
            c = D.Column()
            c.type, c.flags, c.length, c.offset, c.null_count = ty, 0, self.rows, 0, 0
            c.values = t.data_ptr()
            c.validity = None
            cols.append(c)
        return cols


def exchange_batch(ctx: D.Context, cols, key_cols: Sequence[int], dist) -> ExchangedBatch:
    """hash-partition `cols` (device, no NULLs) on `key_cols` across the process group and exchange them"""
    import torch
    world = dist.get_world_size()
    batch, offs = D.hash_partition_device(ctx, cols, list(key_cols), world)
    send_counts = [offs[p + 1] - offs[p] for p in range(world)]
    dev = torch.device("cuda", ctx.device)
    shared_stream = (ctx.lib.dfgpu_ctx_stream(ctx.h) or 0) == torch.cuda.current_stream().cuda_stream
    if not shared_stream:
        ctx.sync()  # partitioned buffers must be complete before NCCL (ordered on torch's stream) reads them
    recv_counts = exchange_counts(dist, send_counts, dev)
    send_tensors, types = [], []
    for i in range(batch.num_columns):
        c = batch.column(i)
        if c.validity:
            raise NotImplementedError("exchange of nullable columns is not implemented yet")
        view = _CudaView(c.values, max(c.length, 1), _NP[c.type], batch)
        send_tensors.append(torch.as_tensor(view, device=dev)[: c.length])
        types.append(c.type)
    recv = all_to_all_columns(dist, send_tensors, send_counts, recv_counts)
    if not shared_stream:
        torch.cuda.current_stream().synchronize()
    return ExchangedBatch(ctx, recv, types, int(sum(recv_counts)), batch)
