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


class PeerExchange:
    """Fused partition + exchange over NVLink peer memory (no NCCL payload transfer).

    Every rank owns one persistent receive buffer per column (capacity `cap_rows`), exported once through CUDA IPC;
    an exchange is: count rows per destination (CUDA) -> all-gather the world x world count matrix (tiny NCCL
    collective, which also orders this exchange after everybody's previous use of the buffers) -> ONE scatter
    kernel that writes each row directly into its owner's buffer -> a one-element all-reduce as the completion
    barrier.  Rows arrive grouped by source rank, in source order (same layout as the all-to-all path)."""

    def __init__(self, ctx: D.Context, dist, col_types: Sequence[int], cap_rows: int):
        import torch
        self.ctx, self.dist, self.types, self.cap = ctx, dist, list(col_types), int(cap_rows)
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.dev = torch.device("cuda", ctx.device)
        self.bufs = [D.DeviceBuffer(ctx, self.cap * D.WIDTH[t]) for t in self.types]
        import ctypes as C
        mine = torch.zeros(len(self.types) * 64, dtype=torch.uint8)
        for i, b in enumerate(self.bufs):
            h = C.create_string_buffer(64)
            ctx.check(ctx.lib.dfgpu_ipc_export(ctx.h, C.c_void_p(b.ptr), h))
            mine[i * 64:(i + 1) * 64] = torch.frombuffer(bytearray(h.raw), dtype=torch.uint8)
        allh = [torch.zeros_like(mine).to(self.dev) for _ in range(self.world)]
        dist.all_gather(allh, mine.to(self.dev))
        self.peer_ptrs = []   # [rank][col]
        for r in range(self.world):
            ptrs = []
            hb = bytes(allh[r].cpu().numpy().tobytes())
            for i in range(len(self.types)):
                if r == self.rank:
                    ptrs.append(self.bufs[i].ptr)
                else:
                    out = C.c_void_p()
                    ctx.check(ctx.lib.dfgpu_ipc_import(ctx.h, hb[i * 64:(i + 1) * 64], C.byref(out)))
                    ptrs.append(out.value)
            self.peer_ptrs.append(ptrs)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.dev)

    def exchange(self, cols, key_cols: Sequence[int]) -> "ExchangedPeerBatch":
        import ctypes as C
        import torch
        ctx, world, nc = self.ctx, self.world, len(self.types)
        arr = D._cols(cols)
        counts = (C.c_int64 * world)()
        plan = C.c_void_p()
        ctx.check(ctx.lib.dfgpu_partition_plan_create(ctx.h, arr, nc, D._i32arr(list(key_cols)), len(key_cols), world, counts, C.byref(plan)))
        try:
            mine = torch.tensor(list(counts), dtype=torch.int64, device=self.dev)
            allc = torch.empty(world * world, dtype=torch.int64, device=self.dev)
            self.dist.all_gather_into_tensor(allc, mine)          # [src][dst]; also the "buffers are free again" barrier
            m = allc.view(world, world).cpu().numpy()
            recv_rows = int(m[:, self.rank].sum())
            if m.sum(axis=0).max() > self.cap:
                raise RuntimeError(f"PeerExchange: a receive buffer would overflow ({int(m.sum(axis=0).max())} rows > capacity {self.cap})")
            dst_row = (C.c_int64 * world)(*[int(m[:self.rank, p].sum()) for p in range(world)])   # my block starts after lower ranks' blocks
            bases = (C.c_void_p * (world * nc))(*[self.peer_ptrs[p][c] for p in range(world) for c in range(nc)])
            ctx.check(ctx.lib.dfgpu_partition_plan_scatter_peer(plan, bases, dst_row))
            shared_stream = (ctx.lib.dfgpu_ctx_stream(ctx.h) or 0) == torch.cuda.current_stream().cuda_stream
            if not shared_stream:
                ctx.sync()
            self.dist.all_reduce(self._flag)                       # completion barrier: every rank's scatter has finished
            if not shared_stream:
                torch.cuda.current_stream().synchronize()
        finally:
            ctx.lib.dfgpu_partition_plan_destroy(plan)
        return ExchangedPeerBatch(self, recv_rows)


class ExchangedPeerBatch:
    def __init__(self, px: PeerExchange, rows: int):
        self.px, self.rows = px, rows

    def columns(self) -> List[D.Column]:
        cols = []
        for b, ty in zip(self.px.bufs, self.px.types):
            c = D.Column()
            c.type, c.flags, c.length, c.offset, c.null_count = ty, 0, self.rows, 0, 0
            c.values, c.validity = b.ptr, None
            cols.append(c)
        return cols
