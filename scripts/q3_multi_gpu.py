"""BASELINE config C4 on N GPUs of one box (weak scaling: every rank owns the rows [rank n, (rank+1) n) of each table of the
SF(100 N) database) — the reference's Q3 physical plan with its RepartitionExec(Hash) exchanges (tpch/plans/q3.slt.part:60-76)
re-cut for NVLink:

  customer  : FilterExec(c_mktsegment = 1) -> keys                      -> all-gather (24 MB / rank, the CollectLeft idea, exec.rs:1326-1336)
              -> every rank builds the GLOBAL key bitmap L1
  orders    : FilterExec(o_orderdate < CUT) -> RightSemi vs L1 -> (o_orderkey, o_orderdate, o_shippriority)   [one fused pipeline]
              -> its keys set the bits of a membership filter F (global geometry) -> OR-all-reduce of F over peer memory (NVLink)
              -> RepartitionExec Hash(o_orderkey): fused partition + peer-memory scatter -> owner builds L2 = {o_orderkey -> (date, prio)}
  lineitem  : FilterExec(l_shipdate > CUT) -> MAYBE vs F (the dynamic filter the downstream join pushes into this scan,
              joins/hash_join/shared_bounds.rs: 9 of 10 rows have no partner and never reach the exchange) -> (l_orderkey, price, discount)
              -> RepartitionExec Hash(l_orderkey): fused partition + peer-memory scatter (~5 % of the rows cross NVLink)
              -> owner: Inner vs L2 -> AggregateExec SinglePartitioned (group keys contain the partition key: no final exchange)

NCCL carries only counts, barriers and the 24 MB key all-gather; row payloads move by peer stores.  Launch: torchrun, one rank per GPU."""
import ctypes as C
import time

import numpy as np

from datafusion_b200 import capi as D
from datafusion_b200 import exchange
import q3_device_pipeline as Q

M64 = (1 << 64) - 1


class PartitionedQ3:
    def __init__(self, device, dist, sf, seed=1):
        import torch
        self.torch, self.dist, self.sf = torch, dist, sf
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.stream = torch.cuda.Stream(device)
        torch.cuda.set_stream(self.stream)            # ONE stream for libdfgpu and NCCL: collectives and kernels are ordered by it
        self.ctx = ctx = D.Context(device, self.stream.cuda_stream)
        self.dev = torch.device("cuda", device)
        self.cu, self.orr, self.li = Q.gen_tables(ctx, sf, seed, self.rank, self.world)
        self.input_rows = self.cu.rows + self.orr.rows + self.li.rows
        nc, no, nl = self.cu.rows, self.orr.rows, self.li.rows
        self.NC = nc * self.world
        # persistent exchange buffers (CUDA IPC mapped once): qualified orders and filtered lineitems, with head room
        self.px_o = exchange.PeerExchange(ctx, dist, [D.INT64, D.INT32, D.INT32], int(no * 0.14) + 4096)
        self.px_l = exchange.PeerExchange(ctx, dist, [D.INT64, D.INT64, D.INT64], int(nl * 0.09) + 4096)
        # the membership filter: identical geometry on every rank, sized for the GLOBAL number of qualifying orders
        self.F = D.Lookup(ctx, D.INT64, [], expected_rows=int(no * self.world * 0.11) + 1024, filter_only=True)
        ptr, nbytes = self.F.filter_buffer()
        h = C.create_string_buffer(64)
        ctx.check(ctx.lib.dfgpu_ipc_export(ctx.h, C.c_void_p(ptr), h))
        mine = torch.frombuffer(bytearray(h.raw), dtype=torch.uint8).to(self.dev)
        allh = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine)
        self.f_ptrs = []
        for r in range(self.world):
            if r == self.rank:
                self.f_ptrs.append(ptr)
            else:
                out = C.c_void_p()
                ctx.check(ctx.lib.dfgpu_ipc_import(ctx.h, bytes(allh[r].cpu().numpy().tobytes()), C.byref(out)))
                self.f_ptrs.append(out.value)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.last_res, self.last_stages = [], {}

    def extra_launches(self):
        return 0

    def exchange_description(self):
        return ("RepartitionExec(Hash) x2 as fused hash partition + peer-memory scatter over NVLink (CUDA IPC) — qualified orders (16 B/row) and "
                "membership-filtered lineitems (24 B/row, ~5 % of the scan); the join's membership filter is OR-all-reduced over peer memory; "
                "NCCL carries counts / barriers and the 24 MB customer-key all-gather")

    def _barrier(self):
        self.dist.all_reduce(self._flag)

    def step(self):
        ctx, cu, orr, li = self.ctx, self.cu, self.orr, self.li
        B, Cc, L = Q.B, Q.C, Q.L
        st = {}
        for b in self.last_res:
            b.release()
        self.last_res = []
        # ---- customer: filter -> keys -> all-gather -> global bitmap ----
        p = D.Pipeline(ctx, cu.types, B(D.OP_EQ, Cc(1), L(1)))
        p.sink_output([0], ordered=False)
        p.push_device(cu.cols); p.finish()
        ck = p.drain(host=False)
        st["customer_building"] = p.metric("sink_rows")
        p.close()
        kcols = [ck[0].column(0)] if ck else [_empty_col(ctx, D.INT64)]
        g = exchange.all_gather_columns(ctx, kcols, self.dist)
        l1 = D.Lookup(ctx, D.INT64, [], key_range=(1, self.NC))
        p = D.Pipeline(ctx, [D.INT64]); p.sink_build(l1, 0, [])
        p.push_device(g.columns()); p.finish(); p.close()
        for b in ck:
            b.release()
        # ---- orders: filter + semi probe -> qualified orders; their keys -> membership filter F ----
        p = D.Pipeline(ctx, orr.types, B(D.OP_LT, Cc(2), L(Q.CUT, D.INT32)), [(D.STAGE_SEMI, 1, l1)], name="orders")
        p.sink_output([0, 2, 3], ordered=False)
        p.push_device(orr.cols); p.finish()
        qo = p.drain(host=False)
        st["orders_of_building_customers_local"] = p.metric("sink_rows")
        p.close()
        qcols = [qo[0].column(i) for i in range(3)] if qo else [_empty_col(ctx, t) for t in (D.INT64, D.INT32, D.INT32)]
        self.F.clear()                                    # safe: the previous step's merge finished everywhere before its last barrier
        p = D.Pipeline(ctx, [D.INT64]); p.sink_build(self.F, 0, [])
        p.push_device([qcols[0]]); p.finish(); p.close()
        self._barrier()                                   # every rank's bits are set
        self.F.filter_allreduce_peer(self.f_ptrs, self.rank)
        self._barrier()                                   # every slice is merged everywhere
        xo = self.px_o.exchange(qcols, [0])               # RepartitionExec Hash(o_orderkey)
        for b in qo:
            b.release()
        l2 = D.Lookup(ctx, D.INT64, [D.INT32, D.INT32], n_acc_words=2, membership_filter=0, expected_rows=max(xo.rows, 1))
        p = D.Pipeline(ctx, [D.INT64, D.INT32, D.INT32]); p.sink_build(l2, 0, [1, 2])
        p.push_device(xo.columns()); p.finish(); p.close()
        st["orders_owned"] = xo.rows
        # ---- lineitem: filter + pushed-down membership filter -> exchange -> owner probes + aggregates ----
        p = D.Pipeline(ctx, li.types, B(D.OP_GT, Cc(3), L(Q.CUT, D.INT32)), [(D.STAGE_MAYBE, 0, self.F)], name="lineitem")
        p.sink_output([0, 1, 2], ordered=False)
        p.push_device(li.cols); p.finish()
        ql = p.drain(host=False)
        st["lineitems_past_filter_local"] = p.metric("sink_rows")
        p.close()
        lcols = [ql[0].column(i) for i in range(3)] if ql else [_empty_col(ctx, D.INT64) for _ in range(3)]
        xl = self.px_l.exchange(lcols, [0])               # RepartitionExec Hash(l_orderkey)
        for b in ql:
            b.release()
        p = D.Pipeline(ctx, [D.INT64, D.INT64, D.INT64], None, [(D.STAGE_INNER, 0, l2)], name="owner_probe_agg")
        p.sink_aggregate([0, 3, 4], [(D.AGG_SUM, B(D.OP_MULTIPLY, Cc(1), B(D.OP_MINUS, L(100), Cc(2))))], D.AGG_SINGLE_PARTITIONED)
        p.push_device(xl.columns()); p.finish()
        self.last_res = p.drain(host=False)
        st["joined_rows_owned"], st["groups_owned"] = p.metric("sink_rows"), p.metric("num_groups")
        p.close(); l2.close(); l1.close()
        self.last_stages = st
        return st

    def fingerprint(self):
        """this rank's share: [groups, sums of the four result columns] + [joined rows, qualified orders owned]"""
        fp = Q.result_fingerprint(self.ctx, self.last_res)
        return fp + [self.last_stages.get("joined_rows_owned", 0), self.last_stages.get("orders_owned", 0)]

    def all_reduce_fingerprint(self, fp_local):
        """wrapping sum over ranks (every group is owned by exactly one rank)"""
        return exchange.allgather_wrapping_sum(self.dist, fp_local, self.dev)

    def e2e(self, steps, barrier):
        """host leg: every rank uploads its shard from pinned host memory each step (H2D inside the timed region), runs the step and
        downloads its result rows; wall clock between barriers, max over ranks"""
        torch, ctx = self.torch, self.ctx
        err, secs, d2h = None, 0.0, 0
        try:
            if steps <= 0:
                raise RuntimeError("--e2e-steps 0: host leg skipped")
            host = []
            for t in (self.cu, self.orr, self.li):
                hs = []
                for c, ty in zip(t.cols, t.types):
                    h = ctx.pinned_empty(t.rows, D.NP_OF_TYPE[ty])
                    ctx.check(ctx.lib.dfgpu_memcpy_d2h(ctx.h, h.ctypes.data_as(C.c_void_p), C.c_void_p(c.values), t.rows * D.WIDTH[ty]))
                    hs.append(h)
                host.append(hs)
            ctx.sync()

            def one():
                for t, hs in zip((self.cu, self.orr, self.li), host):       # H2D of the shard into the resident column buffers
                    for c, ty, h in zip(t.cols, t.types, hs):
                        ctx.check(ctx.lib.dfgpu_memcpy_h2d(ctx.h, C.c_void_p(c.values), h.ctypes.data_as(C.c_void_p), t.rows * D.WIDTH[ty]))
                self.step()
                n = 0
                for b in self.last_res:                                      # D2H of this rank's result rows
                    for i in range(4):
                        cc = b.column(i)
                        ctx.to_host(cc.values, b.num_rows * D.WIDTH[cc.type])
                    n += b.num_rows * 24
                return n
            one()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                d2h = one()
            barrier()
            secs = time.perf_counter() - t0
        except Exception as exc:
            err = f"{type(exc).__name__}: {exc}"[:300]
        dt = torch.tensor([secs, 1.0 if err else 0.0, float(d2h)], device=self.dev, dtype=torch.float64)
        mx = dt.clone(); self.dist.all_reduce(mx, op=self.dist.ReduceOp.MAX)
        sm = dt.clone(); self.dist.all_reduce(sm, op=self.dist.ReduceOp.SUM)
        h2d = sum(t.rows * D.WIDTH[ty] for t in (self.cu, self.orr, self.li) for ty in t.types) * self.world
        if float(mx[1].item()) > 0 or float(mx[0].item()) <= 0:
            return {"value": None, "unit": "rows/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": None, "error": err or "failed on another rank"}
        s = float(mx[0].item())
        return {"value": self.input_rows * self.world * steps / s, "unit": "rows/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(float(sm[2].item())),
                "steps": steps, "ms_per_step": 1000 * s / steps,
                "timer": "host wall clock between barriers, max over ranks; per rank: H2D of its SF shard from pinned memory -> the partitioned pipeline -> D2H of its result rows"}


def _empty_col(ctx, ty):
    c = D.Column()
    buf = D.DeviceBuffer(ctx, 64)
    c.type, c.flags, c.length, c.offset, c.null_count, c.values, c.validity = ty, 0, 0, 0, 0, buf.ptr, None
    c._keep = buf
    return c
