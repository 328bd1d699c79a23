#!/usr/bin/env python
"""bench.py — BASELINE.json metric on its config: rows/sec of the join hot path (HashJoinExec, config C2).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

One step = one complete inner hash join (build + probe + output materialisation) of
  probe 100M rows {k:int64, pp:int64}  JOIN  build 10M rows {k:int64, pb:int64}   -> {k, pb, pp}
with SPARSE UNIQUE build keys k = splitmix64(42, i) and probe keys drawn from them (100 % hit rate,
fan-out 1, 100M output rows): SURVEY.md §8d C2(ii) — the case where the reference takes its hashbrown
path (the dense-key / ArrayMap case is reported under "extra").  rows/sec = (build + probe rows) / time.

* value  : device-resident inputs (generated in HBM by the shared counter-based generators), timed with
           CUDA events on the launching stream; inputs (1.76 GB) exceed L2 (126 MB) so no L2 flush is needed.
* e2e    : the same join through the C ABI with HOST buffers (pinned): H2D of both inputs and D2H of the
           100M-row result inside the timed region.
* roofline: the dominant kernel (join_probe: fused probe + materialise), algorithmic bytes = 40 B per
           probe row (16 B read + 24 B written), duration from CUDA events around that kernel.
* cpu_baseline / --impl reference: oracle/ C restatement of the reference's partitioned HashJoinExec
           (RepartitionExec + per-partition build/probe/take), all host threads — kind "port": the
           Rust reference cannot be built in this image.
N > 1 (torchrun): weak scaling — every rank owns 100M probe + 10M build rows, rows are hash-partitioned
on the GPU, exchanged with ONE all-to-all per column (NCCL via torch.distributed), then joined locally.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NB_PER_GPU = 10_000_000
NP_PER_GPU = 100_000_000
METRIC = "rows/sec join (HashJoinExec build+probe+emit, TPC-H-shaped int64 keys)"
UNIT = "rows/s"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""

    def __init__(self, device=0):
        self.device, self.samples, self.proc = device, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args, rank, world):
    """the reference arm: CPU restatement (oracle/, kind "port") on all host cores, bounded sample per step"""
    if rank != 0:
        return
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    nb, npr = NB_PER_GPU, NP_PER_GPU
    # bounded sample: a 1/2-size instance of the same workload per step keeps K steps within minutes on small hosts
    frac = 1.0 if threads >= 32 else 0.5
    nb, npr = int(nb * frac), int(npr * frac)
    bk = O.generate_i64(2, 42, 0, 0, nb, threads); bp = O.generate_i64(2, 7, 0, 0, nb, threads)
    pk = O.generate_i64(4, 42, 43, nb, npr, threads); pp = O.generate_i64(2, 8, 0, 0, npr, threads)
    for _ in range(max(1, min(args.warmup, 1))):
        O.bench_join(bk, bp, pk, pp, threads=threads)
    t = 0.0
    for _ in range(args.steps):
        secs, rows, _ = O.bench_join(bk, bp, pk, pp, threads=threads)
        assert rows == npr
        t += secs
    value = (nb + npr) * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic (counter-based generators, identical to the GPU arm)",
            "config": {"workload": "C2 HashJoinExec inner 100M x 10M int64, sparse unique keys, 100% hit (sample below)", "batch_size": 8192,
                       "partition_mode": "Partitioned", "target_partitions": threads},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{npr} probe x {nb} build rows per step (fraction {frac} of C2), oracle/oracle.c oracle_bench_join: RepartitionExec(Hash) both sides + per-partition JoinHashMap build/probe/take, batch_size 8192; inputs and repartition buffers resident (pre-faulted) before the clock starts"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def make_inputs(ctx, D, rank, world, nb, npr):
    """rank r owns global rows [r*n, (r+1)*n) of both tables (the generators are counter-based)"""
    nb_all = nb * world
    bk = ctx.generate_i64(D.GEN_SPLITMIX, 42, 0, 0, rank * nb, nb)
    bp = ctx.generate_i64(D.GEN_SPLITMIX, 7, 0, 0, rank * nb, nb)
    pk = ctx.generate_i64(D.GEN_SPARSE_OF, 42, 43, nb_all, rank * npr, npr)
    pp = ctx.generate_i64(D.GEN_SPLITMIX, 8, 0, 0, rank * npr, npr)
    return bk, bp, pk, pp


def join_step(ctx, D, build_cols, probe_cols, keep_output=False):
    j = D.HashJoinHandle(ctx, [D.INT64, D.INT64], [D.INT64, D.INT64], [0], [0], [0, 0, 1], [0, 1, 1])
    j.push_build_device(build_cols)
    j.finish_build()
    j.push_probe_device(probe_cols)
    j.finish_probe()
    rows = j.metric("output_rows")
    outs = j.drain(host=False)
    if not keep_output:
        for b in outs:
            b.release()
        outs = []
    j.close()
    return rows, outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if world == 1:
        args.e2e_steps = max(args.e2e_steps, 1)     # the single-GPU line always carries the host leg

    from datafusion_b200 import capi as D
    dist = None
    xmode = os.environ.get("DFGPU_EXCHANGE", "pipelined") if world > 1 else "none"
    nb, npr = NB_PER_GPU, NP_PER_GPU
    pj = px_b = px_p = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        from datafusion_b200 import exchange
        dist = dist_
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if xmode == "pipelined":
            # exchange stream + join stream; persistent receive buffers mapped into every peer through CUDA IPC (25 % headroom)
            pj = exchange.PartitionedHashJoin(local, dist, [D.INT64, D.INT64], [D.INT64, D.INT64], [0], [0], [0, 0, 1], [0, 1, 1],
                                              int(nb * 1.25), int(npr * 1.25), n_chunks=int(os.environ.get("DFGPU_CHUNKS", "1")))
            ctx = pj.ctx
            torch.cuda.set_stream(pj.js)
        else:
            tstream = torch.cuda.Stream()           # one explicit stream shared by torch (NCCL ordering) and libdfgpu
            torch.cuda.set_stream(tstream)
            ctx = D.Context(local, tstream.cuda_stream)
            if xmode == "peer":
                px_b = exchange.PeerExchange(ctx, dist, [D.INT64, D.INT64], int(nb * 1.25))
                px_p = exchange.PeerExchange(ctx, dist, [D.INT64, D.INT64], int(npr * 1.25))
    else:
        ctx = D.Context(local)
    bk, bp, pk, pp = make_inputs(ctx, D, rank, world, nb, npr)
    col = lambda buf, n: D.DeviceColumn(ctx, D.INT64, n, buf)
    build_cols, probe_cols = [col(bk, nb), col(bp, nb)], [col(pk, npr), col(pp, npr)]

    def step():
        if world == 1:
            return join_step(ctx, D, build_cols, probe_cols)[0]
        if pj is not None:     # chunked peer scatter on the exchange stream overlapped with build/probe on the join stream
            return pj.run(build_cols, probe_cols, keep_output=False)[0]
        if px_b is not None:   # fused partition + exchange, then the join (no overlap)
            b2 = px_b.exchange(build_cols, [0])
            p2 = px_p.exchange(probe_cols, [0])
        else:                  # local partition + one NCCL all-to-all per column
            b2 = exchange.exchange_batch(ctx, build_cols, [0], dist)
            p2 = exchange.exchange_batch(ctx, probe_cols, [0], dist)
        return join_step(ctx, D, b2.columns(), p2.columns())[0]

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        out_rows = step()
    barrier()
    ctx.set_kernel_timing(True)
    ctx.kernel_time_reset()
    all_launches = lambda: ctx.launches + (pj.ctx_x.launches if pj is not None else 0)
    launches0 = all_launches()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(args.steps):
        out_rows = step()
    ctx.record(e1)
    ms = ctx.elapsed_ms(e0, e1)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    launches = all_launches() - launches0
    ctx.set_kernel_timing(False)
    if dist is not None:
        import torch
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    probe_ms, probe_n = ctx.kernel_time("join_probe")
    build_ms, build_n = ctx.kernel_time("join_build")
    ms_per_step = ms / args.steps
    value = (nb + npr) * world / (ms_per_step / 1000.0)

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        launches_per_step = max(probe_n, 1) / args.steps          # 1 at N=1; one per exchanged chunk in the pipelined N>1 path
        algo_bytes = 40.0 * float(out_rows) / launches_per_step   # 16 B read + 24 B written per probe row
        k_ms = probe_ms / max(probe_n, 1)
        achieved = algo_bytes / (k_ms / 1000.0) / 1e9 if k_ms > 0 else 0.0
        # DRAM bytes per launch of the probe kernel from the ncu --set full capture of this workload (profiles/r1h_final_kernels.csv:
        # dram__bytes_read.sum 12.94 GB + dram__bytes_write.sum 2.38 GB); only meaningful for the one-launch-per-step N=1 shape
        traffic = 15.32e9 if world == 1 else None
        roofline = {"bound": "hbm", "kernel": "join_probe_inline_v0_kernel<2>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r1h_final_kernels.csv)",
                    "traffic_note": "3.8x the algorithmic 4.0 GB: every probe row is one random 16 B table lookup that costs a 64 B DRAM fetch (+ L2 pair-sector fill); the kernel sits on the measured DRAM random-access rate, not on bytes (DESIGN.md §4)",
                    "peak_source": peak_src, "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
                    "algorithmic_bytes_per_launch": algo_bytes, "build_kernel_ms": build_ms / max(build_n, 1),
                    "whole_join_achieved_gbs": (16.0 * nb + 16.0 * npr + 24.0 * npr) / (ms_per_step / 1000.0) / 1e9 if world == 1 else None}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic (generated in HBM, counter-based)",
                "config": {"workload": "C2 HashJoinExec inner 100M x 10M int64 per GPU, sparse unique build keys, 100% hit, output {k,pb,pp}",
                           "rows_per_step": (nb + npr) * world, "output_rows_per_gpu": int(out_rows), "l2": "inputs (1.76 GB/GPU) exceed L2; no flush",
                           "exchange": {"none": "none (single GPU)", "pipelined": "fused hash partition + peer-memory scatter over NVLink (CUDA IPC) in chunks on an exchange stream, overlapped with build/probe on the join stream; NCCL only for counts/barriers",
                                        "peer": "fused hash partition + direct peer-memory scatter over NVLink (CUDA IPC), NCCL only for counts/barrier",
                                        "nccl": "hash partition + one NCCL all-to-all per column"}.get(xmode, xmode)},
                "clocks": clk, "gpu_launches": int(launches), "roofline": roofline}

    # ---- e2e through the C ABI with host (pinned) buffers (N = 1: the host leg has no exchange) ----
    h2d = 16 * (nb + npr)
    import ctypes as C
    if world == 1:
        hb = [ctx.pinned_empty(nb, np.int64), ctx.pinned_empty(nb, np.int64)]
        hp = [ctx.pinned_empty(npr, np.int64), ctx.pinned_empty(npr, np.int64)]
        for dst, src, n in ((hb[0], bk, nb), (hb[1], bp, nb), (hp[0], pk, npr), (hp[1], pp, npr)):
            ctx.check(ctx.lib.dfgpu_memcpy_d2h(ctx.h, dst.ctypes.data_as(C.c_void_p), C.c_void_p(src.ptr), n * 8))
        ctx.sync()

    def e2e_step():
        j = D.HashJoinHandle(ctx, [D.INT64, D.INT64], [D.INT64, D.INT64], [0], [0], [0, 0, 1], [0, 1, 1])
        j.push_build_host([D.HostColumn(hb[0]), D.HostColumn(hb[1])])
        j.finish_build()
        j.push_probe_host([D.HostColumn(hp[0]), D.HostColumn(hp[1])])
        j.finish_probe()
        outs = j.drain(host=True)
        rows = sum(o.num_rows for o in outs)
        d2h = rows * 24
        chk = int(np.ctypeslib.as_array((C.c_int64 * 1).from_address(outs[0].column(0).values))[0]) if outs else 0
        for o in outs:
            o.release()
        j.close()
        return rows, d2h, chk

    if world == 1:
        for _ in range(2):
            e2e_step()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            rows, d2h, _ = e2e_step()
        ctx.sync()
        t1 = time.perf_counter()
        e2e_val = (nb + npr) * args.e2e_steps / (t1 - t0)
        line["e2e"] = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
                       "ms_per_step": 1000 * (t1 - t0) / args.e2e_steps, "timer": "host wall clock around the C-ABI calls (includes H2D, kernels, D2H)"}
    elif pj is not None:
        # N > 1: every rank uploads its shard from pinned host memory, runs the exchange + join, and downloads its output rows
        # into pinned host memory; wall clock between barriers, max over ranks
        import torch
        out_cap = int(npr * 1.25)
        setup_err = None
        try:   # 4.3 GB of pinned host memory per rank: agree that every rank got it before any collective step starts
            hb = [ctx.pinned_empty(nb, np.int64), ctx.pinned_empty(nb, np.int64)]
            hp = [ctx.pinned_empty(npr, np.int64), ctx.pinned_empty(npr, np.int64)]
            for dst, src, n in ((hb[0], bk, nb), (hb[1], bp, nb), (hp[0], pk, npr), (hp[1], pp, npr)):
                ctx.check(ctx.lib.dfgpu_memcpy_d2h(ctx.h, dst.ctypes.data_as(C.c_void_p), C.c_void_p(src.ptr), n * 8))
            ctx.sync()
            hout = [ctx.pinned_empty(out_cap, np.int64) for _ in range(3)]
        except Exception as exc:
            setup_err = f"{type(exc).__name__}: {exc}"[:300]
        okf = torch.tensor([0.0 if setup_err else 1.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        setup_ok = float(okf.item()) > 0

        s_in, s_out = torch.cuda.Stream(local), torch.cuda.Stream(local)
        ctx_in, ctx_out = D.Context(local, s_in.cuda_stream), D.Context(local, s_out.cuda_stream)
        K = int(os.environ.get("DFGPU_E2E_CHUNKS", "8"))
        bounds = [npr * k // K for k in range(K + 1)]

        def view(buf, lo, hi):
            c = D.Column()
            c.type, c.flags, c.length, c.offset, c.null_count, c.values, c.validity = D.INT64, 0, hi - lo, 0, 0, buf.ptr + lo * 8, None
            return c

        def e2e_step_multi():
            # all uploads are queued up front on the copy-in stream (one event per piece); the probe batches then stream
            # through exchange + probe on the exchange/join streams while later pieces are still uploading and earlier
            # results are downloading on the copy-out stream (PCIe is full duplex)
            for dst, src, n in ((bk, hb[0], nb), (bp, hb[1], nb)):
                ctx_in.check(ctx_in.lib.dfgpu_memcpy_h2d(ctx_in.h, C.c_void_p(dst.ptr), src.ctypes.data_as(C.c_void_p), n * 8))
            ev_b = torch.cuda.Event(); ev_b.record(s_in)
            evs = []
            for k in range(K):
                lo, hi = bounds[k], bounds[k + 1]
                for dst, src in ((pk, hp[0]), (pp, hp[1])):
                    ctx_in.check(ctx_in.lib.dfgpu_memcpy_h2d(ctx_in.h, C.c_void_p(dst.ptr + lo * 8), C.c_void_p(src.ctypes.data + lo * 8), (hi - lo) * 8))
                e = torch.cuda.Event(); e.record(s_in); evs.append(e)
            ev_b.synchronize()
            pj.build(build_cols)
            off, pending = 0, []
            for k in range(K):
                evs[k].synchronize()
                outs = pj.probe([view(pk, bounds[k], bounds[k + 1]), view(pp, bounds[k], bounds[k + 1])], n_chunks=1)
                for o in outs:
                    if off + o.num_rows > out_cap:
                        raise RuntimeError("e2e: output exceeds the pinned result buffers")
                    for c in range(3):
                        ctx_out.check(ctx_out.lib.dfgpu_memcpy_d2h(ctx_out.h, C.c_void_p(hout[c].ctypes.data + off * 8), C.c_void_p(o.column(c).values), o.num_rows * 8))
                    off += o.num_rows
                pending += outs
            rows, tail = pj.finish()
            s_out.synchronize()
            for o in pending + tail:
                o.release()
            return rows

        e2e_err = None if setup_ok else (setup_err or "pinned host allocation failed on another rank")
        rows = 0
        try:
            if not setup_ok:
                raise RuntimeError(e2e_err)
            if args.e2e_steps <= 0:
                raise RuntimeError("--e2e-steps 0: host leg skipped")
            e2e_step_multi()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                rows = e2e_step_multi()
            barrier()
            secs_local = time.perf_counter() - t0
        except Exception as exc:   # the device-resident line above stays valid; say why the host leg is missing
            e2e_err, secs_local, rows = f"{type(exc).__name__}: {exc}"[:300], 0.0, 0
        dt = torch.tensor([secs_local, 1.0 if e2e_err else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rank == 0:
            secs = float(dt[0].item())
            if float(dt[1].item()) > 0 or secs <= 0:
                line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": None, "error": e2e_err or "failed on another rank"}
            else:
                line["e2e"] = {"value": (nb + npr) * world * args.e2e_steps / secs, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": int(rows) * 24 * world,
                               "steps": args.e2e_steps, "ms_per_step": 1000 * secs / args.e2e_steps,
                               "timer": "host wall clock between barriers, max over ranks; per rank: H2D of its shard from pinned memory in pieces -> exchange + probe per piece -> D2H of the output rows, uploads / compute / downloads overlapped on three streams"}
    elif rank == 0:
        line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": None, "note": "host leg is measured in the default (pipelined) exchange mode only"}

    if rank == 0 and world == 1 and not args.no_extra:
        # ---- other BASELINE configs, device resident (context for the headline; not part of `value`) ----
        extra = {}
        dbk = ctx.generate_i64(D.GEN_PERM, 42, 0, nb, 0, nb); dpk = ctx.generate_i64(D.GEN_UNIFORM, 43, 0, nb, 0, npr)
        dcols_b, dcols_p = [col(dbk, nb), col(bp, nb)], [col(dpk, npr), col(pp, npr)]
        for _ in range(2):
            join_step(ctx, D, dcols_b, dcols_p)
        e0, e1 = ctx.event(), ctx.event()
        ctx.record(e0)
        for _ in range(3):
            join_step(ctx, D, dcols_b, dcols_p)
        ctx.record(e1)
        ms_d = ctx.elapsed_ms(e0, e1) / 3
        extra["C2_dense_keys_join"] = {"ms_per_step": ms_d, "rows_per_s": (nb + npr) / ms_d * 1e3,
                                       "note": "build k = perm(0..10M): the reference's ArrayMap rule applies (direct addressing)"}
        dbk.free(); dpk.free()
        ng, gn = 1_000_000, 1_000_000_000
        gk = ctx.generate_i64(D.GEN_UNIFORM, 5, 0, ng, 0, gn); gv = ctx.generate_i64(D.GEN_UNIFORM, 6, -2**31, 2**32, 0, gn)

        def agg_step():
            a = D.AggHandle(ctx, [D.INT64, D.INT64], [0], [(D.AGG_SUM, 1, -1), (D.AGG_COUNT, 1, -1)], capacity_hint=ng)
            a.push_device([col(gk, gn), col(gv, gn)])
            a.finish()
            g = a.metric("num_groups")
            for b in a.drain(host=False):
                b.release()
            a.close()
            return g
        agg_step()
        ctx.record(e0)
        for _ in range(3):
            groups = agg_step()
        ctx.record(e1)
        ms_g = ctx.elapsed_ms(e0, e1) / 3
        extra["C3_groupby_sum_count_1B_rows_1M_groups"] = {"ms_per_step": ms_g, "rows_per_s": gn / ms_g * 1e3, "groups": int(groups),
                                                          "achieved_gbs": (16.0 * gn + 24.0 * ng) / ms_g / 1e6, "frac_of_hbm_peak": (16.0 * gn + 24.0 * ng) / ms_g / 1e6 / peak,
                                                          "note": "bound by L2 atomics (2 RED + 1 tag read per row; measured RED peak 197 Gop/s), not HBM"}
        gk.free(); gv.free()
        # FilterExec: x:int64 > c over 100M rows x 2 columns, selectivity 20 % (C1's predicate at a size that is not launch-bound)
        fn = 100_000_000
        fx = ctx.generate_i64(D.GEN_UNIFORM, 1, 0, 1 << 32, 0, fn); fy = ctx.generate_i64(D.GEN_SPLITMIX, 2, 0, 0, 0, fn)
        nodes = [(D.EXPR_COLUMN, 0, 0, 0, 0, 0.0), (D.EXPR_LITERAL, 0, D.INT64, 0, int((1 << 32) * 0.8), 0.0), (D.EXPR_BINARY, D.OP_GT, 0, 0, 0, 0.0)]

        def filter_step():
            f = D.FilterHandle(ctx, [D.INT64, D.INT64], nodes, batch_size=0)
            f.push_device([col(fx, fn), col(fy, fn)])
            f.finish()
            outs = f.drain(host=False)
            kept = sum(o.num_rows for o in outs)
            for o in outs:
                o.release()
            f.close()
            return kept
        filter_step()
        ctx.record(e0)
        for _ in range(3):
            kept = filter_step()
        ctx.record(e1)
        ms_f = ctx.elapsed_ms(e0, e1) / 3
        fbytes = 16.0 * fn + 16.0 * kept   # both columns read once, kept rows of both columns written
        extra["C1_shape_filter_100M_rows_sel20"] = {"ms_per_step": ms_f, "rows_per_s": fn / ms_f * 1e3, "kept": int(kept),
                                                   "achieved_gbs": fbytes / ms_f / 1e6, "frac_of_hbm_peak": fbytes / ms_f / 1e6 / peak}
        fx.free(); fy.free()
        # C4: TPC-H Q3-shaped pipeline at SF100, device resident, operator by operator through the C ABI (scripts/q3_device_pipeline.py)
        try:
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts"))
            import q3_device_pipeline as Q
            sf = float(os.environ.get("DFGPU_Q3_SF", "100"))
            cu, orr, li = Q.gen_tables(ctx, sf)
            for _ in range(2):
                res, st = Q.run_q3(ctx, cu, orr, li)
                for b in res:
                    b.release()
            e0, e1 = ctx.event(), ctx.event()
            ctx.record(e0)
            for _ in range(3):
                res, st = Q.run_q3(ctx, cu, orr, li)
                for b in res:
                    b.release()
            ctx.record(e1)
            ms_q = ctx.elapsed_ms(e0, e1) / 3
            in_rows = cu.rows + orr.rows + li.rows
            q_bytes = 9.0 * cu.rows + 24.0 * orr.rows + 28.0 * li.rows            # every input column once
            extra["C4_tpch_q3_pipeline_device_resident"] = {"scale_factor": sf, "ms_per_step": ms_q, "rows_per_s": in_rows / ms_q * 1e3, "input_rows": in_rows, "stages": st,
                                                            "achieved_gbs": q_bytes / ms_q / 1e6, "frac_of_hbm_peak": q_bytes / ms_q / 1e6 / peak,
                                                            "note": "filter x3 -> RightSemi join -> Inner join -> projection -> 3-key group-by SUM; int64 fixed-point money; synthetic TPC-H-shaped tables generated in HBM"}
        except Exception as exc:
            extra["C4_tpch_q3_pipeline_device_resident"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        line["extra"] = extra

    if rank == 0:
        # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload ----
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            threads = os.cpu_count() or 1
            sb, sp = nb // 2, npr // 2
            hbk = O.generate_i64(2, 42, 0, 0, sb, threads); hbp = O.generate_i64(2, 7, 0, 0, sb, threads)
            hpk = O.generate_i64(4, 42, 43, sb, sp, threads); hpp = O.generate_i64(2, 8, 0, 0, sp, threads)
            O.bench_join(hbk, hbp, hpk, hpp, threads=threads)
            secs, rows, _ = O.bench_join(hbk, hbp, hpk, hpp, threads=threads)
            line["cpu_baseline"] = {"value": (sb + sp) / secs, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{sp} probe x {sb} build rows (half of C2), one timed run after one warm-up; oracle_bench_join partitioned hash join on all host threads, buffers pre-faulted"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
