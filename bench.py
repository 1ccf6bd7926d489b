#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (contract in the task statement).

Workload at every N: BASELINE.json configs[1], "TPC-H SF10 q6 (scan+filter+agg) single B200, Parquet
source", per GPU (weak scaling: every rank owns its own SF10 lineitem partition, seed 42+rank):
    Parquet bytes -> device decode (4 columns) -> fused filter + project + DECIMAL128 sum
    -> (N>1) NCCL exchange of the partial aggregates to the final-aggregate owner -> merge.
A "step" is one pass of that pipeline over one rank's whole partition.

value   : lineitem rows/s, whole job (all ranks), Parquet bytes already resident in HBM.
e2e     : same metric through the reference-facing call with HOST (pinned) buffers: the H2D copy of
          the Parquet bytes and the D2H of the result are inside the timed region.
roofline: the dominant kernel of the step (largest share of device time), achieved = algorithmic
          bytes / its CUDA-event time, peak = MEASURED_PEAKS.json hbm_gbs (else 6650 fallback).
--impl reference: the CPU restatement of the same plan (oracle/tpch.py q6_cpu: pyarrow multithreaded
          scan + compute on all host cores; no JVM/Spark exists in this image) on the same input.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SF10_ROWS = 59_986_052
COLS = ["l_shipdate", "l_discount", "l_quantity", "l_extendedprice"]
CACHE = os.environ.get("B2_BENCH_CACHE", "/tmp/b2_bench_cache")


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.proc = gpu, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_q6(m):
    c_ship = m.col(0, m.DATE32, nullable=False)
    c_disc, c_qty, c_price = (m.col(i, m.DECIMAL64, 12, 2, nullable=False) for i in (1, 2, 3))
    from benchdata import tpch
    pred = ((c_ship >= m.lit(tpch.Q6_DATE_LO, m.DATE32)) & (c_ship < m.lit(tpch.Q6_DATE_HI, m.DATE32)) & (c_disc >= m.lit(5, m.DECIMAL64, 3, 2))
            & (c_disc <= m.lit(7, m.DECIMAL64, 3, 2)) & (c_qty < m.lit(2400, m.DECIMAL64, 12, 2)))
    rev = c_price * c_disc
    return m.Program([pred, rev]), [(m.AGG_SUM, 0, m.DECIMAL128, 4, 35)]


def run_reference(args, rank, world):
    """CPU arm: rank 0 only."""
    if rank != 0:
        return
    from benchdata import tpch as gen
    from oracle import tpch          # the CPU arm IS the restatement (no JVM/Spark on the box)
    rows = args.rows
    raw = gen.lineitem_q6_parquet(rows, 42, CACHE)
    cores = os.cpu_count() or 1
    # bounded sample per step: the whole partition if it is small enough, else its first row groups
    sample_rows = rows
    for _ in range(args.warmup):
        res = tpch.q6_cpu(raw, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = tpch.q6_cpu(raw, cores)
    dt = time.perf_counter() - t0
    val = sample_rows * args.steps / dt
    line = {"impl": "reference", "metric": "tpch_q6_rows_per_sec", "value": val, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64/decimal128", "data": "synthetic",
            "config": {"workload": "TPC-H SF10 q6 (scan+filter+agg), Parquet source, CPU plan", "rows": rows, "result": res},
            "cpu_baseline": {"value": val, "unit": "rows/s", "cores": cores, "kind": "port",
                             "sample": "full %d-row partition per step; pyarrow %d threads (CPU restatement, NOT Spark)" % (sample_rows, cores)},
            "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--rows", type=int, default=SF10_ROWS, help="lineitem rows per GPU (default: SF10)")
    ap.add_argument("--cpu-baseline", type=int, default=1)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    # stdout carries exactly one JSON line: anything a library prints (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch  # plumbing only: rendezvous, barrier, max-over-ranks
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import spark_rapids_b200 as m
    from benchdata import tpch       # synthetic inputs only; oracle/ is touched by the cpu_baseline leg alone
    m.init(local, 8 << 30)   # Rmm.initialize analogue: pre-grown stream-ordered pool

    rows = args.rows
    raw = tpch.lineitem_q6_parquet(rows, 42 + rank, CACHE)
    nbytes = raw.nbytes
    m.host_register(raw)                                  # pinned host staging (HostAlloc pinned pool)
    dev = m.DeviceBuffer(nbytes + 64)
    dev.copy_from_host(raw)
    prog, spec = build_q6(m)
    comm = None
    if world > 1:
        uid = [m.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = m.Comm(uid[0], rank, world)

    def step(resident, dev_ptr=None):
        if dev_ptr is not None:
            t = m.parquet_decode_device(raw, dev_ptr, COLS)
        else:
            t = m.parquet_decode_device(raw, dev.ptr, COLS) if resident else m.parquet_decode(raw, COLS)
        part = m.scan_aggregate(prog, True, t, [], spec)           # partial aggregate (1 row)
        if comm is not None:                                        # exchange: SinglePartition -> rank 0 owns the final aggregate
            got = comm.exchange(part, [0] + [1] * world)
            part = m.reduce(got, [(m.AGG_SUM, 0, m.DECIMAL128, 4, 35)])
        return part.to_rows()[0][0]                                 # D2H of the result

    def barrier():
        m.sync(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        m.sync()

    def timed(resident, steps, profile=False):
        barrier()
        if profile:
            m.profile_enable(True)
        l0 = m.kernel_launch_count()
        e0, e1 = m.Event(), m.Event()
        w0 = time.perf_counter()
        e0.record()
        if resident:
            for _ in range(steps):
                res = step(True)
        else:
            # e2e: every step copies its Parquet bytes from pinned host memory; the copy of step k+1 runs on the
            # copy stream while step k decodes (the reference's multithreaded reader keeps host buffers in flight
            # the same way, GpuMultiFileReader.scala)
            nxt = m.AsyncUpload(raw)
            for k in range(steps):
                cur, nxt = nxt, (m.AsyncUpload(raw) if k + 1 < steps else None)
                res = step(False, cur.wait())
                cur.free()
        e1.record()
        m.sync()
        ms = e0.elapsed_ms(e1)
        wall = (time.perf_counter() - w0) * 1000
        launches = m.kernel_launch_count() - l0
        prof = m.profile_report() if profile else None
        if profile:
            m.profile_enable(False)
        barrier()
        if world > 1:
            tt = torch.tensor([ms], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
        return ms, wall, launches, prof, res

    for _ in range(args.warmup):
        res_w = step(True)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, wall, launches, prof, res = timed(True, args.steps, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step(False)
    ms_e2e, wall_e2e, _, _, res_e2e = timed(False, args.steps)

    if rank != 0:
        if comm:
            comm.close()
        return
    total_rows = rows * world
    value = total_rows * args.steps / (ms / 1000)
    e2e = total_rows * args.steps / (ms_e2e / 1000)
    peak, peak_src = hbm_peak()
    # per-kernel shares and the dominant kernel's roofline.  Algorithmic bytes per launch:
    #   snappy_kernel: compressed bytes read + uncompressed bytes written
    #   values_kernel: uncompressed page bytes read + 28 B/row columns written
    #   aggregate_smem_kernel (fused filter+project+sum): 28 B/row read
    kern = {k["name"]: k for k in prof}
    tot_k = sum(k["ms"] for k in prof) or 1.0
    st = m.parquet_last_stats()  # byte accounting of the decoder for one step (identical every step)
    alg = {"snappy_kernel": st["compressed_in"] + st["decompressed_out"], "values_kernel": st["page_bytes"] + st["column_bytes"],
           "aggregate_smem_kernel": rows * 28.0}
    kernels = []
    for k in prof:
        per = k["ms"] / max(1, k["launches"])
        ent = {"name": k["name"], "launches_per_step": k["launches"] / args.steps, "ms_per_launch": per, "share": k["ms"] / tot_k}
        if k["name"] in alg:
            ent["alg_GBps"] = alg[k["name"]] / 1e9 / (per / 1000)
        kernels.append(ent)
    kernels.sort(key=lambda e: -e["share"])
    dom = kernels[0] if kernels else {"name": None, "alg_GBps": 0.0}
    # DRAM traffic of the dominant kernel: one `ncu --set full` capture of this very workload, committed under profiles/
    traffic = None
    try:
        if rows == SF10_ROWS:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_sf10_traffic.json")))[dom["name"]]["traffic_bytes_per_launch"]
    except Exception:
        traffic = None
    roof = {"bound": "hbm", "kernel": dom["name"], "achieved": dom.get("alg_GBps", 0.0), "peak": peak, "unit": "GB/s",
            "frac": dom.get("alg_GBps", 0.0) / peak, "traffic": traffic, "algorithmic_bytes": alg.get(dom["name"]),
            "traffic_source": "profiles/r1_ncu_sf10_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum, one launch)" if traffic else None,
            "peak_source": peak_src,
            "note": "dominant kernel by CUDA-event share of the step; per-kernel list in `kernels`"}
    line = {"metric": "tpch_q6_rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "query_sec": ms / args.steps / 1000, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64/decimal128", "data": "synthetic",
            "config": {"workload": "TPC-H SF10 q6 (scan+filter+agg), Parquet source (snappy, dictionary, INT64 decimals)", "rows_per_gpu": rows,
                       "parquet_bytes_per_gpu": int(nbytes), "l2": "inputs larger than L2 (parquet %.0f MB + 1.68 GB decoded per step)" % (nbytes / 1e6),
                       "result_unscaled_dec25_4": res, "parallelism": "partition per GPU, NCCL exchange of partial aggregates" if world > 1 else "1 GPU"},
            "e2e": {"value": e2e, "unit": "rows/s", "h2d_bytes_per_step": int(nbytes) * world, "d2h_bytes_per_step": 16 * world,
                    "ms_per_step": ms_e2e / args.steps},
            "parquet_stats": st, "gpu_launches": int(launches), "wall_ms_per_step": wall / args.steps, "clocks": clocks, "roofline": roof, "kernels": kernels[:8]}
    assert res_w == res and (world > 1 or res_e2e == res), ("q6 result differs between steps", res_w, res, res_e2e)
    if args.cpu_baseline and world >= 1:
        from oracle import tpch as cpu   # checker + reported baseline: the only use of oracle/ in this arm
        cores = os.cpu_count() or 1
        cpu.q6_cpu(raw, cores)
        t0 = time.perf_counter(); reps = 2
        for _ in range(reps):
            cres = cpu.q6_cpu(raw, cores)
        dt = (time.perf_counter() - t0) / reps
        assert cres == res or world > 1, ("CPU restatement disagrees with the GPU result", cres, res)
        if world == 1 and rows <= SF10_ROWS:   # exact integer restatement over the raw columns pins both
            expect = cpu.q6_numpy_chunks(tpch.lineitem_q6_chunks(rows, 42 + rank))
            assert res == expect, ("q6 result mismatch", res, expect)
        line["cpu_baseline"] = {"value": rows / dt, "unit": "rows/s", "cores": cores, "kind": "port",
                                "sample": "one full %d-row partition, %d reps; pyarrow scan+compute on %d threads (CPU restatement, NOT Spark)" % (rows, reps, cores)}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if comm:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
