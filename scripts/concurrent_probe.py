"""q6 step with T concurrent tasks over S row-group splits (spark.rapids.sql.concurrentGpuTasks analogue)."""
import sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
sys.path.insert(0, ".")
import spark_rapids_b200 as m
from benchdata import tpch
import bench
m.init(0, 12 << 30)
rows = 59_986_052
raw = tpch.lineitem_q6_parquet(rows, 42, "/tmp/b2_bench_cache")
m.host_register(raw)
prog, spec = bench.build_q6(m)
COLS = bench.COLS
dev = m.DeviceBuffer(raw.nbytes + 64); dev.copy_from_host(raw)
nrg = m.parquet_num_row_groups(raw)
expect = None
def run(T, S, n=8):
    bounds = [round(i * nrg / S) for i in range(S + 1)]
    splits = [(bounds[i], bounds[i + 1]) for i in range(S)]
    pool = ThreadPoolExecutor(T)
    def task(tid):
        out = []
        for (a, b) in splits[tid::T]:
            tb = m.parquet_decode_row_groups(raw, COLS, a, b, dev.ptr)
            out.append(m.scan_aggregate(prog, True, tb, [], spec))
        m.sync()
        return out
    def step():
        parts = [p for f in [pool.submit(task, t) for t in range(T)] for p in f.result()]
        return m.reduce(m.concat(parts), [(m.AGG_SUM, 0, m.DECIMAL128, 4, 35)]).to_rows()[0][0]
    for _ in range(3): r = step()
    m.sync(); t0 = time.perf_counter()
    for _ in range(n): r = step()
    m.sync(); dt = (time.perf_counter() - t0) / n * 1e3
    pool.shutdown()
    return dt, r
for T, S in [(1, 1), (1, 2), (2, 2), (2, 4), (2, 6), (3, 6), (4, 4), (4, 8), (2, 13)]:
    dt, r = run(T, S)
    if expect is None: expect = r
    print("tasks %d splits %2d: %.2f ms/step  %s" % (T, S, dt, "ok" if r == expect else "MISMATCH %s %s" % (r, expect)))
