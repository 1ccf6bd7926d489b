import sys, time
import numpy as np
sys.path.insert(0, ".")
import spark_rapids_b200 as m
from benchdata import tpch
import bench
m.init(0, 8 << 30)
rows = 59_986_052
raw = tpch.lineitem_q6_parquet(rows, 42, "/tmp/b2_bench_cache")
m.host_register(raw)
prog, spec = bench.build_q6(m)
COLS = bench.COLS
def t(fn, n=5):
    fn(); m.sync()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    m.sync()
    return (time.perf_counter() - t0) / n * 1e3
def up_only():
    u = m.AsyncUpload(raw); u.wait(); m.sync(); u.free()
print("upload only ms", t(up_only))
dev = m.DeviceBuffer(raw.nbytes + 64); dev.copy_from_host(raw)
def dec_only():
    tb = m.parquet_decode_device(raw, dev.ptr, COLS); return m.scan_aggregate(prog, True, tb, [], spec).to_rows()
print("decode+agg only ms", t(dec_only))
def pipelined(n=6):
    nxt = m.AsyncUpload(raw)
    for k in range(n):
        cur, nxt = nxt, (m.AsyncUpload(raw) if k + 1 < n else None)
        tb = m.parquet_decode_device(raw, cur.wait(), COLS)
        r = m.scan_aggregate(prog, True, tb, [], spec).to_rows()
        cur.free()
m.sync(); t0 = time.perf_counter(); pipelined(6); m.sync(); print("pipelined ms/step", (time.perf_counter() - t0) / 6 * 1e3)
for name, fn in (("resident", lambda: [dec_only() for _ in range(6)]), ("pipelined", lambda: pipelined(6))):
    m.sync(); m.profile_enable(True); fn(); m.sync(); rep = m.profile_report(); m.profile_enable(False)
    print(name, {k["name"]: round(k["ms"] / 6, 3) for k in rep})
t0 = time.perf_counter(); u = m.AsyncUpload(raw); t1 = time.perf_counter(); u.wait(); m.sync(); print("AsyncUpload call returns in ms", (t1 - t0) * 1e3); u.free()
