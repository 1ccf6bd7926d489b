import sys, time
import numpy as np
sys.path.insert(0, ".")
import spark_rapids_b200 as m
m.init(0)
n = 50_000_000
rng = np.random.default_rng(1)
t = m.Table.from_columns([m.Column.from_numpy(rng.integers(0, 1 << 40, n, dtype=np.int64)), m.Column.from_numpy(rng.integers(0, 100, n, dtype=np.int64)),
                          m.Column.from_numpy(rng.integers(8036, 10562, n, dtype=np.int32), dtype=m.DATE32)])
pred = m.Program([m.col(2, m.DATE32, nullable=False) < m.lit(9299, m.DATE32)])
for i in range(6):
    m.sync(); t0 = time.perf_counter()
    r = m.filter(pred, t)
    m.sync(); t1 = time.perf_counter()
    del r
    m.sync(); t2 = time.perf_counter()
    print("filter call %.2f ms, release %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
for i in range(4):
    m.sync(); t0 = time.perf_counter()
    b = m.DeviceBuffer(400_000_000)
    m.sync(); t1 = time.perf_counter()
    del b
    m.sync(); t2 = time.perf_counter()
    print("alloc 400MB %.3f ms, free %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
