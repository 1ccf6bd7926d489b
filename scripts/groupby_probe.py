import sys
import numpy as np
sys.path.insert(0, ".")
import spark_rapids_b200 as m
m.init(0, 8 << 30)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
rng = np.random.default_rng(1)
k = m.Column.from_numpy(rng.integers(0, 100, n).astype(np.int32))
v = m.Column.from_numpy(rng.integers(90000, 10494951, n, dtype=np.int64), dtype=m.DECIMAL64, scale=2)
t = m.Table.from_columns([k, v])
spec = [(m.AGG_SUM, 1, m.DECIMAL128, 2, 22), (m.AGG_COUNT_ALL, 0)]
for _ in range(3):
    r = m.groupby(t, [0], spec)
m.profile_enable(True)
for _ in range(3):
    r = m.groupby(t, [0], spec)
for kk in m.profile_report():
    print(kk["name"], kk["ms"] / kk["launches"])
