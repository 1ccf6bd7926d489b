"""TPC-H q1 fused filter+project+group-by (2 one-char string keys, 4 groups) over resident columns: 46 B/row."""
import sys
import numpy as np
sys.path.insert(0, ".")
import spark_rapids_b200 as m
from spark_rapids_b200 import _init as mi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
m.init(0, 16 << 30)
rng = np.random.default_rng(42)
def strcol(choices, idx):
    chars = np.frombuffer("".join(choices).encode(), dtype=np.uint8)[idx].copy()
    offs = np.arange(len(idx) + 1, dtype=np.int32)
    import ctypes
    out = ctypes.c_int64()
    m.check(m.lib.b2_column_from_host(m.STRING, 0, len(idx), mi._ptr(chars), None, mi._ptr(offs), ctypes.byref(out)))
    return m.Column(out.value)
rf = strcol("ANR", rng.integers(0, 3, n))
ls = strcol("FO", rng.integers(0, 2, n))
dec = lambda lo, hi, mul=1: m.Column.from_numpy(rng.integers(lo, hi, n, dtype=np.int64) * mul, dtype=m.DECIMAL64, scale=2)
qty, price, disc, tax = dec(1, 51, 100), dec(90000, 10494951), dec(0, 11), dec(0, 9)
ship = m.Column.from_numpy(rng.integers(8036, 10562, n, dtype=np.int32), dtype=m.DATE32)
t = m.Table.from_columns([rf, ls, qty, price, disc, tax, ship])
c = [m.col(0, m.STRING, nullable=False), m.col(1, m.STRING, nullable=False)] + [m.col(i, m.DECIMAL64, 12, 2, nullable=False) for i in (2, 3, 4, 5)] + [m.col(6, m.DATE32, nullable=False)]
one = m.lit(1, m.DECIMAL32, 1, 0)
pred = c[6] <= m.lit(10471, m.DATE32)
disc_price = c[3] * (one - c[4])
charge = disc_price * (one + c[5])
prog = m.Program([pred, c[0], c[1], c[2], c[3], disc_price, charge, c[4]])
specs = [(m.AGG_SUM, 2, m.DECIMAL128, 2, 22), (m.AGG_SUM, 3, m.DECIMAL128, 2, 22), (m.AGG_SUM, 4, m.DECIMAL128, 4, 36),
         (m.AGG_SUM, 5, m.DECIMAL128, 6, 38), (m.AGG_COUNT, 2), (m.AGG_COUNT, 3), (m.AGG_COUNT, 6), (m.AGG_COUNT_ALL, 0)]
for _ in range(3):
    r = m.scan_aggregate(prog, True, t, [0, 1], specs)
print(sorted(r.to_rows())[:2])
m.profile_enable(True)
for _ in range(5):
    r = m.scan_aggregate(prog, True, t, [0, 1], specs)
for k in m.profile_report():
    per = k["ms"] / k["launches"]
    print("%-28s %.4f ms  %8.1f GB/s  %8.1f Mrows/s" % (k["name"], per, n * 46 / 1e9 / (per / 1e3), n / per / 1e3))

# ---- breakdown
def tm(name, fn, bpr):
    for _ in range(2):
        fn()
    m.profile_enable(True)
    for _ in range(3):
        fn()
    for k in m.profile_report():
        per = k["ms"] / k["launches"]
        print("%-44s %-24s %.4f ms %8.1f GB/s" % (name, k["name"], per, n * bpr / 1e9 / (per / 1e3)))
    m.profile_enable(False)
p_proj = m.Program([disc_price, charge])
tm("project disc_price, charge", lambda: m.project(p_proj, t), 24 + 32)
p_dp = m.Program([disc_price])
tm("project disc_price only", lambda: m.project(p_dp, t), 16 + 16)
p_keys = m.Program([pred, c[0], c[1]])
tm("group-by keys only, count(*)", lambda: m.scan_aggregate(p_keys, True, t, [0, 1], [(m.AGG_COUNT_ALL, 0)]), 14)
p_one = m.Program([pred, c[0], c[1], c[2]])
tm("group-by, sum(qty)", lambda: m.scan_aggregate(p_one, True, t, [0, 1], [(m.AGG_SUM, 2, m.DECIMAL128, 2, 22)]), 22)
p_four = m.Program([pred, c[0], c[1], c[2], c[3], c[4], c[5]])
tm("group-by, 4 plain sums", lambda: m.scan_aggregate(p_four, True, t, [0, 1], [(m.AGG_SUM, k, m.DECIMAL128, 2, 22) for k in (2, 3, 4, 5)]), 46)
