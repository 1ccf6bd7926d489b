"""q6 fused filter+project+sum over resident columns (no Parquet): used for ncu captures of the VM kernels."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import spark_rapids_b200 as m
from oracle import tpch
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
m.init(0)
cols = tpch.lineitem_q6_columns(n, 42)
t = m.Table.from_columns([m.Column.from_numpy(cols["l_shipdate"], dtype=m.DATE32),
                          m.Column.from_numpy(cols["l_discount"], dtype=m.DECIMAL64, scale=2),
                          m.Column.from_numpy(cols["l_quantity"], dtype=m.DECIMAL64, scale=2),
                          m.Column.from_numpy(cols["l_extendedprice"], dtype=m.DECIMAL64, scale=2)])
prog, spec = bench.build_q6(m)
pred_prog = m.Program([prog.exprs[0]])
rev_prog = m.Program([prog.exprs[1]])
one = m.Program([m.col(0, m.DATE32, nullable=False) >= m.lit(8766, m.DATE32)])
variants = {
    "q6 fused filter+project+sum (28 B/row)": (lambda: m.scan_aggregate(prog, True, t, [], spec), 28),
    "predicate count only (28 B/row)": (lambda: m.filter_count(pred_prog, t), 28),
    "single compare count (4 B/row)": (lambda: m.filter_count(one, t), 4),
    "project+sum, no predicate (16 B/row)": (lambda: m.scan_aggregate(rev_prog, False, t, [], spec), 16),
    "filter (materialise 4 cols, 2% pass)": (lambda: m.filter(pred_prog, t), 28),
    "project revenue column (16 B in, 16 B out)": (lambda: m.project(rev_prog, t), 32),
}
r = m.scan_aggregate(prog, True, t, [], spec).to_rows()
assert r[0][0] == tpch.q6_numpy(cols), r
for name, (fn, bpr) in variants.items():
    for _ in range(3):
        fn()
    m.profile_enable(True)
    for _ in range(reps):
        fn()
    rep = m.profile_report()
    m.profile_enable(False)
    for k in rep:
        per = k["ms"] / k["launches"]
        print("%-45s %-26s %.4f ms  %7.1f GB/s" % (name, k["name"], per, n * bpr / 1e9 / (per / 1e3)))
