"""Per-operator rows/s and achieved GB/s against the measured HBM roofline (BASELINE metric:
"rows/sec per operator").  Columns are generated on the host, copied once, operators timed with the
library's per-kernel CUDA events (profile hooks) + whole-call events.  Algorithmic bytes per SURVEY §8d."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spark_rapids_b200 as m

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
REPS = 3
m.init(0, 24 << 30)   # pre-grown pool (Rmm.initialize analogue): first-touch pool growth costs ~45 ms/GB
rng = np.random.default_rng(42)
peak = 6585.1
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def timed(fn, reps=REPS):
    r = None
    for _ in range(2):
        del r
        r = fn()
    del r
    r = None
    m.sync()
    m.profile_enable(True)
    e0, e1 = m.Event(), m.Event()
    e0.record()
    for _ in range(reps):
        del r          # release the previous result first: otherwise the pool has to grow to hold two
        r = fn()
    e1.record()
    m.sync()
    ms = e0.elapsed_ms(e1) / reps
    prof = m.profile_report()
    m.profile_enable(False)
    return ms, {k["name"]: k["ms"] / reps for k in prof}, r


rows = []


def report(name, n, alg_bytes, ms, kernels):
    gbs = alg_bytes / 1e9 / (ms / 1e3)
    rows.append({"operator": name, "rows": n, "ms": round(ms, 4), "rows_per_sec": n / (ms / 1e3), "alg_GBps": round(gbs, 1), "frac_of_hbm": round(gbs / peak, 4),
                 "kernels_ms": {k: round(v, 4) for k, v in kernels.items()}})
    print("%-46s %12d rows %9.3f ms %10.2f Mrows/s %8.1f GB/s (%.1f%% of %.0f)" % (name, n, ms, n / ms / 1e3, gbs, 100 * gbs / peak, peak), flush=True)


# ---- inputs (q3-shaped)
key = rng.integers(0, N // 4, N, dtype=np.int64)            # l_orderkey-like foreign key (4 rows per key)
val = rng.integers(90000, 10494951, N, dtype=np.int64)      # decimal(12,2)
date = rng.integers(8036, 10562, N, dtype=np.int32)
t = m.Table.from_columns([m.Column.from_numpy(key), m.Column.from_numpy(val, dtype=m.DECIMAL64, scale=2), m.Column.from_numpy(date, dtype=m.DATE32)])
ck, cv, cd = m.col(0, m.INT64, nullable=False), m.col(1, m.DECIMAL64, 12, 2, nullable=False), m.col(2, m.DATE32, nullable=False)

# a2 filter, 50% selectivity, 3 columns compacted (20 B/row in, 10 B/row out)
pred = m.Program([cd < m.lit(9299, m.DATE32)])
ms, k, out = timed(lambda: m.filter(pred, t))
report("filter 50%% (3 cols, 20 B/row)", N, N * 20 + out.num_rows * 20, ms, k)
ms, k, _ = timed(lambda: m.filter_count(pred, t))
report("filter count-only (date < k, 4 B/row)", N, N * 4, ms, k)

# a1 project: decimal multiply -> DECIMAL128 column
proj = m.Program([cv * cv])
ms, k, _ = timed(lambda: m.project(proj, t))
report("project dec64*dec64 -> dec128 (8 in, 16 out)", N, N * 24, ms, k)

# a4 group-by, high cardinality (N/4 groups): global table regime
spec = [(m.AGG_SUM, 1, m.DECIMAL128, 2, 22), (m.AGG_COUNT_ALL, 0)]
gb_rows = min(N, 30_000_000)
tg = m.slice_table(t, 0, gb_rows)
ms, k, out = timed(lambda: m.groupby(tg, [0], spec), reps=2)
report("group-by i64 key, %d groups, sum+count" % out.num_rows, gb_rows, gb_rows * 16 + out.num_rows * 32, ms, k)
# low cardinality (2557 dates > 128 -> global; use 100 keys for the shared-memory regime)
k100 = m.Column.from_numpy((key % 100).astype(np.int32))
tl = m.Table.from_columns([k100, t.column(1)])
ms, k, out = timed(lambda: m.groupby(tl, [0], spec))
report("group-by i32 key, 100 groups, sum+count", N, N * 12, ms, k)

# a6/a7 join: build N/4 unique keys, probe N rows, gather 2 payload columns per side
nb = N // 4
bkeys = rng.permutation(nb).astype(np.int64)
bt = m.Table.from_columns([m.Column.from_numpy(bkeys), m.Column.from_numpy(rng.integers(8036, 10562, nb, dtype=np.int32), dtype=m.DATE32)])
bk = m.Table.from_columns([bt.column(0)])
pk = m.Table.from_columns([t.column(0)])
ms, k, ht = timed(lambda: m.JoinHashTable(bk), reps=2)
report("join build (i64 key)", nb, nb * 8 + nb * 2 * 8, ms, k)
ms, k, maps = timed(lambda: ht.probe(pk, m.JOIN_INNER), reps=2)
matched = len(maps[0])
report("join probe inner (i64 key, %d matches)" % matched, N, 2 * N * 8 + matched * 8, ms, k)
ms, k, _ = timed(lambda: (m.gather(t, maps[0]), m.gather(bt, maps[1])), reps=2)
report("gather payload (20 B stream + 12 B build per row)", matched, matched * (8 + 2 * 32), ms, k)

# a8 sort: single i64 key argsort + gather of 3 columns
ns = min(N, 30_000_000)
ts = m.slice_table(t, 0, ns)
ms, k, _ = timed(lambda: m.sort_order(ts, [(0, 1, 1)]), reps=2)
report("sort_order i64 key (argsort)", ns, ns * 8 + ns * 4, ms, k)
ms, k, _ = timed(lambda: m.order_by(ts, [(1, 0, 0), (2, 1, 1)]), reps=2)
report("order_by (dec64 desc, date asc) + gather 3 cols", ns, ns * 12 + ns * 4 + 2 * ns * 20, ms, k)
ms, k, _ = timed(lambda: m.top_n(ts, [(1, 0, 0), (2, 1, 1)], 10), reps=2)
report("top-10 (dec64 desc, date asc)", ns, ns * 12, ms, k)

# a9 murmur3 + partition into 8 / 200
ms, k, _ = timed(lambda: m.murmur3(t, [0], 42))
report("murmur3 (i64 key)", N, N * 12, ms, k)
ms, k, _ = timed(lambda: m.hash_partition(ts, [0], 8), reps=2)
report("hash_partition 8 parts (3 cols, 20 B/row)", ns, ns * 8 + 2 * ns * 20, ms, k)

# a11 rows, a12 concat
nr = min(N, 10_000_000)
tr = m.slice_table(t, 0, nr)
ms, k, _ = timed(lambda: m.concat([tr, tr]), reps=2)
report("concat 2 x (3 cols, 20 B/row)", 2 * nr, 2 * 2 * nr * 20, ms, k)

json.dump({"n": N, "hbm_peak_gbs": peak, "operators": rows}, open("gpurun_out/operators_bench.json", "w"), indent=1)
