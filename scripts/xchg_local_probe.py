"""Single-rank probe of the fused exchange kernel: a world-of-1 communicator sends every row to its own arena, so the
scatter kernel runs with local stores only.  Separates the kernel's own efficiency from the NVLink store rate seen at N > 1."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spark_rapids_b200 as m

m.init(0, 8 << 30)
comm = m.Comm(m.Comm.unique_id(), 0, 1)
n = int(os.environ.get("ROWS", 20_000_000))
rng = np.random.default_rng(1)
t = m.Table.from_columns([m.Column.from_numpy(rng.integers(0, 1 << 40, n)), m.Column.from_numpy(rng.integers(0, 1 << 40, n)),
                          m.Column.from_numpy(rng.integers(0, 1 << 40, n))])
print("fused_ready", comm.fused_ready())
for _ in range(3):
    out, _ = comm.exchange_hash(t, [0])
m.sync()
e0, e1 = m.Event(), m.Event()
m.profile_enable(True)
e0.record()
for _ in range(5):
    out, _ = comm.exchange_hash(t, [0])
e1.record(); m.sync()
print("rows", n, "ms/call", e0.elapsed_ms(e1) / 5, "out rows", out.num_rows)
for k in m.profile_report():
    print(k["name"], k["ms"] / k["launches"], "ms/launch")
comm.close()
