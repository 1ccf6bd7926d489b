"""Known answers the reference's own Scala suites assert (tests/src/test/scala/com/nvidia/spark/rapids/), reproduced through
the C ABI / operator layer.  These are the few reference-held vectors for the sort / coalesce / sub-partition rows."""
import numpy as np
import pytest

from oracle import spark_cpu as O
from tests import datagen as G

pytestmark = pytest.mark.gpu
I32 = (O.INT32, 0, 0)


def _ints(vals):
    return O.OCol(np.asarray(vals, dtype=np.int32), np.ones(len(vals), bool), I32)


def test_sort_each_batch_like_GpuSortRetrySuite(b2):
    """GpuSortRetrySuite.scala:36-48,178-200: every batch holds (50 until 100) ++ (0 until 50); sorted each batch the values
    read 0, 1, 2, ... 99 and every batch keeps its 100 rows"""
    from spark_rapids_b200 import execs as E
    batch = list(range(50, 100)) + list(range(0, 50))
    src = E.GpuBatchSource([G.to_b2_table(b2, [_ints(batch)]) for _ in range(2)])
    outs = list(E.GpuSortExec([(0, 1, 1)], src, global_sort=False))
    assert len(outs) == 2
    for t in outs:
        assert t.num_rows == 100 and t.column(0).to_pylist() == list(range(100))
    # "GPU out-of-core sort": one final batch with all 200 rows (GpuSortRetrySuite.scala:50-62), ascending
    full = E.GpuSortExec([(0, 1, 1)], E.GpuBatchSource([G.to_b2_table(b2, [_ints(batch)]) for _ in range(2)])).collect()
    assert full.num_rows == 200 and full.column(0).to_pylist() == sorted(batch + batch)


def test_coalesce_like_GpuCoalesceBatchesSuite(b2):
    """GpuCoalesceBatchesSuite.scala:55-75: mixedDf (SparkQueryCompareTestSuite.scala:1335-1357: 14 rows, 5 columns: int, long,
    double, string, decimal(15,5)) arriving one row per batch (TargetSize(1)) coalesces under TargetSize(100000) into ONE batch
    of 14 rows x 5 columns; numOutputRows = 14, numOutputBatches = 1"""
    from spark_rapids_b200 import execs as E
    nul = "\x00"
    rows = [(99, 100, 1.0, "A", 120000), (98, 200, 2.0, "B", 130000), (97, 300, 3.0, "C", 140000), (99, 400, 4.0, "D", 150000),
            (98, 500, 5.0, "E", 160000), (97, -100, 6.0, "F", 170000), (96, -500, 0.0, "G", None), (95, -700, 8.0, "EҀҁ", 190000),
            (2**31 - 1, -2**63, float("inf"), nul, 200000), (-2**31, 2**63 - 1, float("nan"), nul, 10012300),
            (None, None, None, "actions are judged by intentions", None), (94, -900, 9.0, "g\nH", 30036900),
            (92, -1200, 12.0, "IJ\"ĀāԀԁ", -147000000), (90, 1500, 15.0, "휠휡", -2223450)]
    assert len(rows) == 14
    typs = [I32, (O.INT64, 0, 0), (O.FLOAT64, 0, 0), (O.STRING, 0, 0), (O.DECIMAL64, 15, 5)]

    def one_row(r):
        return G.to_b2_table(b2, [O.ocol([v], t) for v, t in zip(r, typs)])
    co = E.GpuCoalesceBatches(E.GpuBatchSource([one_row(r) for r in rows]), 100000)
    out = list(co)
    assert len(out) == 1 and out[0].num_rows == 14 and out[0].num_columns == 5
    assert co.metrics["numOutputRows"] == 14 and co.metrics["numOutputBatches"] == 1
    got = out[0].to_rows()
    for g, e in zip(got, rows):
        assert g[0] == e[0] and g[1] == e[1] and g[3] == e[3] and g[4] == e[4]
        assert (g[2] is None and e[2] is None) or (g[2] != g[2] and e[2] != e[2]) or g[2] == e[2]


def test_sub_partitioner_like_GpuSubPartitionSuite(b2):
    """GpuSubPartitionSuite.scala:86-105: ints (1,2,2,3,3,3) hashed with seed 100 into 5 partitions: every row lands in exactly
    one partition (6 rows in total) and equal keys share a partition"""
    keys = [1, 2, 2, 3, 3, 3]
    t = G.to_b2_table(b2, [_ints(keys)])
    out, offs = b2.hash_partition(t, [0], 5, seed=100)
    assert offs[0] == 0 and offs[-1] == 6 and len(offs) == 6
    vals = out.column(0).to_pylist()
    assert sorted(vals) == sorted(keys)
    where = {}
    for p in range(5):
        for v in vals[offs[p]:offs[p + 1]]:
            assert where.setdefault(v, p) == p
    # Spark's Murmur3 with seed 100 is what the sub-partitioner uses (GpuSubPartitionHashJoin.scala:86-226): ids match the oracle
    from oracle import spark_hash as H
    assert [int(x) for x in H.partition_ids([_ints(keys)], 5, seed=100)] == [where[v] for v in keys]
