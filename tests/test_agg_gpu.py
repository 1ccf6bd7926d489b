"""a3/a4/a5 parity: reductions and hash group-by incl. Spark-exact decimal SUM
(GpuAggregateExec.scala:540-585; aggregateFunctions.scala:1106-1290)."""
import numpy as np
import pytest

from oracle import spark_cpu as O
from tests import datagen as G

pytestmark = pytest.mark.gpu


def check_groupby(b2, ocols, keys, specs):
    t = G.to_b2_table(b2, ocols)
    got = b2.groupby(t, keys, specs).to_rows()
    exp = O.rows_of(O.groupby_cols(ocols, keys, specs))
    assert G.norm_rows(got) == G.norm_rows(exp)


def test_reference_known_answers(b2):
    """HashAggregateRetrySuite.scala:34-52, 117-206"""
    t = b2.Table.from_columns([b2.Column.from_numpy(np.array([5, 0, 3, 1], np.int64), valid=[True, False, True, True])])
    assert b2.reduce(t, [(b2.AGG_SUM, 0, b2.INT64)]).to_rows() == [(9,)]
    k = b2.Column.from_numpy(np.array([5, 0, 1, 1], np.int32), valid=[True, False, True, True])
    v = b2.Column.from_numpy(np.array([1, 2, 3, 4], np.int64))
    got = b2.groupby(b2.Table.from_columns([k, v]), [0], [(b2.AGG_SUM, 1, b2.INT64)]).to_rows()
    assert sorted(got, key=repr) == sorted([(5, 1), (None, 2), (1, 7)], key=repr)


@pytest.mark.parametrize("n", [0, 1, 100, 5000])
def test_reduce(b2, n):
    rng = np.random.default_rng(n)
    ocols = [G.gen_column(rng, (O.INT64, 0, 0), n), G.gen_column(rng, (O.INT32, 0, 0), n), G.gen_column(rng, (O.DECIMAL64, 12, 2), n),
             G.gen_column(rng, (O.DECIMAL128, 38, 4), n), G.gen_column(rng, (O.DECIMAL128, 20, 4), n)]
    specs = [(O.AGG_SUM, 0, O.INT64, 0, 0), (O.AGG_COUNT, 0), (O.AGG_MIN, 0), (O.AGG_MAX, 0), (O.AGG_SUM, 1, O.INT64, 0, 0),
             (O.AGG_MIN, 1), (O.AGG_SUM, 2, O.DECIMAL128, 2, 22), (O.AGG_SUM, 3, O.DECIMAL128, 4, 38), (O.AGG_SUM, 4, O.DECIMAL128, 4, 30),
             (O.AGG_COUNT_ALL, 0)]
    t = G.to_b2_table(b2, ocols)
    got = b2.reduce(t, specs).to_rows()
    exp = O.rows_of(O.reduce_cols(ocols, specs))
    assert got == exp


def test_reduce_all_null_and_overflow(b2):
    n = 100
    nulls = O.OCol(np.zeros(n, np.int64), np.zeros(n, bool), (O.INT64, 0, 0))
    big = O.OCol(np.array([10**38 - 1] * n, dtype=object), np.ones(n, bool), (O.DECIMAL128, 38, 0))
    specs = [(O.AGG_SUM, 0, O.INT64, 0, 0), (O.AGG_COUNT, 0), (O.AGG_MAX, 0), (O.AGG_SUM, 1, O.DECIMAL128, 0, 38)]
    got = b2.reduce(G.to_b2_table(b2, [nulls, big]), specs).to_rows()
    assert got == [(None, 0, None, None)]
    assert got == O.rows_of(O.reduce_cols([nulls, big], specs))


@pytest.mark.parametrize("keytyp", [(O.INT32, 0, 0), (O.INT64, 0, 0), (O.STRING, 0, 0), (O.FLOAT64, 0, 0), (O.DECIMAL64, 12, 2), (O.DATE32, 0, 0)])
@pytest.mark.parametrize("card", [1, 4, 100, 3000])
def test_groupby_single_key(b2, keytyp, card):
    rng = np.random.default_rng(card + keytyp[0])
    n = 6000
    ocols = [G.gen_column(rng, keytyp, n, distinct=card), G.gen_column(rng, (O.INT64, 0, 0), n), G.gen_column(rng, (O.DECIMAL64, 12, 2), n),
             G.gen_column(rng, (O.DECIMAL128, 38, 4), n)]
    specs = [(O.AGG_SUM, 1, O.INT64, 0, 0), (O.AGG_COUNT, 1), (O.AGG_MIN, 1), (O.AGG_MAX, 1), (O.AGG_SUM, 2, O.DECIMAL128, 2, 22),
             (O.AGG_SUM, 3, O.DECIMAL128, 4, 38), (O.AGG_COUNT_ALL, 0)]
    check_groupby(b2, ocols, [0], specs)


def test_groupby_float_keys_nan_zero(b2):
    """NaN keys are one group, -0.0 and 0.0 are one group (GpuAggregateExec.scala:565-568)"""
    k = O.OCol(np.array([np.nan, 0.0, -0.0, np.nan, 1.5, 0.0]), np.array([1, 1, 1, 1, 1, 0], bool), (O.FLOAT64, 0, 0))
    v = O.OCol(np.arange(6, dtype=np.int64), np.ones(6, bool), (O.INT64, 0, 0))
    got = b2.groupby(G.to_b2_table(b2, [k, v]), [0], [(O.AGG_SUM, 1, O.INT64, 0, 0)]).to_rows()
    sums = sorted(r[1] for r in got)
    assert sums == sorted([0 + 3, 1 + 2, 4, 5])


def test_groupby_multi_key_q3_shape(b2):
    """q3 keys: (l_orderkey i64, o_orderdate date, o_shippriority i32), many groups -> global-table regime"""
    rng = np.random.default_rng(3)
    n = 50000
    ok = G.gen_column(rng, (O.INT64, 0, 0), n, null_frac=0.0, distinct=9000)
    od = O.OCol((ok.values % 2400 + 8000).astype(np.int32), np.ones(n, bool), (O.DATE32, 0, 0))
    sp = O.OCol(np.zeros(n, np.int32), np.ones(n, bool), (O.INT32, 0, 0))
    rev = G.gen_column(rng, (O.DECIMAL128, 26, 4), n, null_frac=0.0, small=True)
    check_groupby(b2, [ok, od, sp, rev], [0, 1, 2], [(O.AGG_SUM, 3, O.DECIMAL128, 4, 36)])


def test_scan_aggregate_q1_shape(b2):
    """q1: filter + project + group-by (2 one-char string keys, 4 groups) in ONE kernel"""
    rng = np.random.default_rng(1)
    n = 30000
    dec = (O.DECIMAL64, 12, 2)
    rf = O.OCol(np.array([b"A", b"N", b"R"], dtype=object)[rng.integers(0, 3, n)], np.ones(n, bool), (O.STRING, 0, 0))
    ls = O.OCol(np.array([b"F", b"O"], dtype=object)[rng.integers(0, 2, n)], np.ones(n, bool), (O.STRING, 0, 0))
    qty = G.gen_column(rng, dec, n, 0.0, distinct=50)
    price = G.gen_column(rng, dec, n, 0.0, small=True)
    disc = G.gen_column(rng, dec, n, 0.0, distinct=11)
    tax = G.gen_column(rng, dec, n, 0.0, distinct=9)
    ship = O.OCol(rng.integers(8036, 10561, n).astype(np.int32), np.ones(n, bool), (O.DATE32, 0, 0))
    ocols = [rf, ls, qty, price, disc, tax, ship]
    t = G.to_b2_table(b2, ocols)
    c = [G.b2_expr_col(b2, i, oc) for i, oc in enumerate(ocols)]
    one = b2.lit(1, b2.DECIMAL32, 1, 0)
    pred = c[6] <= b2.lit(10471, b2.DATE32)
    disc_price = c[3] * (one - c[4])
    charge = disc_price * (one + c[5])
    outs = [pred, c[0], c[1], c[2], c[3], disc_price, charge, c[4]]
    specs = [(O.AGG_SUM, 2, O.DECIMAL128, 2, 22), (O.AGG_SUM, 3, O.DECIMAL128, 2, 22), (O.AGG_SUM, 4, O.DECIMAL128, 4, 36),
             (O.AGG_SUM, 5, O.DECIMAL128, 6, 38), (O.AGG_COUNT, 2), (O.AGG_SUM, 6, O.DECIMAL128, 2, 22), (O.AGG_COUNT_ALL, 0)]
    got = b2.scan_aggregate(b2.Program(outs), True, t, [0, 1], specs).to_rows()
    keep = O.eval_expr(pred.sexpr, ocols)
    proj = O.filter_cols([O.eval_expr(e.sexpr, ocols) for e in outs[1:]], keep)
    exp = O.rows_of(O.groupby_cols(proj, [0, 1], specs))
    assert len(got) == 6
    assert G.norm_rows(got) == G.norm_rows(exp)


def test_distinct_count(b2):
    rng = np.random.default_rng(8)
    a = G.gen_column(rng, (O.INT64, 0, 0), 20000, distinct=777, null_frac=0.05)
    t = G.to_b2_table(b2, [a])
    keys = set(None if not ok else int(v) for v, ok in zip(a.values, a.valid))
    assert b2.distinct_count(t, [0]) == len(keys)
