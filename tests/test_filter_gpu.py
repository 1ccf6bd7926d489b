"""a2 parity: GpuFilter (basicPhysicalOperators.scala:1148-1224): NULL predicate drops the row,
row order is preserved, every column (incl. strings, validity) is compacted; count-only path."""
import numpy as np
import pytest

from oracle import spark_cpu as O
from tests import datagen as G

pytestmark = pytest.mark.gpu

SCHEMA = [(O.INT64, 0, 0), (O.INT32, 0, 0), (O.INT8, 0, 0), (O.INT16, 0, 0), (O.FLOAT64, 0, 0), (O.DECIMAL128, 30, 4),
          (O.DECIMAL64, 12, 2), (O.STRING, 0, 0), (O.BOOL8, 0, 0)]


@pytest.mark.parametrize("n", [0, 1, 33, 1024, 1025, 40000])
@pytest.mark.parametrize("sel", [0.0, 0.02, 0.5, 1.0])
def test_filter_all_types(b2, n, sel):
    rng = np.random.default_rng(n + int(sel * 100))
    ocols = [G.gen_column(rng, t, n) for t in SCHEMA]
    key = O.OCol(rng.random(n), rng.random(n) > 0.1, (O.FLOAT64, 0, 0))
    ocols.append(key)
    t = G.to_b2_table(b2, ocols)
    pred = G.b2_expr_col(b2, len(SCHEMA), key) < b2.lit(float(sel), b2.FLOAT64)
    prog = b2.Program([pred])
    out = b2.filter(prog, t)
    keep = O.eval_expr(pred.sexpr, ocols)
    exp = O.filter_cols(ocols, keep)
    assert out.num_rows == len(exp[0])
    for i in range(len(ocols)):
        G.assert_col_equal(out.column(i), exp[i])
    assert b2.filter_count(prog, t) == len(exp[0])


def test_filter_config0_long_gt_k(b2):
    """BASELINE config 0: 1M-row single long column, filter(col > k).count() at 1% / 50% / 99%"""
    rng = np.random.default_rng(42)
    n = 1_000_000
    vals = rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64)
    t = b2.Table.from_columns([b2.Column.from_numpy(vals)])
    c = b2.col(0, b2.INT64, nullable=False)
    for q in (0.01, 0.5, 0.99):
        k = int(np.quantile(vals.astype(np.float64), q))
        prog = b2.Program([c > b2.lit(k, b2.INT64)])
        assert b2.filter_count(prog, t) == int((vals > k).sum())
        out = b2.filter(prog, t)
        got, _ = out.column(0).to_numpy()
        assert np.array_equal(got, vals[vals > k])


def test_filter_mask_column(b2):
    rng = np.random.default_rng(9)
    n = 5000
    ocols = [G.gen_column(rng, (O.INT64, 0, 0), n), G.gen_column(rng, (O.STRING, 0, 0), n)]
    mask = G.gen_column(rng, (O.BOOL8, 0, 0), n, null_frac=0.2)
    t = G.to_b2_table(b2, ocols)
    out = b2.filter_mask(t, G.to_b2_column(b2, mask))
    exp = O.filter_cols(ocols, mask)
    for i in range(2):
        G.assert_col_equal(out.column(i), exp[i])


@pytest.mark.parametrize("n", [1, 255, 4096, 4097, 70001])
def test_conjunct_chain_non_nullable(b2, n):
    """WHERE chains over NOT NULL columns compile to the fused `acc AND (x cmp y)` instruction: every comparison
    operator, 4/8/16-byte and floating operands (NaN, -0.0), literal on either side, column-vs-column, a boolean
    column as the accumulator; compared with numpy on raw values"""
    rng = np.random.default_rng(n)
    i32 = rng.integers(-50, 50, n).astype(np.int32)
    i64 = rng.integers(-2**40, 2**40, n)
    j64 = i64 + rng.integers(-1, 2, n)
    f64 = rng.standard_normal(n)
    f64[rng.random(n) < 0.05] = np.nan
    f64[rng.random(n) < 0.05] = -0.0
    flag = rng.random(n) < 0.8
    t = b2.Table.from_columns([b2.Column.from_numpy(i32), b2.Column.from_numpy(i64), b2.Column.from_numpy(j64), b2.Column.from_numpy(f64),
                               b2.Column.from_numpy(flag)])
    a, b, c, d = (b2.col(0, b2.INT32, nullable=False), b2.col(1, b2.INT64, nullable=False), b2.col(2, b2.INT64, nullable=False),
                  b2.col(3, b2.FLOAT64, nullable=False))
    fl = b2.col(4, b2.BOOL8, nullable=False)
    L = lambda v, dt: b2.lit(v, dt)  # noqa: E731
    nan = np.isnan(f64)
    cases = [
        ((a >= L(-20, b2.INT32)) & (a < L(30, b2.INT32)) & (b != L(0, b2.INT64)) & (b <= c), (i32 >= -20) & (i32 < 30) & (i64 != 0) & (i64 <= j64)),
        (fl & (b == c) & (a > L(-45, b2.INT32)), flag & (i64 == j64) & (i32 > -45)),
        ((L(0.0, b2.FLOAT64) <= d) & (d < L(1.5, b2.FLOAT64)) & fl, (~nan) & (f64 >= 0.0) & (f64 < 1.5) & flag),   # -0.0 == 0.0, NaN is greatest
        ((d > L(0.5, b2.FLOAT64)) & (b > c), (nan | (f64 > 0.5)) & (i64 > j64)),
        ((b > c) & ((a == L(3, b2.INT32)) | (a == L(4, b2.INT32))) & (a != L(5, b2.INT32)), (i64 > j64) & ((i32 == 3) | (i32 == 4)) & (i32 != 5)),
    ]
    for pred, keep in cases:
        prog = b2.Program([pred])
        assert b2.filter_count(prog, t) == int(keep.sum())
        out = b2.filter(prog, t)
        assert np.array_equal(out.column(1).to_numpy()[0], i64[keep])
        # the same predicate in front of an aggregate (fused filter -> sum)
        r = b2.scan_aggregate(b2.Program([pred, b]), True, t, [], [(b2.AGG_SUM, 0, b2.INT64, 0, 0), (b2.AGG_COUNT_ALL, 0, b2.INT64, 0, 0)]).to_rows()[0]
        assert r[1] == int(keep.sum()) and (r[0] == int(i64[keep].sum()) if keep.any() else r[0] is None)
