"""a1 parity: fused expression evaluation (GpuProjectExec / GpuExpression.columnarEval) vs the oracle.
Modelled on the reference's integration_tests arithmetic_ops_test.py / cmp_test.py / logic_test.py:
same seeded generators with special values, CPU result == GPU result, bit-exact."""
import numpy as np
import pytest

from oracle import spark_cpu as O
from tests import datagen as G

pytestmark = pytest.mark.gpu

INT_TYPES = [(O.INT8, 0, 0), (O.INT16, 0, 0), (O.INT32, 0, 0), (O.INT64, 0, 0)]
FLOAT_TYPES = [(O.FLOAT32, 0, 0), (O.FLOAT64, 0, 0)]


def run(b2, exprs, ocols, approx=False):
    t = G.to_b2_table(b2, ocols)
    out = b2.project(b2.Program(exprs), t)
    assert out.num_rows == len(ocols[0])
    for i, e in enumerate(exprs):
        exp = O.eval_expr(e.sexpr, ocols)
        got = out.column(i)
        assert got.dtype == exp.typ[0], (got.dtype, exp.typ)
        if O.is_decimal(exp.typ[0]):
            assert got.scale == exp.typ[2]
        G.assert_col_equal(got, exp, approx)


@pytest.mark.parametrize("typ", INT_TYPES + FLOAT_TYPES)
@pytest.mark.parametrize("n", [0, 1, 31, 1024, 5000])
def test_binary_arith(b2, typ, n):
    rng = np.random.default_rng(n + typ[0])
    a, b = G.gen_column(rng, typ, n), G.gen_column(rng, typ, n, small=True)
    ca, cb = G.b2_expr_col(b2, 0, a), G.b2_expr_col(b2, 1, b)
    run(b2, [ca + cb, ca - cb, ca * cb, ca / cb, ca % cb, ca.pmod(cb), -ca, ca.abs()], [a, b])


@pytest.mark.parametrize("typ", INT_TYPES + FLOAT_TYPES + [(O.DATE32, 0, 0)])
def test_compare(b2, typ):
    rng = np.random.default_rng(11 + typ[0])
    n = 3000
    a = G.gen_column(rng, typ, n, distinct=7)
    b = G.gen_column(rng, typ, n, distinct=7)
    if typ[0] in (O.FLOAT32, O.FLOAT64):  # NaN / -0.0 cases (predicates.scala:155-331)
        for col in (a, b):
            col.values[rng.choice(n, 300, replace=False)] = np.nan
            col.values[rng.choice(n, 300, replace=False)] = -0.0
    ca, cb = G.b2_expr_col(b2, 0, a), G.b2_expr_col(b2, 1, b)
    run(b2, [ca == cb, ca != cb, ca < cb, ca <= cb, ca > cb, ca >= cb, ca.eq_null_safe(cb)], [a, b])


def test_compare_literal(b2):
    rng = np.random.default_rng(3)
    a = G.gen_column(rng, (O.INT64, 0, 0), 4096)
    ca = G.b2_expr_col(b2, 0, a)
    k = b2.lit(12345, b2.INT64)
    run(b2, [ca > k, k > ca, ca == k, ca + k, k - ca], [a])


def test_kleene_logic(b2):
    rng = np.random.default_rng(5)
    n = 2000
    a, b = G.gen_column(rng, (O.BOOL8, 0, 0), n, null_frac=0.3), G.gen_column(rng, (O.BOOL8, 0, 0), n, null_frac=0.3)
    ca, cb = G.b2_expr_col(b2, 0, a), G.b2_expr_col(b2, 1, b)
    run(b2, [ca & cb, ca | cb, ~ca, (ca & cb) | ~cb, ca.is_null(), cb.is_not_null(), ca & b2.lit(None, b2.BOOL8)], [a, b])


def test_if_coalesce(b2):
    rng = np.random.default_rng(6)
    n = 2000
    p = G.gen_column(rng, (O.BOOL8, 0, 0), n, null_frac=0.2)
    a, b = G.gen_column(rng, (O.INT64, 0, 0), n, null_frac=0.3), G.gen_column(rng, (O.INT64, 0, 0), n, null_frac=0.3)
    cp, ca, cb = G.b2_expr_col(b2, 0, p), G.b2_expr_col(b2, 1, a), G.b2_expr_col(b2, 2, b)
    run(b2, [b2.if_else(cp, ca, cb), ca.coalesce(cb), ca.coalesce(b2.lit(7, b2.INT64))], [p, a, b])


@pytest.mark.parametrize("src", INT_TYPES + FLOAT_TYPES)
@pytest.mark.parametrize("dst", INT_TYPES + FLOAT_TYPES)
def test_cast_numeric(b2, src, dst):
    rng = np.random.default_rng(src[0] * 13 + dst[0])
    a = G.gen_column(rng, src, 2000)
    ca = G.b2_expr_col(b2, 0, a)
    run(b2, [ca.cast(dst[0])], [a])


DEC = [((O.DECIMAL64, 12, 2), (O.DECIMAL64, 12, 2)), ((O.DECIMAL32, 7, 3), (O.DECIMAL64, 15, 1)),
       ((O.DECIMAL64, 18, 6), (O.DECIMAL64, 18, 2)), ((O.DECIMAL128, 30, 4), (O.DECIMAL64, 12, 2)),
       ((O.DECIMAL128, 38, 10), (O.DECIMAL128, 38, 10)), ((O.DECIMAL128, 26, 4), (O.DECIMAL64, 13, 2))]


@pytest.mark.parametrize("ta,tb", DEC)
def test_decimal_arith(b2, ta, tb):
    """arithmetic.scala:78-125 (add/sub overflow -> NULL), :411-640 (multiply incl. 256-bit path)"""
    rng = np.random.default_rng(ta[1] * 100 + tb[1])
    n = 1500
    a, b = G.gen_column(rng, ta, n), G.gen_column(rng, tb, n)
    ca, cb = G.b2_expr_col(b2, 0, a), G.b2_expr_col(b2, 1, b)
    exprs = [ca * cb, ca < cb, ca == cb]
    s = max(ta[2], tb[2])
    if max(ta[1] - ta[2], tb[1] - tb[2]) + s + 1 <= 38:
        exprs += [ca + cb, ca - cb]
    run(b2, exprs, [a, b])


def test_decimal_q1_expression(b2):
    """TPC-H q1: l_extendedprice * (1 - l_discount) * (1 + l_tax): (12,2)x(13,2)->(26,4)x(13,2)->(38,6)"""
    rng = np.random.default_rng(1)
    n = 4000
    t = (O.DECIMAL64, 12, 2)
    price, disc, tax = G.gen_column(rng, t, n, 0.05), G.gen_column(rng, t, n, 0.05, distinct=11), G.gen_column(rng, t, n, 0.05, distinct=9)
    cp, cd, ct = (G.b2_expr_col(b2, i, c) for i, c in enumerate((price, disc, tax)))
    one = b2.lit(1, b2.DECIMAL32, 1, 0)
    disc_price = cp * (one - cd)
    charge = disc_price * (one + ct)
    run(b2, [disc_price, charge], [price, disc, tax])


def test_decimal_cast(b2):
    rng = np.random.default_rng(2)
    a = G.gen_column(rng, (O.DECIMAL64, 12, 2), 2000)
    ca = G.b2_expr_col(b2, 0, a)
    run(b2, [ca.cast(b2.DECIMAL128, 22, 2), ca.cast(b2.DECIMAL64, 14, 4), ca.cast(b2.DECIMAL64, 10, 0), ca.cast(b2.DECIMAL32, 9, 1),
             ca.cast(b2.FLOAT64)], [a])
    i = G.gen_column(rng, (O.INT32, 0, 0), 2000)
    run(b2, [G.b2_expr_col(b2, 0, i).cast(b2.DECIMAL64, 12, 2), G.b2_expr_col(b2, 0, i).cast(b2.DECIMAL32, 9, 0)], [i])


def test_year_and_normalize(b2):
    rng = np.random.default_rng(4)
    d = G.gen_column(rng, (O.DATE32, 0, 0), 3000)
    f = G.gen_column(rng, (O.FLOAT64, 0, 0), 3000)
    run(b2, [G.b2_expr_col(b2, 0, d).year(), G.b2_expr_col(b2, 1, f).normalize_nan_zero()], [d, f])


def test_type_errors(b2):
    a = b2.col(0, b2.INT32)
    b = b2.col(1, b2.INT64)
    with pytest.raises(b2.B2Error):
        b2.Program([a + b])
    with pytest.raises(b2.B2Error):
        b2.Program([b2.col(0, b2.STRING) + b2.col(1, b2.STRING)])


@pytest.mark.parametrize("ta,tb", [((O.DECIMAL128, 22, 2), (O.DECIMAL128, 20, 0)), ((O.DECIMAL64, 12, 2), (O.DECIMAL64, 12, 2)),
                                   ((O.DECIMAL128, 38, 6), (O.DECIMAL32, 5, 0)), ((O.DECIMAL64, 18, 4), (O.DECIMAL128, 25, 3))])
def test_decimal_divide(b2, ta, tb):
    """GpuDecimalDivide (arithmetic.scala:903-1000): Spark result type, HALF_UP, divide by zero -> NULL"""
    rng = np.random.default_rng(ta[1] * 7 + tb[1])
    n = 1500
    a, b = G.gen_column(rng, ta, n), G.gen_column(rng, tb, n, small=True)
    b.values[rng.choice(n, 20, replace=False)] = 0
    ca, cb = G.b2_expr_col(b2, 0, a), G.b2_expr_col(b2, 1, b)
    run(b2, [ca / cb], [a, b])


def test_decimal_average_finalisation(b2):
    """avg = sum / count evaluated on the aggregation buffers (aggregateFunctions.scala:1555-1556, 1606-1616):
    decimal(22,2) sum / cast(count as decimal(20,0)), then cast to decimal(p+4, s+4) = (16,6)"""
    rng = np.random.default_rng(9)
    n = 800
    s = G.gen_column(rng, (O.DECIMAL128, 22, 2), n, small=True)
    c = O.OCol(rng.integers(0, 50, n).astype(np.int64), np.ones(n, bool), (O.INT64, 0, 0))
    cs, cc = G.b2_expr_col(b2, 0, s), G.b2_expr_col(b2, 1, c)
    avg = (cs / cc.cast(b2.DECIMAL128, 20, 0)).cast(b2.DECIMAL64, 16, 6)
    run(b2, [avg], [s, c])
