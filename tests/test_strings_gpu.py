"""f3: string predicates, Substring, In/InSet, CaseWhen through the fused VM (stringFunctions.scala:163,189,396,524,972;
GpuInSet.scala; conditionalExpressions.scala:322) vs the oracle; fused filter+project pruning; bound-reference pass-through."""
import numpy as np
import pytest

from oracle import spark_cpu as O
from tests import datagen as G

pytestmark = pytest.mark.gpu
STR = (O.STRING, 0, 0)
WORDS = [b"", b"a", b"ab", b"abc", b"BUILDING", b"BUILD", b"AUTOMOBILE", b"special requests", b"the special packages requests", "été".encode(),
         "日本語テキスト".encode(), b"a_b%c", b"100%", b"MAIL", b"SHIP", b"13-345", b"31-999", b"x" * 70]


def str_col(rng, n, null_frac=0.1):
    vals = np.array([WORDS[i] for i in rng.integers(0, len(WORDS), n)], dtype=object)
    return O.OCol(vals, rng.random(n) >= null_frac if null_frac else np.ones(n, bool), STR)


def run(b2, exprs, ocols):
    t = G.to_b2_table(b2, ocols)
    out = b2.project(b2.Program(exprs), t)
    for i, e in enumerate(exprs):
        G.assert_col_equal(out.column(i), O.eval_expr(e.sexpr, ocols))


@pytest.mark.parametrize("n", [0, 1, 33, 5000])
def test_string_compare_and_predicates(b2, n):
    rng = np.random.default_rng(n)
    a, b = str_col(rng, n), str_col(rng, n)
    ca, cb = G.b2_expr_col(b2, 0, a), G.b2_expr_col(b2, 1, b)
    run(b2, [ca == "BUILDING", ca != "BUILDING", ca < "abc", ca <= "abc", ca > "MAIL", ca >= "MAIL", b2.strlit("abc") > ca,
             ca == cb, ca < cb, ca >= cb,
             ca.startswith("BUILD"), ca.endswith("requests"), ca.contains("special"), ca.contains(""), ca.startswith("日本"),
             ca == b2.strlit(None)], [a, b])


def test_like_patterns(b2):
    rng = np.random.default_rng(5)
    a = str_col(rng, 4000)
    ca = G.b2_expr_col(b2, 0, a)
    pats = ["%special%requests%", "BUILD%", "%ING", "a_c", "_", "%", "", "a\\_b\\%c", "100\\%", "___", "%é%", "日_語%", "%x%x%", "ab%c%"]
    run(b2, [ca.like(p) for p in pats], [a])
    run(b2, [ca.like("a#_b#%c", escape="#")], [a])


def test_substring_windows_and_materialised(b2):
    rng = np.random.default_rng(6)
    a = str_col(rng, 3000)
    ca = G.b2_expr_col(b2, 0, a)
    # q22 shape: substring(c_phone, 1, 2) in ('13', '31', ...)
    run(b2, [ca.substr(1, 2) == "13", ca.substr(1, 2).isin(["13", "31", "ab"]), ca.substr(-3, 2) == "NG", ca.substr(0, 3) == "abc",
             ca.substr(2) == "b", ca.substr(-100, 102) == "ab", ca.substr(3, 0) == "", ca.substr(2, 2).startswith("本")], [a])
    col = G.to_b2_column(b2, a)
    for pos, ln in [(1, 2), (-3, 2), (0, 3), (2, 2**31 - 1), (-100, 102), (5, 0), (2, 3)]:
        got = b2.substring(col, pos, ln)
        exp = O.eval_expr(("substr", ("col", 0, STR), pos, ln), [a])
        G.assert_col_equal(got, exp)


def test_in_and_case_when(b2):
    rng = np.random.default_rng(7)
    n = 5000
    x = G.gen_column(rng, (O.INT32, 0, 0), n, distinct=20)
    d = G.gen_column(rng, (O.DECIMAL64, 12, 2), n, distinct=50)
    s = str_col(rng, n)
    cx, cd, cs = G.b2_expr_col(b2, 0, x), G.b2_expr_col(b2, 1, d), G.b2_expr_col(b2, 2, s)
    one, zero = b2.lit(100, b2.DECIMAL64, 12, 2), b2.lit(0, b2.DECIMAL64, 12, 2)
    run(b2, [cx.isin([1, 5, 7, 19]), cx.isin([3, b2.lit(None, b2.INT32)]), cx.isin([]), cs.isin(["MAIL", "SHIP"]),
             # q12 shape: sum(case when o_orderpriority in (...) then 1 else 0 end)
             b2.case_when([(cs.isin(["MAIL", "SHIP"]), one)], zero),
             b2.case_when([(cx < 5, cd), (cx < 10, cd * 2)], zero),
             b2.case_when([(cx == 3, cd)]),                         # no ELSE -> NULL
             b2.case_when([(cs.like("%special%"), cx), (cs.is_null(), b2.lit(-1, b2.INT32))], b2.lit(0, b2.INT32))], [x, d, s])


def test_filter_with_string_predicate_and_pruning(b2):
    """q3's customer side: Filter(c_mktsegment = 'BUILDING') under a pruning project keeps c_custkey only"""
    rng = np.random.default_rng(8)
    n = 20000
    key = O.OCol(np.arange(n, dtype=np.int64), np.ones(n, bool), (O.INT64, 0, 0))
    seg = str_col(rng, n, null_frac=0.05)
    t = G.to_b2_table(b2, [key, seg])
    pred = G.b2_expr_col(b2, 1, seg) == "BUILDING"
    keep = O.eval_expr(pred.sexpr, [key, seg])
    out = b2.filter_select(b2.Program([pred]), t, [0])
    assert out.num_columns == 1
    G.assert_col_equal(out.column(0), O.filter_cols([key], keep)[0])
    both = b2.filter_select(b2.Program([pred]), t, [1, 0])
    exp = O.filter_cols([seg, key], keep)
    G.assert_col_equal(both.column(0), exp[0]); G.assert_col_equal(both.column(1), exp[1])
    assert b2.filter_count(b2.Program([pred]), t) == len(exp[0])


def test_project_bound_reference_passes_through(b2):
    """GpuBoundReference outputs are the input columns themselves (no copy), strings included"""
    rng = np.random.default_rng(9)
    n = 3000
    a, s = G.gen_column(rng, (O.INT64, 0, 0), n), str_col(rng, n)
    t = G.to_b2_table(b2, [a, s])
    ca, cs = G.b2_expr_col(b2, 0, a), G.b2_expr_col(b2, 1, s)
    one = b2.lit(1, b2.INT64)
    out = b2.project(b2.Program([cs, ca + one, ca]), t)
    G.assert_col_equal(out.column(0), s)
    G.assert_col_equal(out.column(1), O.eval_expr((ca + one).sexpr, [a, s]))
    G.assert_col_equal(out.column(2), a)
    assert out.column(0).info().data == t.column(1).info().data and out.column(2).info().data == t.column(0).info().data
