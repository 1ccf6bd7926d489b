"""Seeded test data in both representations (library Table + oracle OCol list).  Follows the
reference's generators (integration_tests/src/main/python/data_gen.py:863-889 gen_df with
special-case injection: nulls, min/max, NaN, +-0.0)."""
import numpy as np

from oracle import spark_cpu as O

_RANGES = {O.INT8: (-2**7, 2**7 - 1), O.INT16: (-2**15, 2**15 - 1), O.INT32: (-2**31, 2**31 - 1),
           O.INT64: (-2**63, 2**63 - 1), O.DATE32: (-25567, 47482)}


def gen_column(rng, typ, n, null_frac=0.1, small=False, distinct=None):
    """-> oracle OCol.  typ = (dtype, precision, scale)"""
    dt = typ[0]
    valid = rng.random(n) >= null_frac if null_frac > 0 else np.ones(n, bool)
    if dt == O.BOOL8:
        vals = rng.integers(0, 2, n).astype(np.int8)
    elif dt in _RANGES:
        lo, hi = _RANGES[dt]
        if distinct:
            vals = rng.integers(0, distinct, n)
        elif small:
            vals = rng.integers(-1000, 1000, n)
        else:
            vals = rng.integers(lo, hi, n, endpoint=True, dtype=np.int64)
            special = np.array([lo, hi, 0, -1, 1], dtype=np.int64)
            k = min(n, 5)
            if n:
                vals[rng.choice(n, k, replace=False)] = special[:k]
        vals = vals.astype(O._NP[dt])
    elif dt in (O.FLOAT32, O.FLOAT64):
        vals = rng.standard_normal(n) * (10.0 if small else 1e6)
        special = np.array([np.nan, 0.0, -0.0, np.inf, -np.inf, 1.0])
        k = min(n, len(special))
        if n and not distinct:
            vals[rng.choice(n, k, replace=False)] = special[:k]
        if distinct:
            vals = rng.integers(0, distinct, n).astype(np.float64)
        vals = vals.astype(O._NP[dt])
    elif O.is_decimal(dt):
        p = typ[1]
        lim = 10 ** (min(p, 6) if small else p) - 1
        if distinct:
            vals = np.array([int(v) for v in rng.integers(0, distinct, n)], dtype=object)
        else:
            vals = np.array([int(rng.integers(-2**62, 2**62)) * int(rng.integers(1, 2**62)) % (2 * lim + 1) - lim for _ in range(n)], dtype=object)
            if n >= 3 and not small:
                idx = rng.choice(n, 3, replace=False)
                vals[idx[0]], vals[idx[1]], vals[idx[2]] = lim, -lim, 0
    elif dt == O.STRING:
        alphabet = [b"", b"a", b"A", b"N", b"R", b"F", b"O", b"abc", b"hello world", "été".encode(), b"zzzzzzzzzzzzzzzzzzzzzzzz"]
        k = distinct if distinct else len(alphabet)
        vals = np.array([alphabet[i % len(alphabet)] + (str(i // len(alphabet)).encode() if i >= len(alphabet) else b"")
                         for i in rng.integers(0, k, n)], dtype=object)
    else:
        raise NotImplementedError(dt)
    return O.OCol(vals, valid, typ)


def to_b2_column(m, oc):
    dt, _, scale = oc.typ
    valid = None if oc.valid.all() else oc.valid
    if dt == O.STRING:
        return m.Column.from_strings([v if ok else None for v, ok in zip(oc.values, oc.valid)])
    if O.is_decimal(dt):
        if dt == O.DECIMAL128:
            return m.Column.from_numpy(np.array([int(v) for v in oc.values], dtype=object), dtype=dt, valid=valid, scale=scale)
        return m.Column.from_numpy(np.array([int(v) for v in oc.values], dtype=np.int64), dtype=dt, valid=valid, scale=scale)
    return m.Column.from_numpy(oc.values, dtype=dt, valid=valid, scale=scale)


def to_b2_table(m, ocols):
    return m.Table.from_columns([to_b2_column(m, c) for c in ocols])


def b2_expr_col(m, i, oc):
    return m.col(i, oc.typ[0], oc.typ[1], oc.typ[2], nullable=not oc.valid.all())


def assert_col_equal(b2col, oc, approx=False):
    got = b2col.to_pylist()
    exp = oc.to_pylist()
    assert len(got) == len(exp), (len(got), len(exp))
    for i, (g, e) in enumerate(zip(got, exp)):
        if e is None or g is None:
            assert g is None and e is None, "row %d: got %r expected %r" % (i, g, e)
        elif isinstance(e, float):
            if e != e:
                assert g != g, "row %d: got %r expected NaN" % (i, g)
            elif approx:
                assert abs(g - e) <= 1e-9 * max(1.0, abs(e)), "row %d: got %r expected %r" % (i, g, e)
            else:
                assert g == e, "row %d: got %r expected %r" % (i, g, e)
        else:
            assert g == e, "row %d: got %r expected %r" % (i, g, e)


def norm_rows(rows):
    def k(v):
        if v is None:
            return (0, 0)
        if isinstance(v, float):
            if v != v:
                return (3, 0)
            return (1, 0.0 if v == 0 else v)
        if isinstance(v, str):
            return (2, v)
        return (1, v)
    return sorted([tuple(k(v) for v in r) for r in rows])
