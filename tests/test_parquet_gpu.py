"""a10 parity: Parquet -> device decode vs (1) the parquet-testing golden corpus and (2) pyarrow on
files generated on the fly (page sizes, dictionary on/off, nulls, snappy, v1/v2 pages, row groups)."""
import base64
import io
import json
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from oracle import spark_cpu as O
from oracle import spark_parquet as P
from tests import datagen as G

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "parquet_testing.json")))


def _norm_got(col):
    dt = col.dtype
    out = []
    for v in col.to_pylist():
        if v is None:
            out.append(None)
        elif dt in (5, 6):
            out.append("nan" if v != v else float(v))
        elif dt == 12:
            out.append(v)
        else:
            out.append(v)
    return out


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_parquet_testing_golden(b2, name):
    entry = GOLDEN[name]
    raw = base64.b64decode(entry["b64"])
    t = b2.parquet_decode(raw, entry["columns"])
    for i, c in enumerate(entry["columns"]):
        exp = entry["expect"][c]["values"]
        col = t.column(i)
        if col.dtype == b2.STRING:
            vals, valid = col.to_numpy()
            got = [None if not ok else v.decode("latin-1") for v, ok in zip(vals, valid)]
        else:
            got = _norm_got(col)
        assert got == exp, (name, c)


def _write(table, **kw):
    sink = io.BytesIO()
    pq.write_table(table, sink, **kw)
    return sink.getvalue()


def _tpch_like(n, rng, nulls=False):
    def maybe(arr):
        if not nulls:
            return arr
        mask = rng.random(n) < 0.15
        return pa.array(arr.to_pylist(), type=arr.type, mask=mask)
    import decimal
    qty = pa.array([decimal.Decimal(int(v)) for v in rng.integers(1, 51, n)], type=pa.decimal128(12, 2))
    price = pa.array([decimal.Decimal(int(v)) / 100 for v in rng.integers(90000, 10494951, n)], type=pa.decimal128(12, 2))
    disc = pa.array([decimal.Decimal(int(v)) / 100 for v in rng.integers(0, 11, n)], type=pa.decimal128(12, 2))
    ship = pa.array(rng.integers(8036, 10561, n).astype(np.int32), type=pa.int32()).cast(pa.date32())
    key = pa.array(rng.integers(0, 2**40, n), type=pa.int64())
    flag = pa.array(np.array(["A", "N", "R"])[rng.integers(0, 3, n)])
    comment = pa.array(["comment %d %s" % (i, "x" * int(k)) for i, k in enumerate(rng.integers(0, 30, n))])
    dbl = pa.array(rng.standard_normal(n))
    i32 = pa.array(rng.integers(-2**31, 2**31 - 1, n).astype(np.int32))
    ts = pa.array(rng.integers(0, 2**50, n), type=pa.int64()).cast(pa.timestamp("us"))
    big = pa.array([decimal.Decimal(int(v)) * 10**12 + 7 for v in rng.integers(-10**9, 10**9, n)], type=pa.decimal128(30, 4))
    bools = pa.array(rng.integers(0, 2, n).astype(bool))
    cols = {"l_quantity": qty, "l_extendedprice": price, "l_discount": disc, "l_shipdate": ship, "l_orderkey": key, "l_returnflag": flag,
            "l_comment": comment, "dbl": dbl, "i32": i32, "ts": ts, "big": big, "flag": bools}
    return pa.table({k: maybe(v) for k, v in cols.items()})


@pytest.mark.parametrize("nulls", [False, True])
@pytest.mark.parametrize("opts", [
    dict(compression="snappy", use_dictionary=True),
    dict(compression="none", use_dictionary=False),
    dict(compression="snappy", use_dictionary=True, data_page_size=4096),
    dict(compression="snappy", use_dictionary=False, data_page_version="2.0", data_page_size=16384),
    dict(compression="snappy", use_dictionary=True, row_group_size=7000, store_decimal_as_integer=True),
])
def test_decode_vs_pyarrow(b2, opts, nulls):
    rng = np.random.default_rng(len(str(opts)) + nulls)
    tbl = _tpch_like(30000, rng, nulls)
    raw = _write(tbl, **opts)
    cols = tbl.column_names
    t = b2.parquet_decode(raw, cols)
    exp = P.read_parquet(raw, cols)
    assert t.num_rows == 30000
    for i in range(len(cols)):
        got = t.column(i)
        assert got.dtype == exp[i].typ[0], (cols[i], got.dtype, exp[i].typ)
        G.assert_col_equal(got, exp[i])


def test_column_selection_order_and_case(b2):
    rng = np.random.default_rng(5)
    raw = _write(_tpch_like(2000, rng), compression="snappy")
    t = b2.parquet_decode(raw, ["l_shipdate", "L_QUANTITY", "l_orderkey"])
    exp = P.read_parquet(raw, ["l_shipdate", "l_quantity", "l_orderkey"])
    for i in range(3):
        G.assert_col_equal(t.column(i), exp[i])
    with pytest.raises(b2.B2Error):
        b2.parquet_decode(raw, ["nope"])


def test_empty_and_single_row(b2):
    for n in (0, 1):
        tbl = pa.table({"a": pa.array(np.arange(n, dtype=np.int64)), "s": pa.array(["x"] * n)})
        raw = _write(tbl, compression="snappy")
        t = b2.parquet_decode(raw, ["a", "s"])
        assert t.num_rows == n
        assert t.to_rows() == [(0, "x")][:n]


def test_unsupported_is_loud(b2):
    tbl = pa.table({"a": pa.array(np.arange(100, dtype=np.int64))})
    raw = _write(tbl, compression="zstd")
    with pytest.raises(b2.B2Error):
        b2.parquet_decode(raw, ["a"])
    with pytest.raises(b2.B2Error):
        b2.parquet_decode(b"PAR1garbagePAR1", ["a"])
    nested = pa.table({"l": pa.array([[1, 2], [3]])})
    with pytest.raises(b2.B2Error):
        b2.parquet_decode(_write(nested), ["l"])


def test_q6_from_parquet_end_to_end(b2):
    """config 1 shape: Parquet bytes -> decode -> fused filter+project+sum, vs oracle over pyarrow's read"""
    rng = np.random.default_rng(6)
    raw = _write(_tpch_like(50000, rng), compression="snappy", use_dictionary=True)
    cols = ["l_shipdate", "l_discount", "l_quantity", "l_extendedprice"]
    t = b2.parquet_decode(raw, cols)
    oc = P.read_parquet(raw, cols)
    c = [b2.col(0, b2.DATE32, nullable=False)] + [b2.col(i, b2.DECIMAL64, 12, 2, nullable=False) for i in (1, 2, 3)]
    pred = ((c[0] >= b2.lit(8766, b2.DATE32)) & (c[0] < b2.lit(9131, b2.DATE32)) & (c[1] >= b2.lit(5, b2.DECIMAL64, 3, 2))
            & (c[1] <= b2.lit(7, b2.DECIMAL64, 3, 2)) & (c[2] < b2.lit(2400, b2.DECIMAL64, 12, 2)))
    rev = c[3] * c[1]
    spec = [(b2.AGG_SUM, 0, b2.DECIMAL128, 4, 35)]
    got = b2.scan_aggregate(b2.Program([pred, rev]), True, t, [], spec).to_rows()
    keep = O.eval_expr(pred.sexpr, oc)
    exp = O.rows_of(O.reduce_cols(O.filter_cols([O.eval_expr(rev.sexpr, oc)], keep), spec))
    assert got == exp


def test_row_group_splits(b2):
    """a task decodes only the row groups of its split (GpuParquetScan.scala filterBlocks); the union is the file"""
    rng = np.random.default_rng(11)
    tbl = _tpch_like(25000, rng, True)
    raw = _write(tbl, compression="snappy", row_group_size=6000)
    cols = ["l_orderkey", "l_comment", "l_extendedprice"]
    n = b2.parquet_num_row_groups(raw)
    assert n == 5
    exp = P.read_parquet(raw, cols)
    parts = [b2.parquet_decode_row_groups(raw, cols, g, min(n, g + 2)) for g in range(0, n, 2)]
    assert [p.num_rows for p in parts] == [12000, 12000, 1000]
    whole = b2.concat(parts)
    for i in range(len(cols)):
        G.assert_col_equal(whole.column(i), exp[i])
    assert b2.parquet_decode_row_groups(raw, cols, 5, 9).num_rows == 0


@pytest.mark.parametrize("page", [8 << 10, 200 << 10, 2 << 20])
def test_snappy_patterns(b2, page):
    """LZ77 shapes the decoder treats differently: long literals, in-window chains (small int64 PLAIN), runs
    (offset < length), far back-references (beyond the 4 KB ring history), mixed, across page sizes that take
    the warp-pair kernel and the CTA-wide kernel"""
    rng = np.random.default_rng(page)
    n = 120_000
    small = rng.integers(0, 1 << 20, n)                               # 3 data bytes + 5 zero bytes per value
    runs = np.repeat(rng.integers(0, 1 << 40, n // 500 + 1), 500)[:n]  # long runs: overlapping copies
    period = np.tile(rng.integers(0, 1 << 62, 1500), n // 1500 + 1)[:n]  # 12 KB period: far back-references
    noise = rng.integers(-2**62, 2**62, n)                             # incompressible: long literals
    mixed = np.where((np.arange(n) // 3000) % 2 == 0, noise, small)
    tbl = pa.table({"small": pa.array(small), "runs": pa.array(runs), "period": pa.array(period), "noise": pa.array(noise), "mixed": pa.array(mixed)})
    raw = _write(tbl, compression="snappy", use_dictionary=False, data_page_size=page)
    cols = tbl.column_names
    t = b2.parquet_decode(raw, cols)
    for i, c in enumerate(cols):
        assert np.array_equal(t.column(i).to_numpy()[0], tbl.column(c).to_numpy()), c


@pytest.mark.parametrize("nulls", [False, True])
@pytest.mark.parametrize("opts", [dict(compression="none"), dict(compression="snappy", data_page_size=8192), dict(compression="snappy", data_page_version="2.0")])
def test_delta_binary_packed_generated(b2, opts, nulls):
    """DELTA_BINARY_PACKED INT32 / INT64 pages written by pyarrow: sorted keys (small deltas), random full-range values
    (64-bit wide deltas, wrapping sums), constant runs (width 0), dates, across page boundaries and NULLs"""
    rng = np.random.default_rng(3 + nulls)
    n = 70_001
    cols = {
        "sorted64": np.cumsum(rng.integers(0, 1000, n)).astype(np.int64),
        "rand64": rng.integers(-2**63, 2**63 - 1, n).astype(np.int64),
        "rand32": rng.integers(-2**31, 2**31 - 1, n).astype(np.int32),
        "const32": np.full(n, -7, dtype=np.int32),
        "steps": (np.arange(n) // 1000 * 12345 - 5_000_000).astype(np.int64),
    }
    arrays = {}
    for k, v in cols.items():
        mask = (rng.random(n) < 0.2) if nulls else None
        arrays[k] = pa.array(v, mask=mask)
    arrays["day"] = pa.array((8000 + np.arange(n) % 3000).astype(np.int32), type=pa.int32()).cast(pa.date32())
    tbl = pa.table(arrays)
    sink = io.BytesIO()
    pq.write_table(tbl, sink, use_dictionary=False, column_encoding={k: "DELTA_BINARY_PACKED" for k in tbl.column_names}, **opts)
    raw = sink.getvalue()
    names = tbl.column_names
    t = b2.parquet_decode(raw, names)
    exp = P.read_parquet(raw, names)
    for i in range(len(names)):
        G.assert_col_equal(t.column(i), exp[i])


def test_parquet_chunked_reader_row_group_chunks(b2):
    """ParquetChunkedReader: the reference's 10-row-group fixture (tests/src/test/resources/file-splits.parquet) read under
    byte limits -> same rows in the same order as one whole-buffer decode, chunk boundaries on row groups"""
    entry = GOLDEN["spark_rapids_tests/file-splits.parquet"]
    raw = base64.b64decode(entry["b64"])
    cols = entry["columns"]
    whole = b2.parquet_decode(raw, cols).to_rows()
    assert b2.parquet_num_row_groups(raw) == 10
    for limit, min_chunks in [(0, 1), (1, 10), (5_000, 2), (10**9, 1)]:
        rd = b2.ParquetChunkedReader(raw, cols, limit)
        rows, nchunks = [], 0
        for t in rd:
            rows += t.to_rows(); nchunks += 1
        rd.close()
        assert rows == whole, limit
        assert nchunks >= min_chunks and (limit != 1 or nchunks == 10) and (limit not in (0, 10**9) or nchunks == 1), (limit, nchunks)


def test_parquet_corrupt_page_header_is_rejected(b2):
    """untrusted input: negative sizes / truncated values in a page header must raise, not hang or read out of bounds"""
    rng = np.random.default_rng(12)
    tbl = pa.table({"a": pa.array(rng.integers(0, 1000, 5000), type=pa.int64()), "s": pa.array(["x%d" % i for i in range(5000)])})
    raw = bytearray(_write(tbl, compression="NONE", use_dictionary=False))
    good = b2.parquet_decode(bytes(raw), ["a", "s"])
    assert good.num_rows == 5000
    # truncate the file body while keeping the footer: pages then run past their chunk / promise more values than they hold
    md = pq.ParquetFile(io.BytesIO(bytes(raw))).metadata
    off = md.row_group(0).column(0).data_page_offset
    bad = bytearray(raw)
    bad[off + 2: off + 6] = b"\xff\xff\xff\x0f"   # uncompressed/compressed size varints become huge / negative
    with pytest.raises(b2.B2Error):
        b2.parquet_decode(bytes(bad), ["a", "s"])
