"""benchdata/tpch.py — synthetic TPC-H-shaped inputs: deterministic lineitem columns for q6 and
the Parquet bytes a scan sees (snappy, dictionary on, ~128 MB row groups, decimals stored as INT64).

Neutral input generation only: no query semantics live here.  bench.py, the tests and the CPU restatement
(oracle/tpch.py) all read the same bytes from it.  The data spec mirrors the reference's deterministic generator
usage (datagen/README.md:40-62; datagen/src/main/scala/.../bigDataGen.scala): fixed seed, uniform columns,
TPC-H value domains."""
import io
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

LINEITEM_SF1_ROWS = 6001215 - 0  # TPC-H spec; SF10 = 59,986,052 per SURVEY §8
SF_ROWS = {1: 6001215, 10: 59986052, 100: 600037902}
DATE_1992_01_02, DATE_1998_12_01 = 8036, 10561
Q6_DATE_LO, Q6_DATE_HI = 8766, 9131   # 1994-01-01, 1995-01-01


def _dec128(int64_vals, precision, scale):
    n = len(int64_vals)
    buf = np.empty((n, 2), dtype=np.int64)
    buf[:, 0] = int64_vals
    buf[:, 1] = int64_vals >> 63
    return pa.Array.from_buffers(pa.decimal128(precision, scale), n, [None, pa.py_buffer(buf)])


def lineitem_q6_columns(rows, seed=42):
    """numpy columns: l_shipdate int32 days, l_discount / l_quantity / l_extendedprice int64 unscaled dec(12,2)"""
    rng = np.random.default_rng(seed)
    ship = rng.integers(DATE_1992_01_02, DATE_1998_12_01 + 1, rows, dtype=np.int32)
    disc = rng.integers(0, 11, rows, dtype=np.int64)
    qty = rng.integers(1, 51, rows, dtype=np.int64) * 100
    price = rng.integers(90000, 10494951, rows, dtype=np.int64)
    return {"l_shipdate": ship, "l_discount": disc, "l_quantity": qty, "l_extendedprice": price}


def lineitem_q6_chunks(rows, seed=42, row_group_rows=4_800_000):
    """the table as a sequence of row-group sized column dicts (chunk i is seeded seed + 1000003*i)"""
    done, i = 0, 0
    while done < rows or (rows == 0 and i == 0):
        n = min(row_group_rows, rows - done)
        yield lineitem_q6_columns(n, seed + 1000003 * i)
        done += n
        i += 1
        if rows == 0:
            break


def lineitem_q6_parquet(rows, seed=42, cache_dir=None, row_group_rows=4_800_000):
    """Parquet bytes as the reference's scan sees them: snappy, dictionary on, ~128 MB row groups,
    decimals stored as INT64."""
    path = None
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        path = os.path.join(cache_dir, "lineitem_q6_%d_seed%d.parquet" % (rows, seed))
        if os.path.exists(path):
            return np.fromfile(path, dtype=np.uint8)
    sink = io.BytesIO()
    writer = None
    for cols in lineitem_q6_chunks(rows, seed, row_group_rows):
        tbl = pa.table({
            "l_shipdate": pa.array(cols["l_shipdate"], type=pa.int32()).cast(pa.date32()),
            "l_discount": _dec128(cols["l_discount"], 12, 2),
            "l_quantity": _dec128(cols["l_quantity"], 12, 2),
            "l_extendedprice": _dec128(cols["l_extendedprice"], 12, 2),
        })
        if writer is None:
            writer = pq.ParquetWriter(sink, tbl.schema, compression="snappy", use_dictionary=True, store_decimal_as_integer=True)
        writer.write_table(tbl, row_group_size=row_group_rows)
    writer.close()
    raw = np.frombuffer(sink.getvalue(), dtype=np.uint8).copy()
    if path:
        raw.tofile(path + ".tmp%d" % os.getpid())
        os.replace(path + ".tmp%d" % os.getpid(), path)
    return raw


# ---- TPC-H q3 tables (customer 150k*SF, orders 1.5M*SF, lineitem 6,001,215*SF; dense i64 keys, uniform
# foreign keys, c_mktsegment 5 values, uniform dates, seed 42).  Tables are produced chunk by chunk so that a rank of an N-GPU
# run (and the chunked CPU check) can make exactly its share: chunk i of a table is seeded by (seed, table, i).
SEGMENTS = [b"AUTOMOBILE", b"BUILDING", b"FURNITURE", b"MACHINERY", b"HOUSEHOLD"]
Q3_SEGMENT = b"BUILDING"
Q3_DATE = 9204                      # 1995-03-15
DATE_1992_01_01, DATE_1998_08_02 = 8035, 10440
Q3_CHUNKS = {"customer": 8, "orders": 8, "lineitem": 16}
_TABLE_ID = {"customer": 1, "orders": 2, "lineitem": 3}


def q3_rows(sf):
    """row counts at scale factor sf (fractional sf allowed for tests)"""
    li = SF_ROWS.get(sf, int(round(6001215 * sf)))
    return {"customer": int(round(150_000 * sf)), "orders": int(round(1_500_000 * sf)), "lineitem": li}


def chunk_range(n, nchunks, i):
    return n * i // nchunks, n * (i + 1) // nchunks


def _segment_strings(codes):
    """codes -> (chars uint8, offsets int32) of the Arrow string column"""
    seg_len = np.array([len(s) for s in SEGMENTS], dtype=np.int32)
    width = int(seg_len.max())
    mat = np.zeros((len(SEGMENTS), width), dtype=np.uint8)
    for k, s in enumerate(SEGMENTS):
        mat[k, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    lens = seg_len[codes]
    offsets = np.zeros(len(codes) + 1, dtype=np.int32)
    np.cumsum(lens, out=offsets[1:])
    keep = np.arange(width, dtype=np.int32)[None, :] < lens[:, None]
    chars = mat[codes][keep]
    return np.ascontiguousarray(chars), offsets


def q3_chunk(table, sf, i, seed=42):
    """chunk i of a q3 table -> dict of numpy columns (c_mktsegment: (chars, offsets) plus the raw codes)"""
    rows = q3_rows(sf)
    lo, hi = chunk_range(rows[table], Q3_CHUNKS[table], i)
    n = hi - lo
    rng = np.random.default_rng([seed, _TABLE_ID[table], i])
    if table == "customer":
        codes = rng.integers(0, len(SEGMENTS), n, dtype=np.int8)
        chars, offsets = _segment_strings(codes)
        return {"c_custkey": np.arange(lo, hi, dtype=np.int64), "c_mktsegment": (chars, offsets), "c_mktsegment_code": codes}
    if table == "orders":
        return {"o_orderkey": np.arange(lo, hi, dtype=np.int64), "o_custkey": rng.integers(0, rows["customer"], n, dtype=np.int64),
                "o_orderdate": rng.integers(DATE_1992_01_01, DATE_1998_08_02 + 1, n, dtype=np.int32), "o_shippriority": np.zeros(n, dtype=np.int32)}
    return {"l_orderkey": rng.integers(0, rows["orders"], n, dtype=np.int64), "l_extendedprice": rng.integers(90000, 10494951, n, dtype=np.int64),
            "l_discount": rng.integers(0, 11, n, dtype=np.int64), "l_shipdate": rng.integers(DATE_1992_01_02, DATE_1998_12_01 + 1, n, dtype=np.int32)}


def q3_chunks_of_rank(table, rank, world):
    """the chunks a rank of a strong-scaled run owns (round robin, so every rank gets the same number)"""
    return [i for i in range(Q3_CHUNKS[table]) if i % world == rank]
