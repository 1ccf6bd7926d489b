"""oracle/tpch.py — synthetic TPC-H-shaped inputs and the CPU restatement of the
benchmark queries on them.  TEST / BENCH-BASELINE INFRASTRUCTURE ONLY (see spark_cpu.py header).

The data spec mirrors the reference's deterministic generator usage (datagen/README.md:40-62;
datagen/src/main/scala/.../bigDataGen.scala): fixed seed, uniform columns; TPC-H value domains.
q6_cpu is the "vanilla CPU plan" stand-in for `bench.py --impl reference` / cpu_baseline when no
JVM/Spark exists on the box: the same scan -> filter -> project -> aggregate pipeline the CPU Spark
plan runs (FileScan parquet -> Filter -> Project -> HashAggregate), executed by pyarrow's
multi-threaded reader and compute kernels on all host cores.  It is NOT Spark."""
import io
import os

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
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


def q6_cpu(parquet_bytes, threads=None):
    """scan -> filter -> project -> sum on the host cores; returns the unscaled dec(25,4) revenue (python int) or None"""
    if threads:
        pa.set_cpu_count(threads)
        pa.set_io_thread_count(threads)
    tbl = pq.read_table(pa.BufferReader(pa.py_buffer(parquet_bytes)), columns=["l_shipdate", "l_discount", "l_quantity", "l_extendedprice"],
                        use_threads=True)
    ship = tbl.column("l_shipdate").cast(pa.int32())
    disc, qty, price = tbl.column("l_discount"), tbl.column("l_quantity"), tbl.column("l_extendedprice")
    import decimal
    d = lambda s: pa.scalar(decimal.Decimal(s), type=pa.decimal128(12, 2))  # noqa: E731
    mask = pc.and_(pc.and_(pc.greater_equal(ship, Q6_DATE_LO), pc.less(ship, Q6_DATE_HI)),
                   pc.and_(pc.and_(pc.greater_equal(disc, d("0.05")), pc.less_equal(disc, d("0.07"))), pc.less(qty, d("24.00"))))
    f = tbl.filter(mask)
    rev = pc.multiply(f.column("l_extendedprice"), f.column("l_discount"))   # decimal(25,4), exact
    total = pc.sum(rev).as_py()
    return None if total is None else int(total.scaleb(4))


def q6_numpy_chunks(chunks):
    total, any_rows = 0, False
    for cols in chunks:
        r = q6_numpy(cols)
        if r is not None:
            total += r
            any_rows = True
    return total if any_rows else None


def q6_numpy(cols):
    """exact integer restatement over the raw numpy columns (checks both q6_cpu and the CUDA path)"""
    m = ((cols["l_shipdate"] >= Q6_DATE_LO) & (cols["l_shipdate"] < Q6_DATE_HI) & (cols["l_discount"] >= 5) & (cols["l_discount"] <= 7)
         & (cols["l_quantity"] < 2400))
    if not m.any():
        return None
    return int((cols["l_extendedprice"][m].astype(object) * cols["l_discount"][m].astype(object)).sum())
