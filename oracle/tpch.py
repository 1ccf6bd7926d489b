"""oracle/tpch.py — CPU restatement of the benchmark query on the synthetic TPC-H inputs of benchdata/tpch.py.
TEST / BENCH-BASELINE INFRASTRUCTURE ONLY (see spark_cpu.py header).

q6_cpu is the "vanilla CPU plan" stand-in for `bench.py --impl reference` / cpu_baseline when no JVM/Spark exists on
the box: the same scan -> filter -> project -> aggregate pipeline the CPU Spark plan runs (FileScan parquet ->
Filter -> Project -> HashAggregate), executed by pyarrow's multi-threaded reader and compute kernels on all host
cores.  It is NOT Spark.  q6_numpy is the exact integer restatement over the raw columns that pins both."""
import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.parquet as pq

from benchdata.tpch import (DATE_1992_01_02, DATE_1998_12_01, LINEITEM_SF1_ROWS, Q6_DATE_HI, Q6_DATE_LO, SF_ROWS,  # noqa: F401
                            lineitem_q6_chunks, lineitem_q6_columns, lineitem_q6_parquet)


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
