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


# ---- TPC-H q3 ----------------------------------------------------------------------------------------
#   select l_orderkey, sum(l_extendedprice * (1 - l_discount)) as revenue, o_orderdate, o_shippriority
#   from customer, orders, lineitem
#   where c_mktsegment = 'BUILDING' and c_custkey = o_custkey and l_orderkey = o_orderkey
#     and o_orderdate < date '1995-03-15' and l_shipdate > date '1995-03-15'
#   group by l_orderkey, o_orderdate, o_shippriority order by revenue desc, o_orderdate limit 10
# revenue is decimal(12,2) * decimal(13,2) = decimal(26,4), summed as decimal(36,4): unscaled value = price * (100 - disc).
def _member_index(sorted_keys):
    """-> f(keys) = (index into sorted_keys or -1): a direct-address table when the key range allows it, else binary search"""
    import numpy as np
    n = len(sorted_keys)
    if n and sorted_keys[0] >= 0 and sorted_keys[-1] < (1 << 31):
        lut = np.full(int(sorted_keys[-1]) + 2, -1, dtype=np.int32)
        lut[sorted_keys] = np.arange(n, dtype=np.int32)
        top = len(lut) - 1
        return lambda k: lut[np.where((k >= 0) & (k < top), k, top)]
    def f(k):
        if n == 0:
            return np.full(len(k), -1, dtype=np.int64)
        pos = np.searchsorted(sorted_keys, k)
        pos[pos >= n] = n - 1
        return np.where(sorted_keys[pos] == k, pos, -1)
    return f


def q3_numpy(customer_chunks, orders_chunks, lineitem_chunks, date=None, segment=None, limit=10, threads=1):
    """exact integer restatement over the raw numpy columns, one chunk at a time (never holds a whole table; the lineitem
    chunks may be callables so that `threads` workers generate and reduce them in parallel).
    -> list of (l_orderkey, revenue_unscaled_dec36_4, o_orderdate, o_shippriority), ordered by revenue desc, o_orderdate asc,
    (then l_orderkey asc to make ties deterministic for the comparison)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from benchdata import tpch as gen
    date = gen.Q3_DATE if date is None else date
    segment = gen.Q3_SEGMENT if segment is None else segment
    seg_code = gen.SEGMENTS.index(segment)
    # customer: keys of the wanted segment
    ckeys = []
    for c in customer_chunks:
        chars, offsets = c["c_mktsegment"]
        lens = np.diff(offsets)
        ok = lens == len(segment)
        if ok.any():  # compare the bytes of the candidates
            starts = offsets[:-1][ok]
            mat = chars[starts[:, None] + np.arange(len(segment))[None, :]]
            same = (mat == np.frombuffer(segment, dtype=np.uint8)[None, :]).all(axis=1)
            sel = np.flatnonzero(ok)[same]
            ckeys.append(c["c_custkey"][sel])
            assert (c["c_mktsegment_code"][sel] == seg_code).all()
    ckeys = np.sort(np.concatenate(ckeys)) if ckeys else np.zeros(0, np.int64)
    cust_index = _member_index(ckeys)
    okeys, odate, oprio = [], [], []
    for o in orders_chunks:
        m = (o["o_orderdate"] < date) & (cust_index(o["o_custkey"]) >= 0)
        okeys.append(o["o_orderkey"][m]); odate.append(o["o_orderdate"][m]); oprio.append(o["o_shippriority"][m])
    okeys = np.concatenate(okeys); odate = np.concatenate(odate); oprio = np.concatenate(oprio)
    order = np.argsort(okeys, kind="stable")
    okeys, odate, oprio = okeys[order], odate[order], oprio[order]
    assert len(okeys) == 0 or (np.diff(okeys) > 0).all(), "o_orderkey must be unique"
    if len(okeys) == 0:
        return []
    order_index = _member_index(okeys)

    def reduce_chunk(li):
        li = li() if callable(li) else li
        m = li["l_shipdate"] > date
        gid = order_index(li["l_orderkey"][m])           # index into the qualifying orders = group id, -1 = no match
        hit = gid >= 0
        rev = li["l_extendedprice"][m][hit] * (100 - li["l_discount"][m][hit])
        # per-group partial sums stay far below 2^53 (<= 7 lines x 1.05e9 per order), so float64 accumulation is exact
        return np.bincount(gid[hit], weights=rev.astype(np.float64), minlength=len(okeys)), np.bincount(gid[hit], minlength=len(okeys))

    sums = np.zeros(len(okeys), dtype=np.float64)
    cnts = np.zeros(len(okeys), dtype=np.int64)
    if threads > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            for s_, c_ in ex.map(reduce_chunk, lineitem_chunks):
                sums += s_; cnts += c_
    else:
        for li in lineitem_chunks:
            s_, c_ = reduce_chunk(li)
            sums += s_; cnts += c_
    assert sums.max(initial=0) < 2.0**53
    g = np.flatnonzero(cnts > 0)
    isums = sums[g].astype(np.int64)
    top = np.lexsort((okeys[g], odate[g], -isums))[:limit]
    return [(int(okeys[g[i]]), int(isums[i]), int(odate[g[i]]), int(oprio[g[i]])) for i in top]


def q3_expected(sf, seed=42, threads=8):
    """q3_numpy over every chunk of the synthetic tables at scale factor sf (chunks generated and reduced on `threads` host threads)"""
    from concurrent.futures import ThreadPoolExecutor
    from benchdata import tpch as gen
    with ThreadPoolExecutor(max_workers=threads) as ex:
        cust = list(ex.map(lambda i: gen.q3_chunk("customer", sf, i, seed), range(gen.Q3_CHUNKS["customer"])))
        orders = list(ex.map(lambda i: gen.q3_chunk("orders", sf, i, seed), range(gen.Q3_CHUNKS["orders"])))
    line = [(lambda i=i: gen.q3_chunk("lineitem", sf, i, seed)) for i in range(gen.Q3_CHUNKS["lineitem"])]
    return q3_numpy(cust, orders, line, threads=threads)


def q3_cpu(customer, orders, lineitem, threads=None, date=None, segment=None, limit=10):
    """The "vanilla CPU plan" stand-in for q3 when no JVM/Spark exists on the box: the same Filter -> HashJoin -> HashJoin ->
    HashAggregate -> TakeOrdered pipeline, executed by pyarrow's multi-threaded Acero engine on all host cores.  NOT Spark.
    Inputs are pyarrow Tables (built once, outside the timed region, like Spark's cached columnar input).
    Money columns are int64 unscaled decimals and revenue is computed in int64 (exact here; cheaper than Spark's Decimal
    arithmetic, i.e. favourable to the CPU side)."""
    import pyarrow as pa
    import pyarrow.compute as pc
    from benchdata import tpch as gen
    if threads:
        pa.set_cpu_count(threads)
    date = gen.Q3_DATE if date is None else date
    segment = gen.Q3_SEGMENT if segment is None else segment
    cust = customer.filter(pc.equal(customer["c_mktsegment"], pa.scalar(segment, type=pa.binary()))).select(["c_custkey"])
    ords = orders.filter(pc.less(orders["o_orderdate"], pa.scalar(date, type=pa.int32())))
    j1 = ords.join(cust, keys="o_custkey", right_keys="c_custkey", join_type="inner").select(["o_orderkey", "o_orderdate", "o_shippriority"])
    li = lineitem.filter(pc.greater(lineitem["l_shipdate"], pa.scalar(date, type=pa.int32()))).select(["l_orderkey", "l_extendedprice", "l_discount"])
    j2 = li.join(j1, keys="l_orderkey", right_keys="o_orderkey", join_type="inner")
    rev = pc.multiply(j2["l_extendedprice"], pc.subtract(pa.scalar(100, type=pa.int64()), j2["l_discount"]))
    j2 = j2.append_column("rev", rev)
    agg = j2.group_by(["l_orderkey", "o_orderdate", "o_shippriority"]).aggregate([("rev", "sum")])
    idx = pc.select_k_unstable(agg, k=limit, sort_keys=[("rev_sum", "descending"), ("o_orderdate", "ascending"), ("l_orderkey", "ascending")])
    top = agg.take(idx)
    return [(int(a), int(b), int(c), int(d)) for a, b, c, d in zip(top["l_orderkey"].to_pylist(), top["rev_sum"].to_pylist(), top["o_orderdate"].to_pylist(),
                                                                  top["o_shippriority"].to_pylist())]


def q3_arrow_tables(sf, seed=42, threads=8):
    """the synthetic q3 tables as pyarrow Tables (the CPU arm's cached input)"""
    import numpy as np
    import pyarrow as pa
    from concurrent.futures import ThreadPoolExecutor
    from benchdata import tpch as gen
    def table(name, cols):
        with ThreadPoolExecutor(max_workers=threads) as ex:
            chunks = list(ex.map(lambda i: gen.q3_chunk(name, sf, i, seed), range(gen.Q3_CHUNKS[name])))
        arrays = {}
        for c in cols:
            if c == "c_mktsegment":
                parts = []
                for ch in chunks:
                    chars, offsets = ch[c]
                    parts.append(pa.Array.from_buffers(pa.binary(), len(offsets) - 1, [None, pa.py_buffer(offsets), pa.py_buffer(chars)]))
                arrays[c] = pa.chunked_array(parts)
            else:
                arrays[c] = pa.chunked_array([pa.array(ch[c]) for ch in chunks])
        return pa.table(arrays)
    return (table("customer", ["c_custkey", "c_mktsegment"]), table("orders", ["o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"]),
            table("lineitem", ["l_orderkey", "l_extendedprice", "l_discount", "l_shipdate"]))
