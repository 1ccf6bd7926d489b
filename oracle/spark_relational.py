"""oracle/spark_relational.py — CPU restatement of join / sort / gather / row-format semantics.

TEST INFRASTRUCTURE ONLY (see oracle/spark_cpu.py header).  "parity unpinned": the reference holds
no golden vectors for these (its tests are differential against live Spark); the
rules restated here are the ones its Scala call sites document:

  joins   GpuHashJoin.scala:256-302 (makeLeftOuter/makeSemi/makeAnti), :309-600, :602-640 (NULL keys
          never match unless compareNullsEqual), JoinGatherer.scala:585-599 (OOB index -> NULL row)
  sort    SortUtils.scala:38-43 (asc/desc x nulls first/last), docs/compatibility.md:18-41, 71-84
          (NaN greatest, -0.0 == 0.0)
  rows    shims/CudfUnsafeRowBase.scala:80-90, 234-246 (JCUDF fixed-width row layout)
"""
import functools
import struct

import numpy as np

from . import spark_cpu as O

INT32_MIN = -2**31


def _join_key(cols, i, nulls_equal):
    k = []
    for c in cols:
        if not c.valid[i]:
            if not nulls_equal:
                return None
            k.append(None)
        else:
            k.append(O._key_of(c, i))
    return tuple(k)


def hash_join(build_keys, probe_keys, kind, nulls_equal=False):
    """-> (left_map, right_map|None) as python lists; stream/probe side is 'left'.
    kind: 0 inner, 1 left outer, 2 left semi, 3 left anti, 4 full outer (left outer rows, then the unmatched build
    rows in build order with left = INT32_MIN; GpuHashJoin.scala full-join gather maps)."""
    table = {}
    nb = len(build_keys[0])
    for i in range(nb):
        k = _join_key(build_keys, i, nulls_equal)
        if k is not None:
            table.setdefault(k, []).append(i)
    lm, rm = [], []
    for r in range(len(probe_keys[0])):
        k = _join_key(probe_keys, r, nulls_equal)
        m = table.get(k, []) if k is not None else []
        if kind == 0:
            for b in m:
                lm.append(r); rm.append(b)
        elif kind == 1:
            if m:
                for b in m:
                    lm.append(r); rm.append(b)
            else:
                lm.append(r); rm.append(INT32_MIN)
        elif kind == 2:
            if m:
                lm.append(r)
        else:
            if not m:
                lm.append(r)
    if kind == 4:
        lm, rm = hash_join(build_keys, probe_keys, 1, nulls_equal)
        hit = set(b for b in rm if b >= 0)
        for b in range(nb):
            if b not in hit:
                lm.append(INT32_MIN); rm.append(b)
        return lm, rm
    return lm, (rm if kind in (0, 1) else None)


def gather(cols, gmap, nullify_oob):
    out = []
    n = len(cols[0]) if cols else 0
    gmap = np.asarray(gmap, dtype=np.int64)
    ok = (gmap >= 0) & (gmap < n)
    safe = np.where(ok, gmap, 0)
    for c in cols:
        if n == 0:
            vals = np.zeros(len(gmap), dtype=c.values.dtype)
            valid = np.zeros(len(gmap), bool)
        else:
            vals = c.values[safe]
            valid = c.valid[safe] & ok
        out.append(O.OCol(vals, valid, c.typ))
    return out


def _cmp_vals(c, i, j):
    dt = c.typ[0]
    a, b = c.values[i], c.values[j]
    if dt in (O.FLOAT32, O.FLOAT64):
        a, b = float(a), float(b)
        an, bn = a != a, b != b
        if an or bn:
            return 0 if an and bn else (1 if an else -1)
    return -1 if a < b else (1 if a > b else 0)


def sort_order(cols, keys):
    """stable argsort; keys = [(column, ascending, nulls_first)]"""
    n = len(cols[0]) if cols else 0

    def cmp(i, j):
        for col, asc, nf in keys:
            c = cols[col]
            vi, vj = c.valid[i], c.valid[j]
            if not vi or not vj:
                if vi == vj:
                    continue
                r = -1 if not vi else 1  # null before valid
                return r if nf else -r
            r = _cmp_vals(c, i, j)
            if r:
                return r if asc else -r
        return 0
    return sorted(range(n), key=functools.cmp_to_key(cmp))


def take(cols, idx):
    idx = np.asarray(idx, dtype=np.int64)
    return [O.OCol(c.values[idx], c.valid[idx], c.typ) for c in cols]


# ---- JCUDF fixed-width rows
_W = {O.BOOL8: 1, O.INT8: 1, O.INT16: 2, O.INT32: 4, O.INT64: 8, O.FLOAT32: 4, O.FLOAT64: 8, O.DATE32: 4, O.TIMESTAMP_US: 8,
      O.DECIMAL32: 4, O.DECIMAL64: 8, O.DECIMAL128: 16}
_FMT = {O.BOOL8: "<b", O.INT8: "<b", O.INT16: "<h", O.INT32: "<i", O.INT64: "<q", O.FLOAT32: "<f", O.FLOAT64: "<d", O.DATE32: "<i",
        O.TIMESTAMP_US: "<q", O.DECIMAL32: "<i", O.DECIMAL64: "<q"}


def row_layout(dtypes):
    off, offs = 0, []
    for dt in dtypes:
        w = _W[dt]
        off = (off + w - 1) & -w
        offs.append(off)
        off += w
    validity_off = off
    row_bytes = (off + (len(dtypes) + 7) // 8 + 7) & ~7
    return offs, validity_off, row_bytes


def to_rows(cols):
    dtypes = [c.typ[0] for c in cols]
    offs, voff, rb = row_layout(dtypes)
    n = len(cols[0])
    out = np.zeros((n, rb), dtype=np.uint8)
    for r in range(n):
        row = bytearray(rb)
        for ci, c in enumerate(cols):
            dt = dtypes[ci]
            if dt == O.DECIMAL128:
                row[offs[ci]:offs[ci] + 16] = (int(c.values[r]) & ((1 << 128) - 1)).to_bytes(16, "little")
            else:
                v = c.values[r]
                struct.pack_into(_FMT[dt], row, offs[ci], float(v) if dt in (O.FLOAT32, O.FLOAT64) else int(v))
            if c.valid[r]:
                row[voff + ci // 8] |= 1 << (ci % 8)
        out[r] = np.frombuffer(bytes(row), dtype=np.uint8)
    return out
