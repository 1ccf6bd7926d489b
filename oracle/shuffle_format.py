"""oracle/shuffle_format.py — independent reader/writer of the B2T1 shuffle wire format (csrc/serialize.cu).
TEST INFRASTRUCTURE ONLY.  Pins the byte layout: tests serialise on the GPU and parse here, and build buffers here for the
GPU's coalesce-on-read.  The reference's own format (JCudfSerialization / Kudo, used by GpuColumnarBatchSerializer.scala:169-320)
lives in the un-vendored cudf-java / spark-rapids-jni dependency, so byte compatibility with it is NOT established (parity
unpinned for that row); what is pinned is that this format round-trips every column type and that coalesce-on-read
(GpuShuffleCoalesceExec.scala:371-475 semantics: concatenate in arrival order) equals row-wise concatenation."""
import struct

import numpy as np

from . import spark_cpu as O

MAGIC = 0x31543242
_W = {O.BOOL8: 1, O.INT8: 1, O.INT16: 2, O.INT32: 4, O.INT64: 8, O.FLOAT32: 4, O.FLOAT64: 8, O.DATE32: 4, O.TIMESTAMP_US: 8, O.DECIMAL32: 4,
      O.DECIMAL64: 8, O.DECIMAL128: 16}


def _pad64(n):
    return (n + 63) & ~63


def parse(buf):
    """bytes -> list of OCol (precision is not on the wire: decimals come back with precision 0)"""
    buf = bytes(buf)
    magic, version, ncols, rows, total = struct.unpack_from("<IHHqq", buf, 0)
    assert magic == MAGIC and version == 1 and total <= len(buf), (hex(magic), version, total, len(buf))
    descs = [struct.unpack_from("<iiqqqq", buf, 24 + 40 * i) for i in range(ncols)]
    off = _pad64(24 + 40 * ncols)
    cols = []
    for dtype, scale, nulls, vbytes, obytes, dbytes in descs:
        if vbytes:
            assert vbytes == (rows + 7) // 8
            valid = np.unpackbits(np.frombuffer(buf, np.uint8, vbytes, off), bitorder="little")[:rows].astype(bool)
        else:
            valid = np.ones(rows, bool)
        assert int((~valid).sum()) == nulls
        off += _pad64(vbytes)
        if dtype == O.STRING:
            offsets = np.frombuffer(buf, np.int32, rows + 1, off)
            assert obytes == 4 * (rows + 1) and (rows == 0 or offsets[0] == 0) and offsets[-1] == dbytes
            off += _pad64(obytes)
            chars = buf[off:off + dbytes]
            vals = np.array([chars[offsets[i]:offsets[i + 1]] for i in range(rows)], dtype=object)
        else:
            assert obytes == 0 and dbytes == rows * _W[dtype]
            if dtype == O.DECIMAL128:
                raw = np.frombuffer(buf, np.uint64, rows * 2, off).reshape(rows, 2)
                vals = np.array([((int(hi) << 64) | int(lo)) - ((1 << 128) if int(hi) >> 63 else 0) for lo, hi in raw], dtype=object)
            elif O.is_decimal(dtype):
                vals = np.array([int(v) for v in np.frombuffer(buf, np.int32 if dtype == O.DECIMAL32 else np.int64, rows, off)], dtype=object)
            else:
                vals = np.frombuffer(buf, O._NP[dtype], rows, off).copy()
        off += _pad64(dbytes)
        cols.append(O.OCol(vals, valid, (dtype, 0, scale)))
    assert off == total
    return cols


def build(cols):
    """list of OCol -> bytes in the wire format (what a peer executor would have written)"""
    rows = len(cols[0]) if cols else 0
    descs, payload = [], b""
    for c in cols:
        dtype, _, scale = c.typ
        nulls = int((~c.valid).sum())
        v = np.packbits(c.valid, bitorder="little").tobytes() if nulls else b""
        if dtype == O.STRING:
            enc = [bytes(x) if ok else b"" for x, ok in zip(c.values, c.valid)]
            offsets = np.zeros(rows + 1, np.int32)
            if rows:
                offsets[1:] = np.cumsum([len(e) for e in enc])
            o, d = offsets.tobytes(), b"".join(enc)
        elif dtype == O.DECIMAL128:
            o = b""
            d = b"".join(((int(x) if ok else 0) & ((1 << 128) - 1)).to_bytes(16, "little") for x, ok in zip(c.values, c.valid))
        elif O.is_decimal(dtype):
            o, d = b"", np.array([int(x) if ok else 0 for x, ok in zip(c.values, c.valid)], dtype=np.int32 if dtype == O.DECIMAL32 else np.int64).tobytes()
        else:
            o, d = b"", np.where(c.valid, c.values, 0).astype(O._NP[dtype]).tobytes()
        descs.append(struct.pack("<iiqqqq", dtype, scale, nulls, len(v), len(o), len(d)))
        for part in (v, o, d):
            payload += part + b"\0" * (_pad64(len(part)) - len(part))
    head_len = _pad64(24 + 40 * len(cols))
    total = head_len + len(payload)
    head = struct.pack("<IHHqq", MAGIC, 1, len(cols), rows, total) + b"".join(descs)
    return head + b"\0" * (head_len - len(head)) + payload
