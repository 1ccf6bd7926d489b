"""oracle/spark_hash.py — CPU restatement of Spark's Murmur3 hashing and hash partitioning.

TEST INFRASTRUCTURE ONLY (see oracle/spark_cpu.py header).

Follows: GpuMurmur3Hash.compute (sql-plugin/.../rapids/HashFunctions.scala:196-209: chained over the
key columns, seed 42 for partitioning), GpuHashPartitioningBase.scala:36-54, 82-100 (pmod, then
Table.partition = stable split), shims/HashUtils.scala:53-77 (-0.0 normalised to 0.0), and Spark's
org.apache.spark.unsafe.hash.Murmur3_x86_32 / HashExpression (external published algorithm; the
native implementation lives in spark-rapids-jni Hash.murmurHash32, absent from /root/reference).

PINNED by known answers (tests/test_oracle_golden.py): Spark `hash(1)` = -559580957,
`hash(1L)` = -1712319331, plus hand-derived vectors for strings and null chaining.
"""
import struct

import numpy as np

from . import spark_cpu as O

M32 = 0xFFFFFFFF


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & M32


def _mix_k1(k1):
    k1 = (k1 * 0xCC9E2D51) & M32
    k1 = _rotl(k1, 15)
    return (k1 * 0x1B873593) & M32


def _mix_h1(h1, k1):
    h1 ^= k1
    h1 = _rotl(h1, 13)
    return (h1 * 5 + 0xE6546B64) & M32


def _fmix(h1, length):
    h1 ^= length & M32
    h1 ^= h1 >> 16
    h1 = (h1 * 0x85EBCA6B) & M32
    h1 ^= h1 >> 13
    h1 = (h1 * 0xC2B2AE35) & M32
    h1 ^= h1 >> 16
    return h1


def hash_int(v, seed):
    return _fmix(_mix_h1(seed & M32, _mix_k1(v & M32)), 4)


def hash_long(v, seed):
    v &= 0xFFFFFFFFFFFFFFFF
    h1 = _mix_h1(seed & M32, _mix_k1(v & M32))
    h1 = _mix_h1(h1, _mix_k1(v >> 32))
    return _fmix(h1, 8)


def hash_bytes(b, seed):
    """Murmur3_x86_32.hashUnsafeBytes: 4-byte little-endian words, then every trailing byte (signed) as a block"""
    h1 = seed & M32
    aligned = len(b) & ~3
    for i in range(0, aligned, 4):
        h1 = _mix_h1(h1, _mix_k1(struct.unpack_from("<I", b, i)[0]))
    for i in range(aligned, len(b)):
        byte = b[i] - 256 if b[i] >= 128 else b[i]
        h1 = _mix_h1(h1, _mix_k1(byte & M32))
    return _fmix(h1, len(b))


def _to_signed(h):
    return h - (1 << 32) if h >= 1 << 31 else h


def hash_value(v, typ, seed):
    dt = typ[0]
    if dt == O.BOOL8:
        return hash_int(1 if v else 0, seed)
    if dt in (O.INT8, O.INT16, O.INT32, O.DATE32):
        return hash_int(int(v), seed)
    if dt in (O.INT64, O.TIMESTAMP_US):
        return hash_long(int(v), seed)
    if dt == O.FLOAT32:
        f = np.float32(v)
        bits = 0x7FC00000 if f != f else (0 if f == 0 else int(np.array([f], np.float32).view(np.uint32)[0]))
        return hash_int(bits, seed)
    if dt == O.FLOAT64:
        f = float(v)
        bits = 0x7FF8000000000000 if f != f else (0 if f == 0 else struct.unpack("<Q", struct.pack("<d", f))[0])
        return hash_long(bits, seed)
    if O.is_decimal(dt):
        if typ[1] <= 18 and dt != O.DECIMAL128:
            return hash_long(int(v), seed)
        v = int(v)
        nbytes = (v.bit_length() + 8) // 8 if v >= 0 else ((v + 1).bit_length() + 8) // 8  # BigInteger.toByteArray
        return hash_bytes(v.to_bytes(nbytes, "big", signed=True), seed)
    if dt == O.STRING:
        return hash_bytes(v if isinstance(v, bytes) else v.encode(), seed)
    raise NotImplementedError(dt)


def murmur3_rows(cols, seed=42):
    """-> int32 array: hash chained across columns, NULL leaves the running hash unchanged"""
    n = len(cols[0]) if cols else 0
    out = np.zeros(n, dtype=np.int32)
    for i in range(n):
        h = seed & M32
        for c in cols:
            if c.valid[i]:
                h = hash_value(c.values[i], c.typ, h)
        out[i] = _to_signed(h)
    return out


def partition_ids(cols, num_partitions, seed=42):
    h = murmur3_rows(cols, seed).astype(np.int64)
    return (((h % num_partitions) + num_partitions) % num_partitions).astype(np.int32)


def hash_partition(cols, key_idx, num_partitions, seed=42):
    """-> (reordered cols, offsets): stable split so each partition is contiguous"""
    pids = partition_ids([cols[k] for k in key_idx], num_partitions, seed)
    order = np.argsort(pids, kind="stable")
    counts = np.bincount(pids, minlength=num_partitions)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return [O.OCol(c.values[order], c.valid[order], c.typ) for c in cols], [int(x) for x in offsets]
