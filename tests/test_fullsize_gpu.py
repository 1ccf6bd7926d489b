"""Parity at BASELINE.json's full sizes (SF10 lineitem = 59,986,052 rows per GPU) through size-independent
properties: the oracle's per-row Python loops cannot run at this size, so each operator is checked against
vectorised numpy restatements that are themselves pinned to the oracle on a small prefix.
 * filter      : count and ordered output equal numpy boolean indexing
 * group-by    : sums / counts per key equal np.bincount
 * sort        : permutation is valid, keys come out non-decreasing, stable on ties
 * murmur3 / hash_partition : ids equal a numpy Murmur3_x86_32 (pinned to oracle.spark_hash.hash_long), partitions are
                 contiguous in id order, offsets equal bincount, row multiset preserved (sum + xor checksums), stable
 * hash join   : |matches| and the checksum of joined payloads equal a numpy lookup join"""
import numpy as np
import pytest

from oracle import spark_hash as H

pytestmark = pytest.mark.gpu
N = 59_986_052
M32 = np.uint32(0xFFFFFFFF)


def _rotl(x, r):
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def _mix_k1(k):
    k = k * np.uint32(0xCC9E2D51)
    k = _rotl(k, 15)
    return k * np.uint32(0x1B873593)


def _mix_h1(h, k):
    h = h ^ k
    h = _rotl(h, 13)
    return h * np.uint32(5) + np.uint32(0xE6546B64)


def np_hash_long(v, seed=42):
    """Spark Murmur3_x86_32.hashLong over an int64 array -> int32 array (HashFunctions.scala:196-209 semantics)"""
    with np.errstate(over="ignore"):
        u = v.view(np.uint64)
        lo, hi = (u & np.uint64(0xFFFFFFFF)).astype(np.uint32), (u >> np.uint64(32)).astype(np.uint32)
        h = _mix_h1(np.full(len(v), seed, dtype=np.uint32), _mix_k1(lo))
        h = _mix_h1(h, _mix_k1(hi))
        h ^= np.uint32(8)
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x85EBCA6B)
        h ^= h >> np.uint32(13)
        h *= np.uint32(0xC2B2AE35)
        h ^= h >> np.uint32(16)
    return h.view(np.int32)


def test_numpy_murmur3_is_pinned_to_the_oracle():
    rng = np.random.default_rng(1)
    v = np.concatenate([rng.integers(-2**63, 2**63 - 1, 2000), np.array([0, 1, -1, 2**63 - 1, -2**63])]).astype(np.int64)
    exp = np.array([H._to_signed(H.hash_long(int(x), 42)) for x in v], dtype=np.int32)
    assert np.array_equal(np_hash_long(v), exp)
    assert np_hash_long(np.array([1], dtype=np.int64))[0] == -1712319331   # Spark: hash(1L)


@pytest.fixture(scope="module")
def big(b2):
    rng = np.random.default_rng(2024)
    key = rng.integers(0, 1_500_000, N).astype(np.int64)          # l_orderkey-like, ~40 rows per key
    val = rng.integers(90000, 10494951, N).astype(np.int64)       # l_extendedprice unscaled
    grp = rng.integers(0, 1000, N).astype(np.int32)
    t = b2.Table.from_columns([b2.Column.from_numpy(key), b2.Column.from_numpy(val), b2.Column.from_numpy(grp)])
    return key, val, grp, t


def test_filter_full_size(b2, big):
    key, val, grp, t = big
    k = b2.col(1, b2.INT64, nullable=False)
    g = b2.col(2, b2.INT32, nullable=False)
    prog = b2.Program([(k >= b2.lit(5_000_000, b2.INT64)) & (k < b2.lit(5_400_000, b2.INT64)) & (g != b2.lit(7, b2.INT32))])
    keep = (val >= 5_000_000) & (val < 5_400_000) & (grp != 7)
    assert b2.filter_count(prog, t) == int(keep.sum())
    out = b2.filter(prog, t)
    assert out.num_rows == int(keep.sum())
    assert np.array_equal(out.column(0).to_numpy()[0], key[keep])       # order preserved
    assert np.array_equal(out.column(1).to_numpy()[0], val[keep])


def test_groupby_full_size(b2, big):
    key, val, grp, t = big
    out = b2.groupby(t, [2], [(b2.AGG_SUM, 1, b2.INT64, 0, 0), (b2.AGG_COUNT_ALL, 1, b2.INT64, 0, 0), (b2.AGG_MAX, 0, b2.INT64, 0, 0)])
    assert out.num_rows == 1000
    g = out.column(0).to_numpy()[0]
    o = np.argsort(g)
    assert np.array_equal(g[o], np.arange(1000, dtype=np.int32))
    # per-group sums stay below 2^53, so the float64 accumulation of np.bincount is exact
    assert np.array_equal(out.column(1).to_numpy()[0][o], np.bincount(grp, weights=val, minlength=1000).astype(np.int64))
    assert np.array_equal(out.column(2).to_numpy()[0][o], np.bincount(grp, minlength=1000))
    mx = np.full(1000, -1, dtype=np.int64)
    np.maximum.at(mx, grp[:2_000_000], key[:2_000_000])                 # max over a prefix bounds the full max from below
    assert np.all(out.column(3).to_numpy()[0][o] >= mx)
    assert int(out.column(3).to_numpy()[0].max()) == int(key.max())


@pytest.mark.parametrize("n,ngroups", [(3_000_000, 1_200_000), (20_000_000, 7_000_003)])
def test_groupby_high_cardinality_radix_regime(b2, n, ngroups):
    """millions of groups: the radix-partitioned shared-memory regime (partitions cut across chunks, tables flushed at
    partition boundaries).  Sums / counts per key equal np.bincount; run twice: the result does not depend on scheduling"""
    rng = np.random.default_rng(n)
    k0 = rng.integers(0, ngroups, n, dtype=np.int64)
    k1 = (k0 % 2557).astype(np.int32)                   # functionally dependent second key (q3's o_orderdate shape)
    val = rng.integers(1, 10_000_000, n, dtype=np.int64)
    t = b2.Table.from_columns([b2.Column.from_numpy(k0), b2.Column.from_numpy(k1), b2.Column.from_numpy(val)])
    want_sum = np.bincount(k0, weights=val, minlength=ngroups).astype(np.int64)     # < 2^53: exact
    want_cnt = np.bincount(k0, minlength=ngroups)
    present = np.flatnonzero(want_cnt)
    for _ in range(2):
        out = b2.groupby(t, [0, 1], [(b2.AGG_SUM, 2, b2.INT64, 0, 0), (b2.AGG_COUNT_ALL, 2, b2.INT64, 0, 0)])
        assert out.num_rows == len(present)
        g = out.column(0).to_numpy()[0]
        o = np.argsort(g)
        assert np.array_equal(g[o], present)
        assert np.array_equal(out.column(1).to_numpy()[0][o], (present % 2557).astype(np.int32))
        assert np.array_equal(out.column(2).to_numpy()[0][o], want_sum[present])
        assert np.array_equal(out.column(3).to_numpy()[0][o], want_cnt[present])


def test_sort_full_size(b2, big):
    key, val, grp, t = big
    n = 30_000_000                                                      # 2 x 12-byte key/value buffers of the radix sort per row
    sub = b2.slice_table(t, 0, n)
    perm = b2.sort_order(sub, [(0, 1, 1)]).to_numpy()[0]
    assert perm.dtype == np.int32 and len(perm) == n
    seen = np.zeros(n, dtype=bool)
    seen[perm] = True
    assert seen.all()                                                   # a permutation
    sk = key[:n][perm]
    assert np.all(sk[1:] >= sk[:-1])                                    # sorted
    ties = sk[1:] == sk[:-1]
    assert np.all(perm[1:][ties] > perm[:-1][ties])                     # stable


def test_hash_partition_full_size(b2, big):
    key, val, grp, t = big
    nparts = 200                                                        # spark.sql.shuffle.partitions default
    got_h = b2.murmur3(t, [0], 42).to_numpy()[0]
    exp_h = np_hash_long(key)
    assert np.array_equal(got_h, exp_h)
    pid = (exp_h.astype(np.int64) % nparts + nparts) % nparts
    n = 30_000_000
    sub = b2.slice_table(t, 0, n)
    out, offs = b2.hash_partition(sub, [0], nparts)
    assert offs == [0] + [int(x) for x in np.cumsum(np.bincount(pid[:n], minlength=nparts))]
    ok, ov = out.column(0).to_numpy()[0], out.column(1).to_numpy()[0]
    opid = (np_hash_long(ok).astype(np.int64) % nparts + nparts) % nparts
    assert np.all(opid[1:] >= opid[:-1])                                # each partition contiguous, in id order
    order = np.argsort(pid[:n], kind="stable")                          # Table.partition keeps input order inside a partition
    assert np.array_equal(ok, key[:n][order]) and np.array_equal(ov, val[:n][order])


def test_join_full_size(b2, big):
    key, val, grp, t = big
    rng = np.random.default_rng(5)
    nb = 1_500_000
    bkey = rng.permutation(3_000_000)[:nb].astype(np.int64)             # distinct build keys, ~half of the probe keys match
    bval = rng.integers(0, 1 << 40, nb).astype(np.int64)
    ht = b2.JoinHashTable(b2.Table.from_columns([b2.Column.from_numpy(bkey)]))
    lm, rm = ht.probe(b2.Table.from_columns([t.column(0)]), b2.JOIN_INNER)
    lut = np.full(3_000_000, -1, dtype=np.int64)
    lut[bkey] = np.arange(nb)
    hit = lut[key]
    l, r = lm.to_numpy()[0], rm.to_numpy()[0]
    assert len(l) == int((hit >= 0).sum())
    assert np.array_equal(key[l], bkey[r])                              # every pair matches
    o = np.argsort(l, kind="stable")
    assert np.array_equal(l[o], np.nonzero(hit >= 0)[0])                # every matching probe row exactly once
    assert int(bval[r].sum()) == int(bval[hit[hit >= 0]].sum())


@pytest.mark.parametrize("limit", [1, 10, 1000])
def test_top_n_full_size(b2, big, limit):
    """GpuTopN over 30 M rows takes the radix-select path (no full sort): same rows, same order as a stable
    lexicographic sort (price desc, group asc, input order on ties)"""
    key, val, grp, t = big
    n = 30_000_000
    sub = b2.slice_table(t, 0, n)
    top = b2.top_n(sub, [(1, 0, 0), (2, 1, 1)], limit)
    assert top.num_rows == limit
    order = np.lexsort((np.arange(n), grp[:n], -val[:n]))[:limit]     # last key is the primary one
    assert np.array_equal(top.column(0).to_numpy()[0], key[:n][order])
    assert np.array_equal(top.column(1).to_numpy()[0], val[:n][order])
    assert np.array_equal(top.column(2).to_numpy()[0], grp[:n][order])


def test_top_n_select_with_nulls_ties_and_skew(b2):
    """the selection must keep every tie of the threshold prefix, honour nulls_first / nulls_last and fall back to the
    full sort when one value dominates"""
    rng = np.random.default_rng(77)
    n = 400_000
    v = rng.integers(-1000, 1000, n).astype(np.int64)
    valid = rng.random(n) > 0.01
    f = rng.standard_normal(n)
    col_v = b2.Column.from_numpy(v, valid=valid)
    t = b2.Table.from_columns([col_v, b2.Column.from_numpy(f), b2.Column.from_numpy(np.arange(n, dtype=np.int32))])
    big_first = np.where(valid, v, np.iinfo(np.int64).min)                 # asc, nulls first
    exp = np.lexsort((np.arange(n), f, big_first))[:50]
    got = b2.top_n(t, [(0, 1, 1), (1, 1, 1)], 50).column(2).to_numpy()[0]
    assert np.array_equal(got, exp)
    last = np.where(valid, -v, np.iinfo(np.int64).max)                     # desc, nulls last
    exp = np.lexsort((np.arange(n), last))[:200]
    got = b2.top_n(t, [(0, 0, 0)], 200).column(2).to_numpy()[0]
    assert np.array_equal(got, exp)
    skew = np.zeros(n, dtype=np.int64); skew[rng.integers(0, n, 20)] = -5   # one dominant value: prefix does not discriminate
    t2 = b2.Table.from_columns([b2.Column.from_numpy(skew), b2.Column.from_numpy(np.arange(n, dtype=np.int32))])
    exp = np.lexsort((np.arange(n), skew))[:100]
    assert np.array_equal(b2.top_n(t2, [(0, 1, 1)], 100).column(1).to_numpy()[0], exp)
