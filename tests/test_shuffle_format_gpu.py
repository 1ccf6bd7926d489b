"""(f1) shuffle wire format: GpuColumnarBatchSerializer-style serialisation of row slices and GpuShuffleCoalesceExec-style
coalesce-on-read, against the independent Python reader/writer of the format (oracle/shuffle_format.py)."""
import numpy as np
import pytest

from oracle import shuffle_format as F
from oracle import spark_cpu as O
from tests import datagen as G

pytestmark = pytest.mark.gpu
TYPES = [(O.BOOL8, 0, 0), (O.INT8, 0, 0), (O.INT16, 0, 0), (O.INT32, 0, 0), (O.INT64, 0, 0), (O.FLOAT32, 0, 0), (O.FLOAT64, 0, 0), (O.DATE32, 0, 0),
         (O.DECIMAL32, 7, 2), (O.DECIMAL64, 12, 2), (O.DECIMAL128, 30, 4), (O.STRING, 0, 0)]


def _cols(rng, n, null_frac=0.15):
    return [G.gen_column(rng, t, n, null_frac=null_frac if i % 3 else 0.0) for i, t in enumerate(TYPES)]


def _same(got_cols, exp_cols):
    for g, e in zip(got_cols, exp_cols):
        assert g.typ[0] == e.typ[0] and g.typ[2] == e.typ[2]
        gl, el = g.to_pylist(), e.to_pylist()
        assert len(gl) == len(el)
        for x, y in zip(gl, el):
            assert (x != x and y != y) or x == y, (g.typ, x, y)


@pytest.mark.parametrize("n", [0, 1, 63, 1000])
def test_serialize_slices_parse_with_the_oracle(b2, n):
    rng = np.random.default_rng(n)
    cols = _cols(rng, n)
    t = G.to_b2_table(b2, cols)
    for a, b in [(0, n), (n // 3, n), (n // 3, 2 * n // 3 + (1 if n else 0)), (5 if n > 5 else 0, 13 if n > 13 else n)]:
        b = min(b, n)
        raw = b2.serialize_table(t, a, b)
        assert len(raw) % 64 == 0
        _same(F.parse(raw), [O.OCol(c.values[a:b], c.valid[a:b], c.typ) for c in cols])


def test_coalesce_on_read_concatenates_in_arrival_order(b2):
    rng = np.random.default_rng(3)
    parts = [_cols(rng, n, nf) for n, nf in [(700, 0.2), (0, 0.0), (13, 0.0), (2500, 0.3), (1, 0.5)]]
    # buffers as three different writers would produce them: the GPU serialiser (slices of a table) and the oracle writer
    bufs = []
    for k, p in enumerate(parts):
        bufs.append(b2.serialize_table(G.to_b2_table(b2, p)) if k % 2 == 0 else np.frombuffer(F.build(p), dtype=np.uint8))
    got = b2.deserialize_concat(bufs)
    exp = [O.OCol(np.concatenate([p[i].values for p in parts]), np.concatenate([p[i].valid for p in parts]), parts[0][i].typ) for i in range(len(TYPES))]
    assert got.num_rows == sum(len(p[0]) for p in parts)
    for i in range(len(TYPES)):
        G.assert_col_equal(got.column(i), exp[i])


def test_hash_partition_serialize_roundtrip(b2):
    """the MULTITHREADED shuffle write/read path in one process: partition -> serialise each slice -> coalesce each reducer's
    slices -> the union of the reducers' rows is the input"""
    rng = np.random.default_rng(9)
    cols = [G.gen_column(rng, (O.INT64, 0, 0), 6000, distinct=300), G.gen_column(rng, (O.STRING, 0, 0), 6000), G.gen_column(rng, (O.DECIMAL64, 12, 2), 6000)]
    batches = [[O.OCol(c.values[a:b], c.valid[a:b], c.typ) for c in cols] for a, b in [(0, 2000), (2000, 5000), (5000, 6000)]]
    nparts = 4
    slices = [[] for _ in range(nparts)]
    for bt in batches:
        part, offs = b2.hash_partition(G.to_b2_table(b2, bt), [0], nparts)
        for p in range(nparts):
            slices[p].append(b2.serialize_table(part, offs[p], offs[p + 1]))
    rows = []
    for p in range(nparts):
        rows += b2.deserialize_concat(slices[p]).to_rows()
    assert G.norm_rows(rows) == G.norm_rows(O.rows_of(cols))


def test_bad_buffers_are_rejected(b2):
    good = b2.serialize_table(G.to_b2_table(b2, [G.gen_column(np.random.default_rng(1), (O.INT32, 0, 0), 100)]))
    bad = good.copy(); bad[0] ^= 0xff
    with pytest.raises(b2.B2Error):
        b2.deserialize_concat([bad])
    with pytest.raises(b2.B2Error):
        b2.deserialize_concat([good[:40]])
