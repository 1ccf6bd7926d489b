"""Run under torchrun (one process per GPU): the multi-GPU parity checks of (e).  Exit code 0 = pass.

 1. NCCL path (b2_exchange, table with STRING columns) vs the oracle's hash partition.
 2. FUSED path (b2_exchange_hash: hash partition + NVLink peer stores) vs the oracle, row sets per destination.
 3. GpuShuffleExchangeExec termination protocol: ranks with different batch counts, one rank with none.
 4. partial -> exchange -> final aggregate plan == single-node group-by.
 5. The bench's strong-scaled q3 plan (hash exchanges on the join keys) on a small instance == numpy restatement.
 6. GpuBroadcastHashJoinExec == shuffled join.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import spark_rapids_b200 as m
from spark_rapids_b200 import execs as E
from oracle import spark_cpu as O
from oracle import spark_hash as H
from oracle import tpch as T
from tests import datagen as G
import bench

m.init(local)
uid = [m.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
comm = m.Comm(uid[0], rank, world)


def shard(r, strings=True):
    rng = np.random.default_rng(1000 + r)
    n = 20000 + 1000 * r
    cols = [G.gen_column(rng, (O.INT64, 0, 0), n, null_frac=0.05, distinct=500), G.gen_column(rng, (O.STRING, 0, 0), n),
            G.gen_column(rng, (O.DECIMAL64, 12, 2), n), G.gen_column(rng, (O.DECIMAL128, 30, 2), n, null_frac=0.2)]
    return cols if strings else [cols[0], cols[2], cols[3]]


def expected_for_me(strings, keys=(0,)):
    exp_cols = None
    for r in range(world):
        cols, eo = H.hash_partition(shard(r, strings), list(keys), world)
        sl = [O.OCol(c.values[eo[rank]:eo[rank + 1]], c.valid[eo[rank]:eo[rank + 1]], c.typ) for c in cols]
        exp_cols = sl if exp_cols is None else [O.OCol(np.concatenate([a.values, b.values]), np.concatenate([a.valid, b.valid]), a.typ) for a, b in zip(exp_cols, sl)]
    return exp_cols


# ---- 1. NCCL path with strings: exact order (source rank order, stable partition)
mine = shard(rank)
t = G.to_b2_table(m, mine)
part, offs = m.hash_partition(t, [0], world)
got = comm.exchange(part, offs)
exp_cols = expected_for_me(True)
assert got.num_rows == len(exp_cols[0]), (got.num_rows, len(exp_cols[0]))
for i in range(4):
    G.assert_col_equal(got.column(i), exp_cols[i])

# ---- 2. fused path: same rows per destination (row order inside a destination is unspecified, like Spark's shuffle read)
fused = comm.fused_ready()
mine_f = shard(rank, False)
tf = G.to_b2_table(m, mine_f)
if fused:
    got_f, anyd = comm.exchange_hash(tf, [0])
    assert anyd
    exp_f = expected_for_me(False)
    assert got_f.num_rows == len(exp_f[0]), (got_f.num_rows, len(exp_f[0]))
    assert G.norm_rows(got_f.to_rows()) == G.norm_rows(O.rows_of(exp_f))
    # SinglePartition: everything lands on rank 0
    got_s, _ = comm.exchange_hash(tf, [])
    total = sum(len(shard(r, False)[0]) for r in range(world))
    assert got_s.num_rows == (total if rank == 0 else 0), (rank, got_s.num_rows, total)
    # a small arena forces the collective grow-and-retry path
    st0 = comm.stats()
    none_t, any2 = comm.exchange_hash(None, [0])
    assert any2 is False and none_t is None

# ---- 3. termination protocol: rank r feeds r batches (rank 0: none at all); every rank must come back
rng = np.random.default_rng(77)
nb = rank
bat = []
for b in range(nb):
    k = O.OCol(rng.integers(0, 1000, 3000).astype(np.int64), np.ones(3000, bool), (O.INT64, 0, 0))
    v = O.OCol(np.full(3000, rank * 100 + b, dtype=np.int64), np.ones(3000, bool), (O.INT64, 0, 0))
    bat.append(G.to_b2_table(m, [k, v]))
ex = E.GpuShuffleExchangeExec(E.GpuBatchSource(bat), [0], comm, world)
rows_got = 0
nbat = 0
for tb in ex:
    rows_got += tb.num_rows
    nbat += 1
    keys = tb.column(0).to_pylist()[:64]
    if keys:   # every row that arrived here hashes to this rank (Spark Murmur3 pmod world)
        kc = O.OCol(np.array(keys, dtype=np.int64), np.ones(len(keys), bool), (O.INT64, 0, 0))
        assert all(int(p) == rank for p in H.partition_ids([kc], world)), "row delivered to the wrong rank"
allrows = [None] * world
dist.all_gather_object(allrows, rows_got)
assert sum(allrows) == 3000 * sum(range(world)), (allrows,)
allb = [None] * world
dist.all_gather_object(allb, nbat)
assert len(set(allb)) == 1, ("every rank must see the same number of exchange rounds", allb)

# ---- 4. partial group-by -> exchange on the key -> final group-by; union over ranks == single-node group-by
specs = [(O.AGG_SUM, 1, O.DECIMAL128, 2, 22), (O.AGG_COUNT, 1), (O.AGG_COUNT_ALL, 0)]
pre = [G.b2_expr_col(m, 0, mine[0]), G.b2_expr_col(m, 2, mine[2])]
partial = E.GpuHashAggregateExec(E.GpuBatchSource([t]), [0], specs, pre_project=pre)
final = E.GpuHashAggregateExec(E.GpuShuffleExchangeExec(partial, [0], comm, world), [0], specs, mode="final")
res = final.collect()
rows = res.to_rows() if res is not None else []
allrows = [None] * world
dist.all_gather_object(allrows, rows)
if rank == 0:
    merged = [r for rs in allrows for r in rs]
    every = [shard(r) for r in range(world)]
    cat = [O.OCol(np.concatenate([s[i].values for s in every]), np.concatenate([s[i].valid for s in every]), every[0][i].typ) for i in range(4)]
    exp = O.rows_of(O.groupby_cols([cat[0], cat[2]], [0], specs))
    assert G.norm_rows(merged) == G.norm_rows(exp), (len(merged), len(exp))

# ---- 5. the bench's q3 plan, strong-scaled over the ranks, small instance
sf = 0.05
chunks = bench.q3_host_chunks(sf, rank, world)
progs = bench.q3_programs(m)
dev = bench.q3_device_batches(m, chunks)
root, nodes = bench.build_q3_plan(m, E, progs, {tn: E.GpuBatchSource(dev[tn]) for tn in bench.Q3_SCHEMA}, comm, rank, world)
q3 = bench.q3_rows_of(root.collect())
if rank == 0:
    exp = T.q3_expected(sf, 42, threads=4)
    assert [(r[1], r[2]) for r in q3] == [(r[1], r[2]) for r in exp], (q3, exp)
    assert sorted(q3) == sorted(exp)
else:
    assert q3 == []
xs = comm.stats()

# ---- 6. broadcast hash join == shuffled hash join
i64 = (O.INT64, 0, 0)
rngb = np.random.default_rng(500 + rank)
small = [O.OCol(np.arange(rank * 300, rank * 300 + 300, dtype=np.int64), np.ones(300, bool), i64), O.OCol(rngb.integers(0, 9, 300).astype(np.int64), np.ones(300, bool), i64)]
big = [O.OCol(rngb.integers(0, 300 * world, 8000).astype(np.int64), np.ones(8000, bool), i64), O.OCol(np.arange(8000, dtype=np.int64) + rank * 10**6, np.ones(8000, bool), i64)]
bj = E.GpuBroadcastHashJoinExec([0], [0], m.JOIN_INNER, E.GpuBatchSource([G.to_b2_table(m, big)]), E.GpuBatchSource([G.to_b2_table(m, small)]), comm, rank, world)
brow = bj.collect().to_rows()
assert len(brow) == 8000, len(brow)     # every stream key 0..300*world-1 finds exactly one build row somewhere
for k, _, bk, _ in brow[:200]:
    assert k == bk

if rank == 0:
    print("exchange_check ok: world=%d fused=%s rows_out=%d q3_top1=%s exchange_bytes_sent=%d" % (world, fused, got.num_rows, q3[0], xs["bytes_sent"]))
comm.close()
dist.barrier()
dist.destroy_process_group()
