"""Run under torchrun (one process per GPU): checks b2_exchange (NCCL all-to-all of a hash-partitioned
table) and the partial -> exchange -> final aggregate plan against the oracle.  Exit code 0 = pass."""
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
from tests import datagen as G

m.init(local)
uid = [m.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
comm = m.Comm(uid[0], rank, world)


def shard(r):
    rng = np.random.default_rng(1000 + r)
    n = 20000 + 1000 * r
    return [G.gen_column(rng, (O.INT64, 0, 0), n, null_frac=0.05, distinct=500), G.gen_column(rng, (O.STRING, 0, 0), n),
            G.gen_column(rng, (O.DECIMAL64, 12, 2), n), G.gen_column(rng, (O.DECIMAL128, 30, 2), n, null_frac=0.2)]


mine = shard(rank)
t = G.to_b2_table(m, mine)
part, offs = m.hash_partition(t, [0], world)
got = comm.exchange(part, offs)
# expected: rows of every shard whose partition id is my rank, in source-rank order
exp_cols = None
for r in range(world):
    cols, eo = H.hash_partition(shard(r), [0], world)
    sl = [O.OCol(c.values[eo[rank]:eo[rank + 1]], c.valid[eo[rank]:eo[rank + 1]], c.typ) for c in cols]
    exp_cols = sl if exp_cols is None else [O.OCol(np.concatenate([a.values, b.values]), np.concatenate([a.valid, b.valid]), a.typ) for a, b in zip(exp_cols, sl)]
assert got.num_rows == len(exp_cols[0]), (got.num_rows, len(exp_cols[0]))
for i in range(4):
    G.assert_col_equal(got.column(i), exp_cols[i])

# plan: partial group-by -> exchange on the key -> final group-by; union over ranks == single-node group-by
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
    print("exchange_check ok: world=%d rows_out=%d groups=%d" % (world, got.num_rows, len(merged)))
comm.close()
dist.barrier()
dist.destroy_process_group()
